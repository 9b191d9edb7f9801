// kernels_dna.cu -- 4-state (DNA) GAMMA likelihood kernels for sm_100a.
//
// Replaces, for DNA partitions, the per-site loops of
//   newviewGTRGAMMA_AVX   avxLikelihood.c:64-318          (all three tip cases + scaling)
//   makeP                 newviewGenericSpecial.c:78-132  (folded into the kernel prologue)
//   evaluateGTRGAMMA      evaluateGenericSpecial.c:1879-1985, calcDiagptable :80-119
//   sumGAMMA              makenewzGenericSpecial.c:1798-1848
//   coreGTRGAMMA          makenewzGenericSpecial.c:2309-2393
//
// B200 restructuring of HOT LOOP A/B: the traversal loop lives INSIDE the kernel.
// A CTA owns a tile of sites for the whole traversal and walks the descriptor entry by entry.
// Thread t of the CTA owns (site = t/4 + 64*u, Gamma category = t%4) for every entry, so a parent
// entry reads back exactly the 32 B this same thread stored for the child: no inter-thread
// dependency, no barrier for data, and the child CLV is still in L2 when it is consumed.
// One 256-bit load per child and one 256-bit store per parent (LDG/STG.E.ENL2.256); a warp covers
// 1 KB of contiguous CLV.  The only block barrier per entry publishes the 24 exponentials of the
// two P(t) eigen-factor tables (double-buffered in shared memory).
#include "engine.cuh"

namespace exb {

namespace {

struct DnaShared {
  double EIGN[4];
  double rates[4];
  double EI[16];
  double EV[16];
  double tip[64];        // tipVector[16 codes][4]
  double d[2][2][16];    // [buffer][side q/r][c*4+k] = exp(EIGN[k] * rate[c] * lz)
  double red[8];
};

__device__ __forceinline__ void ld256(const double *p, double v[4])
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(double *p, const double v[4])
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

// a[l] = sum_j P[l][j] x[j] with P[l][j] = d[j] * EI[l][j], P[l][0] = 1 (makeP, newviewGenericSpecial.c:118-129)
__device__ __forceinline__ void apply_p(const double x[4], const double *d /*4, d[0] unused*/,
                                        const double *EI, double a[4])
{
  const double t1 = d[1] * x[1], t2 = d[2] * x[2], t3 = d[3] * x[3];
  #pragma unroll
  for (int l = 0; l < 4; l++)
    a[l] = x[0] + EI[l * 4 + 1] * t1 + EI[l * 4 + 2] * t2 + EI[l * 4 + 3] * t3;
}

__device__ __forceinline__ void load_model(DnaShared &sh, const double *model)
{
  // model: EIGN[4] rates[4] EI[16] EV[16] tip[64]  (kDnaModelDoubles)
  for (int i = threadIdx.x; i < kDnaModelDoubles; i += blockDim.x) {
    const double v = model[i];
    if (i < 4) sh.EIGN[i] = v;
    else if (i < 8) sh.rates[i - 4] = v;
    else if (i < 24) sh.EI[i - 8] = v;
    else if (i < 40) sh.EV[i - 24] = v;
    else sh.tip[i - 40] = v;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// Fused traversal: newviews [first, count) for this tile, then (optionally) evaluate or the sum table
// at the branch described by entry 0.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3)
traverse_dna_kernel(const TraverseArgs a)
{
  __shared__ DnaShared sh;
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;

  const int tid = threadIdx.x, lane = tid & 31;
  const int c = tid & 3;                       // Gamma category owned by this thread
  const int slot = a.nbranches > 1 ? tile.part : 0;
  const size_t N = a.N;
  const int npass = (tile.nsites + 63) >> 6;   // 64 sites per pass (256 threads / 4 categories)

  load_model(sh, a.models + (size_t)tile.part_cls * kDnaModelDoubles);
  __syncthreads();

  // ---------------- newview entries ----------------
  for (int e = a.first; e < a.count; e++) {
    const DevEntry en = a.entries[e];
    const int buf = e & 1;
    if (tid < 32) {                            // 2 sides x 4 categories x 4 eigenvalues
      const int side = tid >> 4, cc = (tid >> 2) & 3, k = tid & 3;
      const double lz = side ? a.lzr[(size_t)e * a.nbranches + slot] : a.lzq[(size_t)e * a.nbranches + slot];
      sh.d[buf][side][cc * 4 + k] = k ? exp(sh.rates[cc] * (sh.EIGN[k] * lz)) : 1.0;
    }
    __syncthreads();
    const double *dq = &sh.d[buf][0][c * 4], *dr = &sh.d[buf][1][c * 4];

    const double *x1p = nullptr, *x2p = nullptr;
    const uint8_t *t1p = nullptr, *t2p = nullptr;
    if (en.tipCase == 2) x1p = a.clv + (size_t)(en.q - a.mxtips - 1) * N * 16;
    else                 t1p = a.tips + (size_t)(en.q - 1) * N;
    if (en.tipCase == 0) t2p = a.tips + (size_t)(en.r - 1) * N;
    else                 x2p = a.clv + (size_t)(en.r - a.mxtips - 1) * N * 16;
    double *x3p = a.clv + (size_t)(en.p - a.mxtips - 1) * N * 16;
    const bool may_scale = en.tipCase != 0;    // TIP_TIP is never rescaled (avxLikelihood.c:87-160)

    unsigned int cnt = 0;
    for (int u = 0; u < npass; u++) {
      const int i = (u << 6) + (tid >> 2);
      const bool valid = i < (int)tile.nsites;
      const size_t site = (size_t)tile.site0 + (valid ? i : 0);
      double x1[4], x2[4], al[4], ar[4], v[4];
      if (x1p) ld256(x1p + site * 16 + c * 4, x1);
      else { const double *tv = &sh.tip[4 * t1p[site]]; x1[0] = tv[0]; x1[1] = tv[1]; x1[2] = tv[2]; x1[3] = tv[3]; }
      if (x2p) ld256(x2p + site * 16 + c * 4, x2);
      else { const double *tv = &sh.tip[4 * t2p[site]]; x2[0] = tv[0]; x2[1] = tv[1]; x2[2] = tv[2]; x2[3] = tv[3]; }
      apply_p(x1, dq, sh.EI, al);
      apply_p(x2, dr, sh.EI, ar);
      #pragma unroll
      for (int m = 0; m < 4; m++) v[m] = 0.0;
      #pragma unroll
      for (int l = 0; l < 4; l++) {
        const double ab = al[l] * ar[l];
        #pragma unroll
        for (int m = 0; m < 4; m++) v[m] += ab * sh.EV[l * 4 + m];
      }
      if (may_scale) {
        const bool small = fabs(v[0]) < kMinLikelihood && fabs(v[1]) < kMinLikelihood &&
                           fabs(v[2]) < kMinLikelihood && fabs(v[3]) < kMinLikelihood;
        const unsigned int b = __ballot_sync(0xffffffffu, small);
        if (((b >> (lane & ~3)) & 0xFu) == 0xFu) {          // all 16 values of the site are tiny
          #pragma unroll
          for (int m = 0; m < 4; m++) v[m] *= kTwoToThe256;
          if (c == 0 && valid) cnt += (unsigned int)a.wgt[site];
        }
      }
      if (valid) st256(x3p + site * 16 + c * 4, v);
    }
    if (may_scale && __any_sync(0xffffffffu, cnt != 0)) {
      const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
    }
  }

  if (a.mode == MODE_NEWVIEW) return;

  // ---------------- branch (entry 0): evaluate or sum table ----------------
  const DevEntry br = a.entries[0];
  const bool ptip = br.p <= a.mxtips, qtip = br.q <= a.mxtips;
  const double *xap = ptip ? nullptr : a.clv + (size_t)(br.p - a.mxtips - 1) * N * 16;
  const double *xbp = qtip ? nullptr : a.clv + (size_t)(br.q - a.mxtips - 1) * N * 16;
  const uint8_t *tap = ptip ? a.tips + (size_t)(br.p - 1) * N : nullptr;
  const uint8_t *tbp = qtip ? a.tips + (size_t)(br.q - 1) * N : nullptr;

  double diag[4] = {1.0, 1.0, 1.0, 1.0};
  if (a.mode == MODE_EVALUATE) {
    __syncthreads();                           // everyone is done with d[][] of the last entry
    if (tid < 16) {
      const int cc = tid >> 2, k = tid & 3;
      const double lz = a.lzq[slot];           // entry 0 carries the branch value
      sh.d[0][0][cc * 4 + k] = k ? exp(sh.rates[cc] * (sh.EIGN[k] * lz)) : 1.0;
    }
    __syncthreads();
    #pragma unroll
    for (int k = 0; k < 4; k++) diag[k] = sh.d[0][0][c * 4 + k];
  }

  double acc = 0.0;
  for (int u = 0; u < npass; u++) {
    const int i = (u << 6) + (tid >> 2);
    const bool valid = i < (int)tile.nsites;
    const size_t site = (size_t)tile.site0 + (valid ? i : 0);
    double xa[4], xb[4];
    if (xap) ld256(xap + site * 16 + c * 4, xa);
    else { const double *tv = &sh.tip[4 * tap[site]]; xa[0] = tv[0]; xa[1] = tv[1]; xa[2] = tv[2]; xa[3] = tv[3]; }
    if (xbp) ld256(xbp + site * 16 + c * 4, xb);
    else { const double *tv = &sh.tip[4 * tbp[site]]; xb[0] = tv[0]; xb[1] = tv[1]; xb[2] = tv[2]; xb[3] = tv[3]; }
    if (a.mode == MODE_SUM) {
      double s[4];
      #pragma unroll
      for (int k = 0; k < 4; k++) s[k] = xa[k] * xb[k];
      if (valid) st256(a.sum + site * 16 + c * 4, s);
    } else {
      double term = 0.0;
      #pragma unroll
      for (int k = 0; k < 4; k++) term += xa[k] * xb[k] * diag[k];
      term += __shfl_xor_sync(0xffffffffu, term, 1);
      term += __shfl_xor_sync(0xffffffffu, term, 2);
      if (c == 0 && valid) acc += (double)a.wgt[site] * log(0.25 * fabs(term));
    }
  }
  if (a.mode == MODE_EVALUATE) {
    const double tot = block_sum(acc, sh.red);
    if (tid == 0) { a.tile_out[2 * (size_t)blockIdx.x] = tot; a.tile_out[2 * (size_t)blockIdx.x + 1] = 0.0; }
  }
}

// ------------------------------------------------------------------------------------------------------
// core: first and second derivative of lnL w.r.t. lz over the sum table (one launch per NR iteration).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 4)
core_dna_kernel(const CoreArgs a)
{
  __shared__ double d0[16], d1[16], d2[16], red[8];
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  const int tid = threadIdx.x, c = tid & 3;
  const int slot = a.nbranches > 1 ? tile.part : 0;
  const double *model = a.models + (size_t)tile.part_cls * kDnaModelDoubles;
  if (tid < 16) {
    const int cc = tid >> 2, k = tid & 3;
    const double ki = model[4 + cc], ek = model[k], lz = a.lz[slot];
    d0[tid] = k ? exp(ek * ki * lz) : 1.0;
    d1[tid] = k ? ek * ki : 0.0;
    d2[tid] = k ? ek * ek * (ki * ki) : 0.0;
  }
  __syncthreads();
  double e0[4], e1[4], e2[4];
  #pragma unroll
  for (int k = 0; k < 4; k++) { e0[k] = d0[c * 4 + k]; e1[k] = d1[c * 4 + k]; e2[k] = d2[c * 4 + k]; }

  double acc1 = 0.0, acc2 = 0.0;
  const int npass = (tile.nsites + 63) >> 6;
  for (int u = 0; u < npass; u++) {
    const int i = (u << 6) + (tid >> 2);
    const bool valid = i < (int)tile.nsites;
    const size_t site = (size_t)tile.site0 + (valid ? i : 0);
    double s[4];
    ld256(a.sum + site * 16 + c * 4, s);
    double L = 0.0, L1 = 0.0, L2 = 0.0;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
      const double t = e0[k] * s[k];
      L += t; L1 += t * e1[k]; L2 += t * e2[k];
    }
    L  += __shfl_xor_sync(0xffffffffu, L, 1);   L  += __shfl_xor_sync(0xffffffffu, L, 2);
    L1 += __shfl_xor_sync(0xffffffffu, L1, 1);  L1 += __shfl_xor_sync(0xffffffffu, L1, 2);
    L2 += __shfl_xor_sync(0xffffffffu, L2, 1);  L2 += __shfl_xor_sync(0xffffffffu, L2, 2);
    if (c == 0 && valid) {
      const double inv = 1.0 / fabs(L);
      const double g = L1 * inv, h = L2 * inv;
      const double w = (double)a.wgt[site];
      acc1 += w * g;
      acc2 += w * (h - g * g);
    }
  }
  const double t1 = block_sum(acc1, red);
  const double t2 = block_sum(acc2, red);
  if (tid == 0) { a.tile_out[2 * (size_t)blockIdx.x] = t1; a.tile_out[2 * (size_t)blockIdx.x + 1] = t2; }
}

void launch_traverse_dna(const TraverseArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles > 0) traverse_dna_kernel<<<ntiles, kThreads, 0, s>>>(a);
}
void launch_core_dna(const CoreArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles > 0) core_dna_kernel<<<ntiles, kThreads, 0, s>>>(a);
}

}  // namespace exb
