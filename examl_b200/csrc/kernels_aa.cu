// kernels_aa.cu -- 20-state (protein) GAMMA likelihood kernels for sm_100a.
//
// Replaces, for AA partitions, the per-site loops of
//   newviewGTRGAMMAPROT_AVX  avxLikelihood.c:1312-1800
//   evaluateGTRGAMMAPROT     evaluateGenericSpecial.c:1393-1461
//   sumGAMMAPROT             makenewzGenericSpecial.c:2083-2153
//   coreGTRGAMMAPROT         makenewzGenericSpecial.c:2581-2655
//
// Same structure as kernels_dna.cu: the traversal loop is inside the kernel, thread t owns
// (site = t/4 + 64*u, category = t%4) for every entry and therefore only ever reads CLV words it
// wrote itself.  All arithmetic is FP64 (the reference's precision; tcgen05 has no f64 kind, see
// DESIGN.md).  P(t) is never materialised: a = EI * (d (.) x) with the category-independent EI
// broadcast from shared memory (every lane of a warp reads the same word: conflict-free), d the 19
// exponentials of the (entry, side, category).
#include "engine.cuh"

namespace exb {

namespace {

constexpr int S = 20;
constexpr int SPAN = 80;

struct AaShared {
  double EI[400];
  double EV[400];
  double tip[460];        // tipVector[23 codes][20]
  double EIGN[20];
  double rates[4];
  double d[2][2][80];     // [buffer][side][c*20+k]
  double red[8];
};

__device__ __forceinline__ void ld256(const double *p, double *v)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(double *p, const double *v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

__device__ __forceinline__ void load_vec(const double *clv, const uint8_t *tips, const double *tipv, size_t site, int c,
                                         double x[S])
{
  if (clv) {
    const double *p = clv + site * SPAN + c * S;
    #pragma unroll
    for (int j = 0; j < S; j += 4) ld256(p + j, x + j);
  } else {
    const double *tv = tipv + S * tips[site];
    #pragma unroll
    for (int j = 0; j < S; j++) x[j] = tv[j];
  }
}

__device__ __forceinline__ void load_model(AaShared &sh, const double *model)
{
  // model: EIGN[20] rates[4] EI[400] EV[400] tip[460]
  for (int i = threadIdx.x; i < kAaModelDoubles; i += blockDim.x) {
    const double v = model[i];
    if (i < 20) sh.EIGN[i] = v;
    else if (i < 24) sh.rates[i - 20] = v;
    else if (i < 424) sh.EI[i - 24] = v;
    else if (i < 824) sh.EV[i - 424] = v;
    else sh.tip[i - 824] = v;
  }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
traverse_aa_kernel(const TraverseArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AaShared &sh = *reinterpret_cast<AaShared *>(smem_raw);
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;

  const int tid = threadIdx.x, lane = tid & 31;
  const int c = tid & 3;
  const int slot = a.nbranches > 1 ? tile.part : 0;
  const size_t N = a.N;
  const int npass = (tile.nsites + 63) >> 6;

  load_model(sh, a.models + (size_t)tile.part_cls * kAaModelDoubles);
  __syncthreads();

  for (int e = a.first; e < a.count; e++) {
    const DevEntry en = a.entries[e];
    const int buf = e & 1;
    if (tid < 160) {                              // 2 sides x 4 categories x 20 eigenvalues
      const int side = tid / 80, r = tid % 80, cc = r / 20, k = r % 20;
      const double lz = side ? a.lzr[(size_t)e * a.nbranches + slot] : a.lzq[(size_t)e * a.nbranches + slot];
      sh.d[buf][side][r] = k ? exp(sh.rates[cc] * (sh.EIGN[k] * lz)) : 1.0;
    }
    __syncthreads();
    const double *dq = &sh.d[buf][0][c * S], *dr = &sh.d[buf][1][c * S];

    const double *x1p = nullptr, *x2p = nullptr;
    const uint8_t *t1p = nullptr, *t2p = nullptr;
    if (en.tipCase == 2) x1p = a.clv + (size_t)(en.q - a.mxtips - 1) * N * SPAN;
    else                 t1p = a.tips + (size_t)(en.q - 1) * N;
    if (en.tipCase == 0) t2p = a.tips + (size_t)(en.r - 1) * N;
    else                 x2p = a.clv + (size_t)(en.r - a.mxtips - 1) * N * SPAN;
    double *x3p = a.clv + (size_t)(en.p - a.mxtips - 1) * N * SPAN;
    const bool may_scale = en.tipCase != 0;

    unsigned int cnt = 0;
    for (int u = 0; u < npass; u++) {
      const int i = (u << 6) + (tid >> 2);
      const bool valid = i < (int)tile.nsites;
      const size_t site = (size_t)tile.site0 + (valid ? i : 0);
      double t[S], ab[S];
      // left child: ab[l] = sum_j EI[l][j] * (dq[j] * x1[j])      (EI[l][0] = 1, dq[0] = 1)
      load_vec(x1p, t1p, sh.tip, site, c, t);
      #pragma unroll
      for (int j = 1; j < S; j++) t[j] *= dq[j];
      #pragma unroll
      for (int l = 0; l < S; l++) {
        double s = t[0];
        #pragma unroll
        for (int j = 1; j < S; j++) s += sh.EI[l * S + j] * t[j];
        ab[l] = s;
      }
      // right child, multiplied in as soon as each component is known
      load_vec(x2p, t2p, sh.tip, site, c, t);
      #pragma unroll
      for (int j = 1; j < S; j++) t[j] *= dr[j];
      #pragma unroll
      for (int l = 0; l < S; l++) {
        double s = t[0];
        #pragma unroll
        for (int j = 1; j < S; j++) s += sh.EI[l * S + j] * t[j];
        ab[l] *= s;
      }
      // back to the eigen-basis: v[m] = sum_l ab[l] * EV[l][m]   (reuses t)
      #pragma unroll
      for (int m = 0; m < S; m++) t[m] = 0.0;
      #pragma unroll
      for (int l = 0; l < S; l++) {
        #pragma unroll
        for (int m = 0; m < S; m++) t[m] += ab[l] * sh.EV[l * S + m];
      }
      if (may_scale) {
        bool small = true;
        #pragma unroll
        for (int m = 0; m < S; m++) small = small && (fabs(t[m]) < kMinLikelihood);
        const unsigned int b = __ballot_sync(0xffffffffu, small);
        if (((b >> (lane & ~3)) & 0xFu) == 0xFu) {
          #pragma unroll
          for (int m = 0; m < S; m++) t[m] *= kTwoToThe256;
          if (c == 0 && valid) cnt += (unsigned int)a.wgt[site];
        }
      }
      if (valid) {
        double *dst = x3p + site * SPAN + c * S;
        #pragma unroll
        for (int m = 0; m < S; m += 4) st256(dst + m, t + m);
      }
    }
    if (may_scale && __any_sync(0xffffffffu, cnt != 0)) {
      const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
    }
  }

  if (a.mode == MODE_NEWVIEW) return;

  const DevEntry br = a.entries[0];
  const bool ptip = br.p <= a.mxtips, qtip = br.q <= a.mxtips;
  const double *xap = ptip ? nullptr : a.clv + (size_t)(br.p - a.mxtips - 1) * N * SPAN;
  const double *xbp = qtip ? nullptr : a.clv + (size_t)(br.q - a.mxtips - 1) * N * SPAN;
  const uint8_t *tap = ptip ? a.tips + (size_t)(br.p - 1) * N : nullptr;
  const uint8_t *tbp = qtip ? a.tips + (size_t)(br.q - 1) * N : nullptr;

  if (a.mode == MODE_EVALUATE) {
    __syncthreads();
    if (tid < 80) {
      const int cc = tid / 20, k = tid % 20;
      const double lz = a.lzq[slot];
      sh.d[0][0][tid] = k ? exp(sh.rates[cc] * (sh.EIGN[k] * lz)) : 1.0;
    }
    __syncthreads();
  }
  const double *diag = &sh.d[0][0][c * S];

  double acc = 0.0;
  for (int u = 0; u < npass; u++) {
    const int i = (u << 6) + (tid >> 2);
    const bool valid = i < (int)tile.nsites;
    const size_t site = (size_t)tile.site0 + (valid ? i : 0);
    double xa[S], xb[S];
    load_vec(xap, tap, sh.tip, site, c, xa);
    load_vec(xbp, tbp, sh.tip, site, c, xb);
    if (a.mode == MODE_SUM) {
      #pragma unroll
      for (int k = 0; k < S; k++) xa[k] *= xb[k];
      if (valid) {
        double *dst = a.sum + site * SPAN + c * S;
        #pragma unroll
        for (int k = 0; k < S; k += 4) st256(dst + k, xa + k);
      }
    } else {
      double term = 0.0;
      #pragma unroll
      for (int k = 0; k < S; k++) term += xa[k] * xb[k] * diag[k];
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

__global__ void __launch_bounds__(kThreads, 2)
core_aa_kernel(const CoreArgs a)
{
  __shared__ double d0[80], d1[80], d2[80], red[8];
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  const int tid = threadIdx.x, c = tid & 3;
  const int slot = a.nbranches > 1 ? tile.part : 0;
  const double *model = a.models + (size_t)tile.part_cls * kAaModelDoubles;
  if (tid < 80) {
    const int cc = tid / 20, k = tid % 20;
    const double ki = model[20 + cc], ek = model[k], lz = a.lz[slot];
    d0[tid] = k ? exp(ek * ki * lz) : 1.0;
    d1[tid] = k ? ek * ki : 0.0;
    d2[tid] = k ? ek * ek * (ki * ki) : 0.0;
  }
  __syncthreads();
  double acc1 = 0.0, acc2 = 0.0;
  const int npass = (tile.nsites + 63) >> 6;
  for (int u = 0; u < npass; u++) {
    const int i = (u << 6) + (tid >> 2);
    const bool valid = i < (int)tile.nsites;
    const size_t site = (size_t)tile.site0 + (valid ? i : 0);
    double s[S];
    const double *p = a.sum + site * SPAN + c * S;
    #pragma unroll
    for (int j = 0; j < S; j += 4) ld256(p + j, s + j);
    double L = 0.0, L1 = 0.0, L2 = 0.0;
    #pragma unroll
    for (int k = 0; k < S; k++) {
      const double t = d0[c * S + k] * s[k];
      L += t; L1 += t * d1[c * S + k]; L2 += t * d2[c * S + k];
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

void launch_traverse_aa(const TraverseArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles <= 0) return;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(traverse_aa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AaShared));
    configured = true;
  }
  traverse_aa_kernel<<<ntiles, kThreads, sizeof(AaShared), s>>>(a);
}
void launch_core_aa(const CoreArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles > 0) core_aa_kernel<<<ntiles, kThreads, 0, s>>>(a);
}

}  // namespace exb
