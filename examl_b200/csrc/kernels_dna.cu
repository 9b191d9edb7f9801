// kernels_dna.cu -- 4-state (DNA) GAMMA likelihood kernels for sm_100a.
//
// Replaces, for DNA partitions, the per-site loops of
//   newviewGTRGAMMA_AVX   avxLikelihood.c:64-318          (all three tip cases + scaling)
//   makeP                 newviewGenericSpecial.c:78-132  (dtab_kernel: the exponentials; the EI factor
//                                                          is applied inside the contraction)
//   evaluateGTRGAMMA      evaluateGenericSpecial.c:1879-1985, calcDiagptable :80-119
//   sumGAMMA              makenewzGenericSpecial.c:1798-1848
//   coreGTRGAMMA          makenewzGenericSpecial.c:2309-2393
//
// B200 restructuring of HOT LOOP A/B: the traversal loop lives INSIDE the kernel.
// A CTA owns a tile of sites for the whole traversal and walks the descriptor entry by entry.
// Thread t owns (site = t/4 + 64*s, Gamma category = t%4), s < NS, for every entry, so
//   * a parent entry only ever consumes CLV words this same thread produced: no barrier, no fence;
//   * the CLV of the entry just computed stays in registers ("carry"): in a post-order descriptor
//     the inner child r of every entry is the previous entry, so ~3/4 of all child reads never
//     leave the register file; the rest are L2 hits (the child was written microseconds ago);
//   * every parent CLV is still written to HBM once (it must persist for later partial traversals):
//     one 256-bit store per (site, category), a warp covers 1 KB contiguous.
// HBM traffic of a full traversal is therefore ~1 write per CLV + the tip bytes, about half of the
// algorithmic read+write bytes of a per-node implementation.
// Eigen-basis constants EI/EV reach the DFMAs as immediate constant-bank operands when the class
// holds one partition (kernel parameters), else from shared memory; the per-(entry, category)
// exponentials come from the d-table built once per call by dtab_kernel (no block barrier in the
// entry loop).
#include "engine.cuh"
#include "core_tile.cuh"
#include "stage.cuh"
#include "finish.cuh"
#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

namespace exb {

static_assert(kThreads == 32 * kDnaWarps, "score_out slots per tile");

namespace {

struct DnaShared {
  double tip[64];        // tipVector[16 codes][4]
  double EI[16];
  double EV[16];
  double red[32];
};

__device__ __forceinline__ void ld256(const double *p, double v[4])
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
// read-only data (d-table): may be cached / reordered freely
__device__ __forceinline__ void ld256_ro(const double *p, double v[4])
{
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ void st256(double *p, const double v[4])
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

}  // namespace

// |v| < 2^-256 for all four values, decided on the integer pipe (the FP64 pipe is the busy one): the biased exponent
// of 2^-256 is 767 = 0x2FF, so fabs(v) < minlikelihood  <=>  (high word & 0x7fffffff) < 0x2FF00000 (NaN: false, like <).
__device__ __forceinline__ bool all_tiny(const double (&v)[4])
{
  const unsigned int h0 = (unsigned int)__double2hiint(v[0]) & 0x7fffffffu, h1 = (unsigned int)__double2hiint(v[1]) & 0x7fffffffu,
                     h2 = (unsigned int)__double2hiint(v[2]) & 0x7fffffffu, h3 = (unsigned int)__double2hiint(v[3]) & 0x7fffffffu;
  return max(max(h0, h1), max(h2, h3)) < 0x2FF00000u;
}

// cp.async (LDGSTS): 32 bytes global -> this thread's slot of the prefetch ring, no registers, L2-only caching
__device__ __forceinline__ void async_copy32(void *smem_dst, const void *gsrc)
{
  const uint32_t d = stage_smem_u32(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + 16u), "l"(reinterpret_cast<const char *>(gsrc) + 16) : "memory");
}
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void async_wait_all_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// One traversal entry for the NS sites of a thread, compiled once per combination of child sources so that the
// NS loop carries no branch on tip case / carry / scaling.
//   source: 0 = tip (staged code -> tip vector in shared memory), 1 = carry registers,
//           2 = inner CLV that was brought into this thread's slot of the prefetch ring by cp.async,
//           3 = inner CLV read straight from global memory (the second of two loaded children: rare)
// Loop order: every EI / EV constant is fetched once per entry and used for all NS sites (constant-outer, site-inner;
// the compiler keeps them in uniform registers), in stages that keep the live set inside the register budget:
//   (A) gather the children, d (.) x      (B) row l of EI for all sites -> ab[s][l] = a_q[l] * a_r[l]
//   (C) column m of EV for all sites -> v[s][m]
template <int QS, int RS, int NS, bool PARAM>
__device__ __forceinline__ unsigned int dna_entry(double (&carry)[NS][4], const TraverseArgs &a, const DnaConsts &kc, const DnaShared &sh,
                                                  const uint8_t *stips, const SEntry &en, const double *sd, const double *ring,
                                                  const size_t (&eoff)[NS], const bool (&valid)[NS], const int (&lsite)[NS], int c, int lane,
                                                  bool store)
{
#define EI_(i) (PARAM ? kc.EI[i] : sh.EI[i])
#define EV_(i) (PARAM ? kc.EV[i] : sh.EV[i])
  const double dq1 = sd[1], dq2 = sd[2], dq3 = sd[3], dr1 = sd[5], dr2 = sd[6], dr3 = sd[7];   // [side][k], d[0] = 1
  double qa[NS][4], ra[NS][4];         // x[0], d[1] x[1], d[2] x[2], d[3] x[3]
  #pragma unroll
  for (int s = 0; s < NS; s++) {                      // (A)
    double x1[4], x2[4];
    if (QS == 1) { x1[0] = carry[s][0]; x1[1] = carry[s][1]; x1[2] = carry[s][2]; x1[3] = carry[s][3]; }
    else if (QS == 2) { const double *r = ring + (size_t)s * kThreads * 4; x1[0] = r[0]; x1[1] = r[1]; x1[2] = r[2]; x1[3] = r[3]; }
    else if (QS == 3) ld256(a.clv + en.qoff + ((en.flags & kFlagGapA) ? (size_t)(c * 4) : eoff[s]), x1);
    else { const double *tv = &sh.tip[4 * stips[en.qoff + lsite[s]]]; x1[0] = tv[0]; x1[1] = tv[1]; x1[2] = tv[2]; x1[3] = tv[3]; }
    if (RS == 1) { x2[0] = carry[s][0]; x2[1] = carry[s][1]; x2[2] = carry[s][2]; x2[3] = carry[s][3]; }
    else if (RS == 2) { const double *r = ring + (size_t)s * kThreads * 4; x2[0] = r[0]; x2[1] = r[1]; x2[2] = r[2]; x2[3] = r[3]; }
    else if (RS == 3) ld256(a.clv + en.roff + ((en.flags & kFlagGapB) ? (size_t)(c * 4) : eoff[s]), x2);
    else { const double *tv = &sh.tip[4 * stips[en.roff + lsite[s]]]; x2[0] = tv[0]; x2[1] = tv[1]; x2[2] = tv[2]; x2[3] = tv[3]; }
    qa[s][0] = x1[0]; qa[s][1] = dq1 * x1[1]; qa[s][2] = dq2 * x1[2]; qa[s][3] = dq3 * x1[3];
    ra[s][0] = x2[0]; ra[s][1] = dr1 * x2[1]; ra[s][2] = dr2 * x2[2]; ra[s][3] = dr3 * x2[3];
  }
  double ab[NS][4];
  #pragma unroll
  for (int l = 0; l < 4; l++) {                       // (B) one row of EI for all sites
    const double e1 = EI_(l * 4 + 1), e2 = EI_(l * 4 + 2), e3 = EI_(l * 4 + 3);
    #pragma unroll
    for (int s = 0; s < NS; s++) {
      const double al = fma(e3, qa[s][3], fma(e2, qa[s][2], fma(e1, qa[s][1], qa[s][0])));
      const double ar = fma(e3, ra[s][3], fma(e2, ra[s][2], fma(e1, ra[s][1], ra[s][0])));
      ab[s][l] = al * ar;
    }
  }
  double v[NS][4];
  #pragma unroll
  for (int m = 0; m < 4; m++) {                       // (C) one column of EV for all sites
    const double f0 = EV_(0 * 4 + m), f1 = EV_(1 * 4 + m), f2 = EV_(2 * 4 + m), f3 = EV_(3 * 4 + m);
    #pragma unroll
    for (int s = 0; s < NS; s++)
      v[s][m] = fma(ab[s][3], f3, fma(ab[s][2], f2, fma(ab[s][1], f1, ab[s][0] * f0)));
  }
  unsigned int cnt = 0;
  double *const dst = a.clv + (store ? en.poff : 0);     // (a scoring entry's poff is its candidate index)
  #pragma unroll
  for (int s = 0; s < NS; s++) {
    if (RS != 0) {                           // TIP_TIP is never rescaled (avxLikelihood.c:87-160)
      const unsigned int b = __ballot_sync(0xffffffffu, all_tiny(v[s]));
      if (((b >> (lane & ~3)) & 0xFu) == 0xFu) {            // all 16 values of the site are tiny
        #pragma unroll
        for (int m = 0; m < 4; m++) v[s][m] *= kTwoToThe256;
        if (c == 0 && valid[s]) cnt += (unsigned int)a.wgt[eoff[s] >> 4];
      }
    }
    if (valid[s] && store) st256(dst + eoff[s], v[s]);
    carry[s][0] = v[s][0]; carry[s][1] = v[s][1]; carry[s][2] = v[s][2]; carry[s][3] = v[s][3];
  }
  return cnt;
#undef EI_
#undef EV_
}

// ------------------------------------------------------------------------------------------------------
// d-table: exp(EIGN[k] * rate[c] * lz) for every (entry, partition of the class, side, category, k).
// Same association as makeP / calcDiagptable: rate * (EIGN * lz).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dtab_kernel(const DtabArgs a)
{
  const size_t total = (size_t)a.count * a.nparts_cls * a.row;
  const int S = a.states;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % a.row);
    const size_t ep = i / a.row;
    const int pc = (int)(ep % a.nparts_cls), e = (int)(ep / a.nparts_cls);
    const int c = w / (2 * S), side = (w / S) & 1, k = w % S;
    const double *model = a.models + (size_t)pc * a.model_doubles;   // EIGN[S] rates[4] ...
    const int slot = a.nbranches > 1 ? a.cls_part_global[pc] : 0;
    const double lz = side ? a.lzr[(size_t)e * a.nbranches + slot] : a.lzq[(size_t)e * a.nbranches + slot];
    const double rate = a.cat_rates ? a.cat_rates[(size_t)pc * a.max_cats + c] : model[a.rates_off + c];
    const double ek = model[(a.cat_rates ? 0 : c) * a.eign_cat_stride + k];
    const bool lg4 = a.lg4_off >= 0 && model[a.lg4_off] != 0.0;
    a.dtab[i] = !k ? 1.0 : (lg4 ? exp(rate * ek * lz)          // makeP_FlexLG4 newviewGenericSpecial.c:189-190
                                : exp(rate * (ek * lz)));       // makeP :95-112
  }
}

void launch_dtab(const DtabArgs &a, cudaStream_t s)
{
  const size_t total = (size_t)a.count * a.nparts_cls * a.row;
  if (total == 0) return;
  int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
  dtab_kernel<<<blocks, 256, 0, s>>>(a);
}

// ------------------------------------------------------------------------------------------------------
// Fused traversal: newviews [first, count) for this tile, then (optionally) evaluate or the sum table
// at the branch described by entry 0.   NS = sites per thread (tile chunk = 64*NS sites).
//
// Prologue per chunk: the descriptor is expanded into shared memory (element offsets instead of node
// numbers) and the tip-code rows the descriptor touches are brought in with one cp.async.bulk
// (TMA, 1-D) per row, all completing on a single mbarrier -- the only cold HBM reads of the kernel
// are issued up front and in bulk instead of one dependent byte load per entry.
// ------------------------------------------------------------------------------------------------------
// Window of the DNA GAMMA kernel: 48 entries keep (entries + tip rows + d-table rows + the prefetch ring) of a
// 256-site chunk at ~102 KB, i.e. two CTAs per SM.
constexpr int kDnaWindow = 48;

// Which child of an entry goes through the prefetch ring: 1 = q, 2 = r, 0 = none (tips and carry need no load)
__device__ __forceinline__ int ring_child(const SEntry &en)
{
  const bool q_load = en.tipCase == 2 && !(en.flags & kFlagCarryA), r_load = en.tipCase != 0 && !(en.flags & kFlagCarryB);
  return q_load ? 1 : (r_load ? 2 : 0);
}

// SMALL: the call (descriptor, mask, d-table rows, finish arguments) arrives in the kernel parameters and the last CTA
// of a `fold` launch finishes the call (engine.cuh SmallCall) -- the per-call path of a tree search on a small alignment.
template <int NS, bool PARAM, int MINB, bool SMALL = false>
__global__ void __launch_bounds__(kThreads, MINB)
traverse_dna_kernel(const TraverseArgs a, const DnaConsts kc, const typename std::conditional<SMALL, SmallCall, NoSmallCall>::type sc)
{
  constexpr int TS = 64 * NS;                  // sites per chunk = bytes per staged tip row
  constexpr int KW = kDnaWindow;
  __shared__ DnaShared sh;
  __shared__ __align__(8) uint64_t mbar;
  extern __shared__ __align__(16) unsigned char dyn[];
  const SEntry *ents = reinterpret_cast<const SEntry *>(dyn);                  // window entries; slot KW = branch entry
  const uint8_t *stips = dyn + (size_t)(KW + 1) * sizeof(SEntry);               // [2 KW + 2][TS] staged tip rows
  const double *sdtab = reinterpret_cast<const double *>(dyn + window_dtab_offset<TS, KW>());     // [KW + 1][32] d-table rows
  double *ring = reinterpret_cast<double *>(dyn + window_smem_bytes<TS, KW>(kDnaDtabRow));          // [2][NS][256 threads][4]

  __shared__ DevEntry s_ent[SMALL ? kSmallEntries : 1];
  __shared__ unsigned char s_mask[SMALL ? kSmallParts : 1];

  const Tile tile = a.tiles[blockIdx.x];
  if constexpr (!SMALL) { if (!a.mask[tile.part]) return; }      // (SMALL: masked partitions are not launched at all)

  const int tid = threadIdx.x, lane = tid & 31;
  const int c = tid & 3;                       // Gamma category owned by this thread
  const size_t N = a.N;
  const DevEntry *entries = a.entries;
  if constexpr (SMALL) {
    if (tid < sc.count) s_ent[tid] = sc.entries[tid];
    if (tid < kSmallParts) s_mask[tid] = sc.mask[tid];
    entries = s_ent;
  }
  {
    const double *model = a.models + (size_t)tile.part_cls * kDnaModelDoubles;   // EIGN[4] rates[4] EI[16] EV[16] tip[64]
    if (tid < 64) sh.tip[tid] = model[40 + tid];
    if (!PARAM && tid < 16) { sh.EI[tid] = model[8 + tid]; sh.EV[tid] = model[24 + tid]; }
    if (tid == 0) mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }

  const double *drow0 = a.dtab + (size_t)tile.part_cls * kDnaDtabRow;       // row of entry e: + e * dstride
  const size_t dstride = (size_t)a.nparts_cls * kDnaDtabRow;
  double *const myring = ring + (size_t)tid * 4;                            // + (buffer * NS + s) * 256 * 4
  double acc = 0.0;
  uint32_t phase = 0;

  for (unsigned chunk0 = 0; chunk0 < tile.nsites; chunk0 += TS) {
    const unsigned nchunk = min(tile.nsites - chunk0, (unsigned)TS);
    size_t eoff[NS];                           // element offset of (site, category) inside a CLV; site = eoff >> 4
    bool valid[NS];
    int lsite[NS];
    #pragma unroll
    for (int s = 0; s < NS; s++) {
      const unsigned i = (unsigned)s * 64 + (tid >> 2);
      valid[s] = i < nchunk;
      lsite[s] = valid[s] ? (int)i : 0;
      eoff[s] = ((size_t)tile.site0 + chunk0 + lsite[s]) * 16 + c * 4;
    }
    double carry[NS][4];                       // CLV of the entry just computed (kFlagCarry* say when it is an operand)
    #pragma unroll
    for (int s = 0; s < NS; s++) { carry[s][0] = carry[s][1] = carry[s][2] = carry[s][3] = 0.0; }

    // ---------------- newview entries, one window at a time ----------------
    int w0 = a.first;
    do {
    const int w1 = min(w0 + KW, a.count);
    phase = stage_window<TS, KW>(dyn, &mbar, phase, entries, a.tips, a.mxtips, N, 16, (size_t)tile.site0 + chunk0, nchunk, w0, w1,
                                 a.first, a.count, w0 == a.first && a.mode != MODE_NEWVIEW, SMALL ? nullptr : drow0, dstride, kDnaDtabRow,
                                 SevView{a.sev_table, a.sev_gapcol, a.clv, a.sev_ntiles, a.sev_tile0 + (int)blockIdx.x, tile.part_cls,
                                         a.nparts_cls, 16, tile.site0}, a.scratch_off);
    if constexpr (SMALL) {
      // the d-table rows came with the launch: this partition's row of every window entry (slot e - w0) and of the branch
      // entry (slot KW) from the parameters into the window's table
      double *sd = const_cast<double *>(sdtab);
      const int nrows = (w1 - w0) + ((w0 == a.first && a.mode != MODE_NEWVIEW) ? 1 : 0);
      for (int i = tid; i < nrows * kDnaDtabRow; i += kThreads) {
        const int j = i / kDnaDtabRow, w = i % kDnaDtabRow;
        const bool is_branch = j == (w1 - w0);
        sd[(size_t)(is_branch ? KW : j) * kDnaDtabRow + w] = sc.dtab[((size_t)(is_branch ? 0 : w0 + j) * a.nparts_cls + tile.part_cls) * kDnaDtabRow + w];
      }
      __syncthreads();
    }
    for (int e = w0; e < w1; e++) {
      const SEntry en = ents[e - w0];
      // The one inner child that is neither a tip nor in the carry registers comes through the prefetch ring: its
      // cp.async was issued while the PREVIOUS entry was being computed (the window's first entry issues its own).
      // (-S: an all-gap child is the node's gap column -- the same 32 bytes for every site of the thread.)
      const int mine = (en.flags & kFlagGapP) ? 0 : ring_child(en);
      if (e == w0 && mine) {
        const double *src = a.clv + (mine == 1 ? en.qoff : en.roff);
        const bool gap = en.flags & (mine == 1 ? kFlagGapA : kFlagGapB);
        #pragma unroll
        for (int s = 0; s < NS; s++) async_copy32(myring + (size_t)((e & 1) * NS + s) * kThreads * 4, src + (gap ? (size_t)(c * 4) : eoff[s]));
      }
      async_commit();
      if (e + 1 < w1) {
        const SEntry nx = ents[e + 1 - w0];
        const int next = (nx.flags & kFlagGapP) ? 0 : ring_child(nx);
        if (next) {
          const double *src = a.clv + (next == 1 ? nx.qoff : nx.roff);
          const bool gap = nx.flags & (next == 1 ? kFlagGapA : kFlagGapB);
          #pragma unroll
          for (int s = 0; s < NS; s++) async_copy32(myring + (size_t)(((e + 1) & 1) * NS + s) * kThreads * 4, src + (gap ? (size_t)(c * 4) : eoff[s]));
        }
      }
      async_commit();
      async_wait_all_but_last();               // this entry's operand has landed; the next entry's stays in flight
      if (en.flags & kFlagGapP) continue;      // -S: all sites of the tile are gaps below this node: nothing to compute or store

      const double *sd = sdtab + (size_t)(e - w0) * kDnaDtabRow + c * 8;
      const double *rg = myring + (size_t)((e & 1) * NS) * kThreads * 4;
      const bool q_inner = en.tipCase == 2, r_inner = en.tipCase != 0;
      // sources: 0 tip, 1 carry, 2 ring, 3 direct load (r, only when q already uses the ring)
      const int qs = !q_inner ? 0 : ((en.flags & kFlagCarryA) ? 1 : 2);
      const int rs = !r_inner ? 0 : ((en.flags & kFlagCarryB) ? 1 : (mine == 2 ? 2 : 3));
      unsigned int cnt;
      const bool score = en.flags & kFlagScore;
#define DNA_ENTRY(QS, RS) dna_entry<QS, RS, NS, PARAM>(carry, a, kc, sh, stips, en, sd, rg, eoff, valid, lsite, c, lane, !score)
      switch (qs * 4 + rs) {                   // block-uniform
      case 0:  cnt = DNA_ENTRY(0, 0); break;   // TIP_TIP
      case 1:  cnt = DNA_ENTRY(0, 1); break;   // TIP_INNER, inner child = previous entry
      case 2:  cnt = DNA_ENTRY(0, 2); break;   // TIP_INNER, inner child loaded
      case 6:  cnt = DNA_ENTRY(1, 2); break;   // INNER_INNER, q = previous entry
      case 9:  cnt = DNA_ENTRY(2, 1); break;   // INNER_INNER, r = previous entry
      default: cnt = DNA_ENTRY(2, 3); break;   // INNER_INNER, both loaded
      }
#undef DNA_ENTRY
      if (score) {
        // exb_score_insertions: the CLV just formed (carry) is the candidate's new node; contract it with the pruned
        // subtree's CLV under the diagonal table of the branch p--s (evaluateGTRGAMMA evaluateGenericSpecial.c:1879) and
        // leave this warp's share in its own slot of score_out -- summed over chunks here, over warps and tiles by
        // score_finish_kernel in a fixed order
        const SEntry br = ents[KW];
        const double *diag = sdtab + (size_t)KW * kDnaDtabRow + c * 8;
        double part = 0.0;
        #pragma unroll
        for (int s = 0; s < NS; s++) {
          double xs[4];
          if (br.flags & kFlagPTip) { const double *tv = &sh.tip[4 * stips[br.poff + lsite[s]]]; xs[0] = tv[0]; xs[1] = tv[1]; xs[2] = tv[2]; xs[3] = tv[3]; }
          else ld256(a.clv + br.poff + eoff[s], xs);
          double term = carry[s][0] * xs[0];
          #pragma unroll
          for (int k = 1; k < 4; k++) term += carry[s][k] * xs[k] * diag[k];
          term += __shfl_xor_sync(0xffffffffu, term, 1);
          term += __shfl_xor_sync(0xffffffffu, term, 2);
          if (c == 0 && valid[s]) part += (double)a.wgt[eoff[s] >> 4] * log(0.25 * fabs(term));
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
        const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) {
          double2 *slot = a.score_out + (size_t)en.poff * a.score_stride + (size_t)(a.sev_tile0 + (int)blockIdx.x) * (kThreads / 32) + (tid >> 5);
          double2 v = make_double2(part, (double)tot);
          if (chunk0) { const double2 old = *slot; v.x += old.x; v.y += old.y; }
          *slot = v;
        }
        continue;
      }
      if (r_inner && __any_sync(0xffffffffu, cnt != 0)) {
        const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
      }
    }
    } while ((w0 += KW) < a.count);

    if (a.mode == MODE_NEWVIEW || a.mode == MODE_SCORE) continue;

    // ---------------- branch (entry 0): evaluate or sum table ----------------
    const SEntry br = ents[KW];
    const bool ptip = br.flags & kFlagPTip, qtip = br.flags & kFlagQTip;
    const bool p_carry = br.flags & kFlagCarryA, q_carry = br.flags & kFlagCarryB;
    const double *diag = sdtab + (size_t)KW * kDnaDtabRow + c * 8;     // row 0, q side: calcDiagptable (diag[0] = 1)

    #pragma unroll
    for (int s = 0; s < NS; s++) {
      double xa[4], xb[4];
      if (p_carry) { xa[0] = carry[s][0]; xa[1] = carry[s][1]; xa[2] = carry[s][2]; xa[3] = carry[s][3]; }
      else if (!ptip) ld256(a.clv + br.poff + ((br.flags & kFlagGapA) ? (size_t)(c * 4) : eoff[s]), xa);
      else { const double *tv = &sh.tip[4 * stips[br.poff + lsite[s]]]; xa[0] = tv[0]; xa[1] = tv[1]; xa[2] = tv[2]; xa[3] = tv[3]; }
      if (q_carry) { xb[0] = carry[s][0]; xb[1] = carry[s][1]; xb[2] = carry[s][2]; xb[3] = carry[s][3]; }
      else if (!qtip) ld256(a.clv + br.qoff + ((br.flags & kFlagGapB) ? (size_t)(c * 4) : eoff[s]), xb);
      else { const double *tv = &sh.tip[4 * stips[br.qoff + lsite[s]]]; xb[0] = tv[0]; xb[1] = tv[1]; xb[2] = tv[2]; xb[3] = tv[3]; }
      if (a.mode == MODE_SUM) {
        double sm[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) sm[k] = xa[k] * xb[k];
        if (valid[s]) st256(a.sum + eoff[s], sm);
      } else {
        double term = xa[0] * xb[0];           // diag[0] = 1
        #pragma unroll
        for (int k = 1; k < 4; k++) term += xa[k] * xb[k] * diag[k];
        term += __shfl_xor_sync(0xffffffffu, term, 1);
        term += __shfl_xor_sync(0xffffffffu, term, 2);
        if (c == 0 && valid[s]) acc += (double)a.wgt[eoff[s] >> 4] * log(0.25 * fabs(term));
      }
    }
  }

  if (a.mode == MODE_EVALUATE) {
    const double tot = block_sum(acc, sh.red);
    if (tid == 0) { a.tile_out[2 * (size_t)blockIdx.x] = tot; a.tile_out[2 * (size_t)blockIdx.x + 1] = 0.0; }
  }
  if constexpr (SMALL) {
    if (sc.fold) {
      // the last CTA of the call's last launch does what finish_kernel would do in a launch of its own
      __shared__ int last, peer_lost;
      __shared__ double red0[4], red1[4];
      __threadfence();                             // CLVs, sum table, tile partials, scaler increments: visible device-wide
      __syncthreads();
      if (tid == 0) last = atomicAdd(sc.fin.ticket, 1u) == gridDim.x - 1;
      __syncthreads();
      if (last) {
        finish_body(sc.fin, s_ent, s_mask, reinterpret_cast<unsigned int *>(dyn), red0, red1, 0, 1);
        __threadfence_system();
        finish_publish(sc.fin, &peer_lost);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------
// CAT / PSR traversal (newviewGTRCAT_AVX avxLikelihood.c:326-485, evaluateGTRCAT evaluateGenericSpecial.c:1988-2083,
// sumCAT makenewzGenericSpecial.c:1650-1796): one rate category per site, a CLV is 4 doubles = 32 B per site.
// Same structure as the GAMMA kernel -- descriptor loop inside, carry, staged tips -- with thread = site; a site
// selects its exponentials through rate_cat[site].  Scaling is decided per thread (all 4 values tiny).
// ------------------------------------------------------------------------------------------------------
template <bool PARAM>
__global__ void __launch_bounds__(kThreads, 3)
traverse_dna_cat_kernel(const TraverseArgs a, const DnaConsts kc)
{
  constexpr int TS = 256;                      // sites per chunk = threads
  __shared__ DnaShared sh;
  __shared__ __align__(8) uint64_t mbar;
  extern __shared__ __align__(16) unsigned char dyn[];
  const SEntry *ents = reinterpret_cast<const SEntry *>(dyn);                  // window entries; slot kWindow = branch entry
  const uint8_t *stips = dyn + (size_t)(kWindow + 1) * sizeof(SEntry);          // [2*kWindow + 2][TS] staged tip rows

  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  const int tid = threadIdx.x, lane = tid & 31;
  const size_t N = a.N;
  {
    const double *model = a.models + (size_t)tile.part_cls * kDnaModelDoubles;
    if (tid < 64) sh.tip[tid] = model[40 + tid];
    if (!PARAM && tid < 16) { sh.EI[tid] = model[8 + tid]; sh.EV[tid] = model[24 + tid]; }
    if (tid == 0) mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }
#define EI_(i) (PARAM ? kc.EI[i] : sh.EI[i])
#define EV_(i) (PARAM ? kc.EV[i] : sh.EV[i])
  const size_t dstride = (size_t)a.nparts_cls * a.dtab_row;
  double acc = 0.0;
  uint32_t phase = 0;

  for (unsigned chunk0 = 0; chunk0 < tile.nsites; chunk0 += TS) {
    const unsigned nchunk = min(tile.nsites - chunk0, (unsigned)TS);
    const bool valid = (unsigned)tid < nchunk;
    const int lsite = valid ? tid : 0;
    const size_t site = (size_t)tile.site0 + chunk0 + lsite;
    const size_t eoff = site * 4;
    const int cat = a.rate_cat[site];
    const double *drow = a.dtab + ((size_t)tile.part_cls * a.max_cats + cat) * 8;    // + e * dstride; q side +0, r side +4
    double carry[4] = {0.0, 0.0, 0.0, 0.0};

    int w0 = a.first;
    do {
    const int w1 = min(w0 + kWindow, a.count);
    phase = stage_window<TS>(dyn, &mbar, phase, a.entries, a.tips, a.mxtips, N, 4, (size_t)tile.site0 + chunk0, nchunk, w0, w1,
                     a.first, a.count, w0 == a.first && a.mode != MODE_NEWVIEW);
    for (int e = w0; e < w1; e++) {
      const SEntry en = ents[e - w0];
      double dq[4], dr[4];
      ld256_ro(drow + (size_t)e * dstride, dq);
      ld256_ro(drow + (size_t)e * dstride + 4, dr);
      const bool q_inner = en.tipCase == 2, r_inner = en.tipCase != 0;
      double x1[4], x2[4];
      if (en.flags & kFlagCarryA) { x1[0] = carry[0]; x1[1] = carry[1]; x1[2] = carry[2]; x1[3] = carry[3]; }
      else if (q_inner) ld256(a.clv + en.qoff + eoff, x1);
      else { const double *tv = &sh.tip[4 * stips[en.qoff + lsite]]; x1[0] = tv[0]; x1[1] = tv[1]; x1[2] = tv[2]; x1[3] = tv[3]; }
      if (en.flags & kFlagCarryB) { x2[0] = carry[0]; x2[1] = carry[1]; x2[2] = carry[2]; x2[3] = carry[3]; }
      else if (r_inner) ld256(a.clv + en.roff + eoff, x2);
      else { const double *tv = &sh.tip[4 * stips[en.roff + lsite]]; x2[0] = tv[0]; x2[1] = tv[1]; x2[2] = tv[2]; x2[3] = tv[3]; }

      const double tq1 = dq[1] * x1[1], tq2 = dq[2] * x1[2], tq3 = dq[3] * x1[3];
      const double tr1 = dr[1] * x2[1], tr2 = dr[2] * x2[2], tr3 = dr[3] * x2[3];
      double v[4] = {0.0, 0.0, 0.0, 0.0};
      #pragma unroll
      for (int l = 0; l < 4; l++) {
        const double al = x1[0] + EI_(l * 4 + 1) * tq1 + EI_(l * 4 + 2) * tq2 + EI_(l * 4 + 3) * tq3;
        const double ar = x2[0] + EI_(l * 4 + 1) * tr1 + EI_(l * 4 + 2) * tr2 + EI_(l * 4 + 3) * tr3;
        const double ab = al * ar;
        #pragma unroll
        for (int m = 0; m < 4; m++) v[m] += ab * EV_(l * 4 + m);
      }
      unsigned int cnt = 0;
      if (r_inner) {
        const bool small = fabs(v[0]) < kMinLikelihood && fabs(v[1]) < kMinLikelihood &&
                           fabs(v[2]) < kMinLikelihood && fabs(v[3]) < kMinLikelihood;
        if (small) {
          #pragma unroll
          for (int m = 0; m < 4; m++) v[m] *= kTwoToThe256;
          if (valid) cnt = (unsigned int)a.wgt[site];
        }
        if (__any_sync(0xffffffffu, cnt != 0)) {
          const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
          if (lane == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
        }
      }
      if (valid) st256(a.clv + en.poff + eoff, v);
      carry[0] = v[0]; carry[1] = v[1]; carry[2] = v[2]; carry[3] = v[3];
    }
    } while ((w0 += kWindow) < a.count);
    if (a.mode == MODE_NEWVIEW) continue;

    const SEntry br = ents[kWindow];
    const bool ptip = br.flags & kFlagPTip, qtip = br.flags & kFlagQTip;
    double xa[4], xb[4];
    if (br.flags & kFlagCarryA) { xa[0] = carry[0]; xa[1] = carry[1]; xa[2] = carry[2]; xa[3] = carry[3]; }
    else if (!ptip) ld256(a.clv + br.poff + eoff, xa);
    else { const double *tv = &sh.tip[4 * stips[br.poff + lsite]]; xa[0] = tv[0]; xa[1] = tv[1]; xa[2] = tv[2]; xa[3] = tv[3]; }
    if (br.flags & kFlagCarryB) { xb[0] = carry[0]; xb[1] = carry[1]; xb[2] = carry[2]; xb[3] = carry[3]; }
    else if (!qtip) ld256(a.clv + br.qoff + eoff, xb);
    else { const double *tv = &sh.tip[4 * stips[br.qoff + lsite]]; xb[0] = tv[0]; xb[1] = tv[1]; xb[2] = tv[2]; xb[3] = tv[3]; }
    if (a.mode == MODE_SUM) {
      double sm[4];
      #pragma unroll
      for (int k = 0; k < 4; k++) sm[k] = xa[k] * xb[k];
      if (valid) st256(a.sum + eoff, sm);
    } else {
      double diag[4];
      ld256_ro(drow, diag);                    // row 0, q side
      double term = xa[0] * xb[0];
      #pragma unroll
      for (int k = 1; k < 4; k++) term += xa[k] * xb[k] * diag[k];
      if (valid) acc += (double)a.wgt[site] * log(fabs(term));     // no 1/4: a single category (evaluateGenericSpecial.c:289)
    }
  }
#undef EI_
#undef EV_
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
  __shared__ union { CoreSmem g; CoreCatSmem c; } sh;
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  double t1, t2;
  const double *model = a.models + (size_t)tile.part_cls * kDnaModelDoubles;
  const double lz = a.lz[a.nbranches > 1 ? tile.part : 0];
  if (a.rate_cat) core_tile_cat<4>(tile, a.sum, a.wgt, model, a.rate_cat, a.cat_rates + (size_t)tile.part_cls * a.max_cats, a.max_cats, lz, sh.c, t1, t2);
  else            core_tile<4>(tile, a.sum, a.wgt, model, lz, sh.g, t1, t2);
  if (threadIdx.x == 0) { a.tile_out[2 * (size_t)blockIdx.x] = t1; a.tile_out[2 * (size_t)blockIdx.x + 1] = t2; }
}

size_t traverse_dna_smem(int ns)
{
  // entries + tip rows + d-table rows of a window, then the prefetch ring [2][ns][256 threads] x 32 B
  return (size_t)(kDnaWindow + 1) * sizeof(SEntry) + (size_t)(2 * kDnaWindow + 2) * 64 * ns +
         (size_t)(kDnaWindow + 1) * kDnaDtabRow * sizeof(double) + (size_t)2 * ns * kThreads * 32;
}

template <int NS, bool PARAM, int MINB>
static void launch_variant(const TraverseArgs &a, const DnaConsts &kc, int ntiles, cudaStream_t s, const SmallCall *small = nullptr)
{
  const size_t smem = traverse_dna_smem(NS);
  if (small) {
    if constexpr (PARAM) {
      static size_t configured[kMaxDevices];
      ensure_dynamic_smem(traverse_dna_kernel<NS, true, MINB, true>, std::min<size_t>(227 * 1024 - 2048, std::max<size_t>(smem, 64 * 1024)), configured);
      traverse_dna_kernel<NS, true, MINB, true><<<ntiles, kThreads, smem, s>>>(a, kc, *small);
    }
    return;
  }
  static size_t configured[kMaxDevices];
  ensure_dynamic_smem(traverse_dna_kernel<NS, PARAM, MINB>, std::min<size_t>(227 * 1024 - 2048, std::max<size_t>(smem, 64 * 1024)), configured);
  traverse_dna_kernel<NS, PARAM, MINB><<<ntiles, kThreads, smem, s>>>(a, kc, NoSmallCall{});
}

// NS sites per thread: 4 for full tiles (128 registers, 2 CTAs per SM), fewer for small tiles (3 CTAs per SM).
// kc: EI / EV of the launch's single partition as kernel parameters (constant bank -> uniform registers);
// null: many partitions in one launch, every CTA reads its partition's EI / EV from shared memory.
void launch_traverse_dna(const TraverseArgs &a, const DnaConsts *kc, int ntiles, cudaStream_t s, const SmallCall *small)
{
  if (ntiles <= 0) return;
  static int ns_max = -1;                        // development knob: EXB_DNA_NS = 1 | 2 | 4 sites per thread at most
  if (ns_max < 0) { const char *v = getenv("EXB_DNA_NS"); ns_max = v ? atoi(v) : 4; }
  const int ns = (a.tile_sites <= 64 || ns_max == 1) ? 1 : ((a.tile_sites <= 128 || ns_max == 2) ? 2 : 4);
  if (!kc) {
    DnaConsts none;
    memset(&none, 0, sizeof none);
    if (ns == 1)      launch_variant<1, false, 3>(a, none, ntiles, s);
    else if (ns == 2) launch_variant<2, false, 3>(a, none, ntiles, s);
    else              launch_variant<4, false, 2>(a, none, ntiles, s);
    return;
  }
  if (ns == 1)      launch_variant<1, true, 3>(a, *kc, ntiles, s, small);
  else if (ns == 2) launch_variant<2, true, 3>(a, *kc, ntiles, s, small);
  else              launch_variant<4, true, 2>(a, *kc, ntiles, s, small);
}
void launch_traverse_dna_cat(const TraverseArgs &a, const DnaConsts *kc, int ntiles, cudaStream_t s)
{
  if (ntiles <= 0) return;
  const size_t smem = window_smem_bytes<256>();
  static size_t configured[2][kMaxDevices];
  const size_t want = std::min<size_t>(227 * 1024 - 2048, std::max<size_t>(smem, 64 * 1024));
  if (kc) ensure_dynamic_smem(traverse_dna_cat_kernel<true>, want, configured[1]);
  else    ensure_dynamic_smem(traverse_dna_cat_kernel<false>, want, configured[0]);
  if (kc) traverse_dna_cat_kernel<true><<<ntiles, kThreads, smem, s>>>(a, *kc);
  else {
    DnaConsts none;
    memset(&none, 0, sizeof none);
    traverse_dna_cat_kernel<false><<<ntiles, kThreads, smem, s>>>(a, none);
  }
}
void launch_core_dna(const CoreArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles > 0) core_dna_kernel<<<ntiles, kThreads, 0, s>>>(a);
}

}  // namespace exb
