// kernels_aa.cu -- 20-state (protein) GAMMA likelihood kernels for sm_100a.
//
// Replaces, for AA partitions, the per-site loops of
//   newviewGTRGAMMAPROT_AVX  avxLikelihood.c:1312-1800
//   evaluateGTRGAMMAPROT     evaluateGenericSpecial.c:1393-1461
//   sumGAMMAPROT             makenewzGenericSpecial.c:2083-2153
//   coreGTRGAMMAPROT         makenewzGenericSpecial.c:2581-2655
//
// Same structure as kernels_dna.cu: the traversal loop is inside the kernel, a CTA owns a tile of sites for the whole
// descriptor, the CLV just produced stays in registers for the parent ("carry"), every CLV is written to HBM once.
// The arithmetic runs on the FP64 tensor pipe: a warp owns 8 sites (GAMMA: all four categories of each; PSR: four
// groups of 8 sites), and the three 20x20 contractions of an entry
//     Y1 = (d_q (.) X1) * EI^T        Y2 = (d_r (.) X2) * EI^T        X3 = (Y1 (.) Y2) * EV
// are [8 x 20] x [20 x 24] GEMMs on mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  All arithmetic is FP64 -- the
// reference's precision; tcgen05 has no f64 kind.  P(t) is never materialised: the exponentials of
// the per-call d-table are folded into the A operand.  EI^T / EV sit in shared memory as ready-made B fragments (one
// set per eigen-system: LG4M / LG4X have four), tip children come from a per-call look-up table (aa_lut_kernel).
#include "engine.cuh"
#include "core_tile.cuh"
#include "stage.cuh"
#include <algorithm>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

namespace exb {

namespace {

constexpr int S = 20;
constexpr int SPAN = 80;
constexpr int kAaThreads = 32 * kAaWarps;  // 12 warps, one CTA per SM (the register budget of the DMMA pipeline)
static_assert(kAaThreads == 32 * kAaWarps, "score_out slots per tile");

struct AaShared { double red[32]; };

// per-category constants in dynamic shared memory, behind the window area (45.4 KB):
struct AaConstsSmem {
  double fragEI[4][15][32];  // B fragments of EI_c^T, [category][k-step * 3 + n-tile][lane]: conflict-free LDS.64
  double fragEV[4][15][32];  // B fragments of EV_c
  double tip[4][460];        // tipVector_c[23 codes][20]
};

__device__ __forceinline__ void ld256(const double *p, double *v)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void ld256_ro(const double *p, double *v)
{
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
__device__ __forceinline__ double2 ld128(const double *p)
{
  double2 v;
  asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st256(double *p, const double *v)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
enum Src : int { SRC_TIP = 0, SRC_CARRY = 1, SRC_LOAD = 2 };

// fabs(v) < 2^-256 on the integer pipe (the FP64 pipe belongs to the DMMAs): the biased exponent of 2^-256 is 0x2FF
__device__ __forceinline__ unsigned int abs_hi(double v) { return (unsigned int)__double2hiint(v) & 0x7fffffffu; }
constexpr unsigned int kMinLikelihoodHi = 0x2FF00000u;

// D(8x8) += A(8x4) * B(4x8), FP64 tensor pipe (SASS: DMMA.8x8x4).
//   A: thread holds A[lane/4][lane%4]; B: thread holds B[lane%4][lane/4]; C/D: C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One traversal entry for the 8 sites of this warp, all four categories.  A child is
//   MEM   read from memory (global CLV with category stride 20, or the shared-memory tip vector with stride 0)
//         and contracted with EI^T on the tensor pipe,
//   CARRY taken from the carry registers (C-fragment layout -> A operand by warp shuffles), same contraction,
//   LUT   a tip whose product P(t) * tipVector[code] was tabulated for this (entry, side) by aa_lut_kernel:
//         its Y comes straight from the table in C-fragment layout, no DMMA at all.
// The combinations that occur are compiled separately so that the k-step loop is branch-free and the
// loads of the next category can be scheduled under the DMMAs.
enum ChildMode : int { CH_MEM = 0, CH_CARRY = 1, CH_LUT = 2 };

// Which of the 20 contraction indices does lane tq feed into k-step s?  The sum over k is order-free, so A and B
// fragments may use any common assignment.  The natural one (k = 4s + tq) would force a C-fragment -- which holds
// columns 8t + 2tq + {0,1} -- to be re-distributed with two shuffles and a select per k-step (measured: SHFL + FSEL
// took more issue slots than the DMMAs).  Instead k-steps 0..3 take the columns the lane already holds (t = s/2,
// element s&1), and only k-step 4 moves data: lanes 2,3 fetch columns 17 / 19 from lanes 0,1 -- one shuffle per
// contraction instead of ten (36.7 -> 32.0 ms per C4 traversal).
__device__ __forceinline__ int kappa(int s, int tq)
{
  return s < 4 ? 8 * (s >> 1) + 2 * tq + (s & 1) : (tq < 2 ? 16 + 2 * tq : 13 + 2 * tq);
}

// A operand of k-step s taken from a value in C-fragment layout (v[t][e] = column 8t + 2tq + e of my row)
__device__ __forceinline__ double c_to_a(const double (&v)[3][2], int s, int lane, int tq)
{
  if (s < 4) return v[s >> 1][s & 1];
  const double moved = __shfl_sync(0xffffffffu, v[2][1], (lane & ~3) | (tq & 1));
  return tq < 2 ? v[2][0] : moved;
}

template <int MODE>
__device__ __forceinline__ void child_y(double (&y)[3][2], const double (&carry)[4][3][2], const double (*fEI)[32],
                                        const double *x, const double *lutrow, const double *dside,
                                        int c, int lane, int tq)
{
  if (MODE == CH_LUT) {
    const double *r = lutrow + 2 * tq;
    #pragma unroll
    for (int t = 0; t < 3; t++) {
      double2 v = make_double2(0.0, 0.0);
      if (t < 2 || tq < 2) v = __ldg(reinterpret_cast<const double2 *>(r + 8 * t));
      y[t][0] = v.x; y[t][1] = v.y;
    }
    return;
  }
  double av[5];
  #pragma unroll
  for (int s = 0; s < 5; s++) {
    const int k = kappa(s, tq);
    if (MODE == CH_CARRY) av[s] = c_to_a(carry[c], s, lane, tq);
    else av[s] = x[k];
    av[s] *= __ldg(dside + k);
  }
  #pragma unroll
  for (int t = 0; t < 3; t++) y[t][0] = y[t][1] = 0.0;
  #pragma unroll
  for (int s = 0; s < 5; s++)
    #pragma unroll
    for (int t = 0; t < 3; t++) dmma(y[t][0], y[t][1], av[s], fEI[s * 3 + t][lane]);
}

// Per-slot operands of one entry.  A "slot" is a Gamma category of the warp's 8 sites (GAMMA) or one of the
// warp's four 8-site groups (CAT / PSR, one category per site).
struct AaSlotsCat {                // CAT: everything differs per slot (four different sites)
  const double *xq_[4], *xr_[4];   // child vectors (memory mode): 20 doubles of my row
  const double *lq_[4], *lr_[4];   // tip table rows (LUT mode): 20 doubles
  const double *dq_[4];            // exponentials of the q side; the r side follows at +20
  __device__ __forceinline__ const double *xq(int c) const { return xq_[c]; }
  __device__ __forceinline__ const double *xr(int c) const { return xr_[c]; }
  __device__ __forceinline__ const double *lq(int c) const { return lq_[c]; }
  __device__ __forceinline__ const double *lr(int c) const { return lr_[c]; }
  __device__ __forceinline__ const double *dq(int c) const { return dq_[c]; }
};
struct AaSlotsGamma {              // GAMMA: one site, slot = category: a base and a constant stride
  const double *x1, *x2, *lut1, *lut2, *drow;
  int st1, st2;                    // 20 for a CLV, 460 for the tip vector of the next category's eigen-system
  __device__ __forceinline__ const double *xq(int c) const { return x1 + c * st1; }
  __device__ __forceinline__ const double *xr(int c) const { return x2 + c * st2; }
  __device__ __forceinline__ const double *lq(int c) const { return lut1 + c * S; }
  __device__ __forceinline__ const double *lr(int c) const { return lut2 + c * S; }
  __device__ __forceinline__ const double *dq(int c) const { return drow + c * 40; }
};

// KC: true -> slot c uses the eigen-system of category c (GAMMA; LG4 makes them differ); false -> system 0 (CAT)
template <int QM, int RM, bool KC, class Slots>
__device__ __forceinline__ void aa_entry(double (&carry)[4][3][2], const AaConstsSmem &K,
                                         const Slots &sl, int lane, int tq)
{
  #pragma unroll
  for (int c = 0; c < 4; c++) {
    const double (*fEI)[32] = K.fragEI[KC ? c : 0];
    const double (*fEV)[32] = K.fragEV[KC ? c : 0];
    double y1[3][2], y2[3][2];
    child_y<QM>(y1, carry, fEI, sl.xq(c), sl.lq(c), sl.dq(c), c, lane, tq);
    child_y<RM>(y2, carry, fEI, sl.xr(c), sl.lr(c), sl.dq(c) + S, c, lane, tq);
    // Z = Y1 (.) Y2 in C layout, then X3 = Z * EV
    #pragma unroll
    for (int t = 0; t < 3; t++) { y1[t][0] *= y2[t][0]; y1[t][1] *= y2[t][1]; }
    double x3[3][2];
    #pragma unroll
    for (int t = 0; t < 3; t++) x3[t][0] = x3[t][1] = 0.0;
    #pragma unroll
    for (int s = 0; s < 5; s++) {
      const double az = c_to_a(y1, s, lane, tq);    // Z[g][kappa(s, tq)]
      #pragma unroll
      for (int t = 0; t < 3; t++) dmma(x3[t][0], x3[t][1], az, fEV[s * 3 + t][lane]);
    }
    #pragma unroll
    for (int t = 0; t < 3; t++) { carry[c][t][0] = x3[t][0]; carry[c][t][1] = x3[t][1]; }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------
// Fused traversal on the FP64 tensor pipe.
//
// A warp owns 8 sites (one per "row" g = lane/4) for all four Gamma categories; a CTA of 12 warps owns a
// 96-site chunk.  For every (entry, category) the three 20x20 contractions
//     Y1 = (d_q (.) X1) * EI^T        Y2 = (d_r (.) X2) * EI^T        X3 = (Y1 (.) Y2) * EV
// are [8 sites x 20] x [20 x 20(24)] GEMMs = 15 DMMA.8x8x4 each.  EI^T and EV are read from shared memory as B
// fragments (they do not depend on entry or side because the exponentials are folded into the A operand; they
// depend on the category only under LG4M / LG4X).  The result stays in registers in C-fragment layout ("carry") and
// is converted to an A operand for the parent (c_to_a); it is written to HBM once, 16 B per thread.  Scaling (all 80
// values of a site < 2^-256) is decided after the four categories are done, from the 24 register values of the 4
// lanes that share a site.
// ------------------------------------------------------------------------------------------------------
// CAT = false: GAMMA, slot = category, 8 sites per warp, 96 per chunk, CLV = 80 doubles per site.
// CAT = true : PSR (newviewGTRCATPROT_AVX avxLikelihood.c:487-812, evaluateGTRCATPROT evaluateGenericSpecial.c:1308-1391):
//              slot = one of four 8-site groups of the warp (32 sites per warp, 384 per chunk), CLV = 20 doubles per
//              site, every row scales its A operand with the exponentials of ITS site's category; scaling and the
//              log are per site.
// LG4 = true : every Gamma category has its own eigen-system (LG4M / LG4X partitions in the class): per-category B
//              fragments and tip vectors.  false: one system -- fragment and tip addresses do not depend on the category.
template <bool CAT, bool LG4>
__global__ void __launch_bounds__(kAaThreads, 1)
traverse_aa_kernel(const TraverseArgs a)
{
  constexpr int NSYS = LG4 ? 4 : 1;            // eigen-systems kept in shared memory
  constexpr int TSITES = CAT ? kAaCatChunkSites : kAaChunkSites;   // sites per chunk = bytes per staged tip row
  constexpr int XSPAN = CAT ? 20 : 80;         // doubles per site in a CLV
  constexpr int XSLOT = CAT ? 8 * 20 : 20;     // distance between consecutive slots of my row in a CLV
  __shared__ AaShared sh;
  __shared__ __align__(8) uint64_t mbar;
  extern __shared__ __align__(16) unsigned char dyn[];
  const SEntry *ents = reinterpret_cast<const SEntry *>(dyn);                     // window entries; slot kWindow = branch entry
  const uint8_t *stips = dyn + (size_t)(kWindow + 1) * sizeof(SEntry);             // [2*kWindow + 2][TSITES] staged tip rows
  AaConstsSmem &K = *reinterpret_cast<AaConstsSmem *>(dyn + ((window_smem_bytes<TSITES>() + 127) & ~(size_t)127));

  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const size_t N = a.N;
  const double *model = a.models + (size_t)tile.part_cls * kAaModelDoubles;   // EIGN[20] rates[4] EI[400] EV[400] tip[460]
  {
    for (int i = tid; i < NSYS * 460; i += kAaThreads) K.tip[i / 460][i % 460] = model[kAaTip + i];
    // B fragments per category: EI_c^T[k = 4s+tq][n = 8t+g] = EI_c[n][k];  EV_c[k][n];  zero padding for n >= 20
    for (int i = tid; i < NSYS * 15 * 32; i += kAaThreads) {
      const int kc = i / 480, f = (i % 480) >> 5, ln = i & 31, fs = f / 3, ft = f % 3;
      const int k = kappa(fs, ln & 3), n = 8 * ft + (ln >> 2);
      K.fragEI[kc][f][ln] = n < S ? model[kAaEI + kc * 400 + n * S + k] : 0.0;
      K.fragEV[kc][f][ln] = n < S ? model[kAaEV + kc * 400 + k * S + n] : 0.0;
    }
    if (tid == 0) mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
  }

  const int ncat = CAT ? a.max_cats : 4;                                  // blocks per d-table / tip-table row
  const int drowlen = CAT ? a.dtab_row : kAaDtabRow;                      // (compile-time constants under GAMMA)
  const double *drow0 = a.dtab + (size_t)tile.part_cls * drowlen;         // + e * dstride + cat*40 + side*20 + k
  const size_t dstride = (size_t)a.nparts_cls * drowlen;
  const size_t lutside = CAT ? (size_t)23 * ncat * S : (size_t)kAaLutRow; // doubles per (entry, partition, side)
  double acc = 0.0;
  uint32_t phase = 0;

  for (unsigned chunk0 = 0; chunk0 < tile.nsites; chunk0 += TSITES) {
    const unsigned nchunk = min(tile.nsites - chunk0, (unsigned)TSITES);
    // my row's site in every slot: GAMMA -- the same site four times; CAT -- four sites 8 apart
    const unsigned i0 = (unsigned)(CAT ? warp * 32 + g : warp * 8 + g);
    bool valid[4]; int lsite[4], cat[4];
    #pragma unroll
    for (int c = 0; c < 4; c++) {
      const unsigned i = CAT ? i0 + 8u * c : i0;
      valid[c] = i < nchunk;
      lsite[c] = valid[c] ? (int)i : 0;
      cat[c] = CAT ? a.rate_cat[(size_t)tile.site0 + chunk0 + lsite[c]] : c;
    }
    const size_t site0 = (size_t)tile.site0 + chunk0 + (CAT ? i0 : lsite[0]);   // CAT: slot c sits at site0 + 8c
    const size_t soff = site0 * XSPAN;
    const bool warp_active = (unsigned)(CAT ? warp * 32 : warp * 8) < nchunk;      // any site of this chunk in my warp?
    double carry[4][3][2];                                // CLV of the entry just computed, C-fragment layout
    #pragma unroll
    for (int c = 0; c < 4; c++)
      #pragma unroll
      for (int t = 0; t < 3; t++) carry[c][t][0] = carry[c][t][1] = 0.0;

    int w0 = a.first;
    do {
    const int w1 = min(w0 + kWindow, a.count);
    phase = stage_window<TSITES>(dyn, &mbar, phase, a.entries, a.tips, a.mxtips, N, XSPAN, (size_t)tile.site0 + chunk0, nchunk, w0, w1,
                         a.first, a.count, w0 == a.first && a.mode != MODE_NEWVIEW, nullptr, 0, 0,
                         SevView{CAT ? nullptr : a.sev_table, a.sev_gapcol, a.clv, a.sev_ntiles, a.sev_tile0 + (int)blockIdx.x, tile.part_cls,
                                 a.nparts_cls, XSPAN, tile.site0}, a.scratch_off);
    for (int e = w0; e < w1; e++) {
      const SEntry en = ents[e - w0];
      if (en.flags & kFlagGapP) continue;       // -S: all sites of the tile are gaps below this node: nothing to compute or store
      if (!warp_active) {
        // a warp without sites in this (last, partial) chunk of a tile does no tensor work -- the chunk then costs what its
        // busiest scheduler costs; its slot of a scoring entry still has to read zero
        if (!CAT && (en.flags & kFlagScore) && lane == 0 && chunk0 == 0)
          a.score_out[(size_t)en.poff * a.score_stride + (size_t)(a.sev_tile0 + (int)blockIdx.x) * (kAaThreads / 32) + warp] = make_double2(0.0, 0.0);
        continue;
      }
      const double *drow = drow0 + (size_t)e * dstride;
      const bool q_inner = en.tipCase == 2, r_inner = en.tipCase != 0;
      const int qsrc = !q_inner ? SRC_TIP : ((en.flags & kFlagCarryA) ? SRC_CARRY : SRC_LOAD);
      const int rsrc = !r_inner ? SRC_TIP : ((en.flags & kFlagCarryB) ? SRC_CARRY : SRC_LOAD);
      const double *lutE = a.lut ? a.lut + ((size_t)e * a.nparts_cls + tile.part_cls) * 2 * lutside : nullptr;
      // memory sources: a global CLV or the tip vector in shared memory; tip tables when they exist
      typename std::conditional<CAT, AaSlotsCat, AaSlotsGamma>::type sl;
      if constexpr (CAT) {
        #pragma unroll
        for (int c = 0; c < 4; c++) {
          const int code1 = q_inner ? 0 : stips[en.qoff + lsite[c]], code2 = r_inner ? 0 : stips[en.roff + lsite[c]];
          sl.xq_[c] = q_inner ? a.clv + en.qoff + soff + c * XSLOT : &K.tip[0][S * code1];
          sl.xr_[c] = r_inner ? a.clv + en.roff + soff + c * XSLOT : &K.tip[0][S * code2];
          sl.dq_[c] = drow + cat[c] * 40;
          sl.lq_[c] = lutE + ((size_t)code1 * ncat + cat[c]) * S;
          sl.lr_[c] = lutE + lutside + ((size_t)code2 * ncat + cat[c]) * S;
        }
      } else {
        const int code1 = q_inner ? 0 : stips[en.qoff + lsite[0]], code2 = r_inner ? 0 : stips[en.roff + lsite[0]];
        sl.x1 = q_inner ? a.clv + en.qoff + ((en.flags & kFlagGapA) ? 0 : soff) : &K.tip[0][S * code1];    // -S: gap column of the node
        sl.x2 = r_inner ? a.clv + en.roff + ((en.flags & kFlagGapB) ? 0 : soff) : &K.tip[0][S * code2];
        sl.st1 = q_inner ? S : (LG4 ? 460 : 0); sl.st2 = r_inner ? S : (LG4 ? 460 : 0);
        sl.lut1 = lutE + (size_t)code1 * SPAN; sl.lut2 = lutE + lutside + (size_t)code2 * SPAN;
        sl.drow = drow;
      }
      const bool qlut = !q_inner && lutE, rlut = !r_inner && lutE;
      if (qlut && rlut)                   aa_entry<CH_LUT, CH_LUT, LG4>(carry, K, sl, lane, tq);
      else if (qlut && rsrc == SRC_CARRY) aa_entry<CH_LUT, CH_CARRY, LG4>(carry, K, sl, lane, tq);
      else if (qlut)                      aa_entry<CH_LUT, CH_MEM, LG4>(carry, K, sl, lane, tq);
      else if (rsrc == SRC_CARRY)         aa_entry<CH_MEM, CH_CARRY, LG4>(carry, K, sl, lane, tq);
      else if (qsrc == SRC_CARRY)         aa_entry<CH_CARRY, CH_MEM, LG4>(carry, K, sl, lane, tq);
      else                                aa_entry<CH_MEM, CH_MEM, LG4>(carry, K, sl, lane, tq);

      unsigned int cnt = 0;
      if (r_inner) {                            // TIP_TIP is never rescaled (avxLikelihood.c:1393-1440)
        if (!CAT) {
          unsigned int hi = 0;
          #pragma unroll
          for (int c = 0; c < 4; c++)
            #pragma unroll
            for (int t = 0; t < 3; t++)            // padding columns hold exact zeros and pass the test
              hi = max(hi, max(abs_hi(carry[c][t][0]), abs_hi(carry[c][t][1])));
          const unsigned int b = __ballot_sync(0xffffffffu, hi < kMinLikelihoodHi);
          if (((b >> (lane & ~3)) & 0xFu) == 0xFu) {            // all 80 values of the site are tiny
            #pragma unroll
            for (int c = 0; c < 4; c++)
              #pragma unroll
              for (int t = 0; t < 3; t++) { carry[c][t][0] *= kTwoToThe256; carry[c][t][1] *= kTwoToThe256; }
            if (tq == 0 && valid[0]) cnt = (unsigned int)a.wgt[site0];
          }
        } else {
          #pragma unroll
          for (int c = 0; c < 4; c++) {            // one site per slot: all 20 values of that site tiny
            unsigned int hi = 0;
            #pragma unroll
            for (int t = 0; t < 3; t++) hi = max(hi, max(abs_hi(carry[c][t][0]), abs_hi(carry[c][t][1])));
            const unsigned int b = __ballot_sync(0xffffffffu, hi < kMinLikelihoodHi);
            if (((b >> (lane & ~3)) & 0xFu) == 0xFu) {
              #pragma unroll
              for (int t = 0; t < 3; t++) { carry[c][t][0] *= kTwoToThe256; carry[c][t][1] *= kTwoToThe256; }
              if (tq == 0 && valid[c]) cnt += (unsigned int)a.wgt[site0 + 8 * c];
            }
          }
        }
      }
      if (!CAT && (en.flags & kFlagScore)) {
        // exb_score_insertions: the CLV just formed (carry, C-fragment layout) is the candidate's new node; contract it with
        // the pruned subtree's CLV under the diagonal table of the branch p--s (evaluateGTRGAMMAPROT
        // evaluateGenericSpecial.c:1393) and leave this warp's share in its own slot of score_out
        const SEntry br = ents[kWindow];
        const bool stip = br.flags & kFlagPTip;
        const bool lg4 = LG4 && model[kAaLg4] != 0.0;
        double term = 0.0;
        #pragma unroll
        for (int c = 0; c < 4; c++) {
          const double *tps = stip ? &K.tip[LG4 ? c : 0][S * stips[br.poff + lsite[0]]] : K.tip[0];
          double tc = 0.0;
          #pragma unroll
          for (int t = 0; t < 3; t++) {
            const int col = 8 * t + 2 * tq;
            if (col >= S) continue;               // t == 2 && tq >= 2: padding
            const double2 xs = stip ? *reinterpret_cast<const double2 *>(tps + col) : ld128(a.clv + br.poff + soff + c * XSLOT + col);
            const double d0 = __ldg(drow0 + c * 40 + col), d1 = __ldg(drow0 + c * 40 + col + 1);   // row 0, q side
            tc += carry[c][t][0] * xs.x * d0;
            tc += carry[c][t][1] * xs.y * d1;
          }
          term += lg4 ? model[kAaWeights + c] * tc : tc;
        }
        term += __shfl_xor_sync(0xffffffffu, term, 1);
        term += __shfl_xor_sync(0xffffffffu, term, 2);
        double part = (tq == 0 && valid[0]) ? (double)a.wgt[site0] * (lg4 ? log(fabs(term)) : log(0.25 * fabs(term))) : 0.0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
        const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) {
          double2 *slot = a.score_out + (size_t)en.poff * a.score_stride + (size_t)(a.sev_tile0 + (int)blockIdx.x) * (kAaThreads / 32) + warp;
          double2 v = make_double2(part, (double)tot);
          if (chunk0) { const double2 old = *slot; v.x += old.x; v.y += old.y; }
          *slot = v;
        }
        __syncwarp();
        continue;
      }
      if (__any_sync(0xffffffffu, cnt != 0)) {
        const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
      }
      if (CAT || valid[0]) {
        double *dst = a.clv + en.poff + soff + 2 * tq;
        #pragma unroll
        for (int c = 0; c < 4; c++) {
          if (CAT && !valid[c]) continue;
          *reinterpret_cast<double2 *>(dst + c * XSLOT) = make_double2(carry[c][0][0], carry[c][0][1]);
          *reinterpret_cast<double2 *>(dst + c * XSLOT + 8) = make_double2(carry[c][1][0], carry[c][1][1]);
          if (tq < 2) *reinterpret_cast<double2 *>(dst + c * XSLOT + 16) = make_double2(carry[c][2][0], carry[c][2][1]);
        }
      }
      __syncwarp();
    }
    } while ((w0 += kWindow) < a.count);

    if (a.mode == MODE_NEWVIEW || a.mode == MODE_SCORE) continue;

    // ---------------- branch (entry 0): evaluate or sum table (C-fragment layout on both ends) ----------------
    const SEntry br = ents[kWindow];
    const bool ptip = br.flags & kFlagPTip, qtip = br.flags & kFlagQTip;
    const bool p_carry = br.flags & kFlagCarryA, q_carry = br.flags & kFlagCarryB;
    const bool lg4 = LG4 && model[kAaLg4] != 0.0;       // an ordinary partition next to LG4 ones keeps log(0.25 |sum|)
    double term = 0.0;
    #pragma unroll
    for (int c = 0; c < 4; c++) {
      const double *tpa = ptip ? &K.tip[LG4 ? c : 0][S * stips[br.poff + lsite[c]]] : K.tip[0];
      const double *tpb = qtip ? &K.tip[LG4 ? c : 0][S * stips[br.qoff + lsite[c]]] : K.tip[0];
      double tc = 0.0;                            // this category's share (weighted under LG4X)
      #pragma unroll
      for (int t = 0; t < 3; t++) {
        const int col = 8 * t + 2 * tq;
        const bool live = col < S;                // t == 2 && tq >= 2: padding
        double2 xa = make_double2(0.0, 0.0), xb = make_double2(0.0, 0.0);
        if (p_carry) xa = make_double2(carry[c][t][0], carry[c][t][1]);
        else if (live) xa = ptip ? *reinterpret_cast<const double2 *>(tpa + col)
                                 : ld128(a.clv + br.poff + ((br.flags & kFlagGapA) ? 0 : soff) + c * XSLOT + col);
        if (q_carry) xb = make_double2(carry[c][t][0], carry[c][t][1]);
        else if (live) xb = qtip ? *reinterpret_cast<const double2 *>(tpb + col)
                                 : ld128(a.clv + br.qoff + ((br.flags & kFlagGapB) ? 0 : soff) + c * XSLOT + col);
        if (a.mode == MODE_SUM) {
          if (valid[c] && live) *reinterpret_cast<double2 *>(a.sum + soff + c * XSLOT + col) = make_double2(xa.x * xb.x, xa.y * xb.y);
        } else if (live) {
          const double d0 = __ldg(drow0 + cat[c] * 40 + col), d1 = __ldg(drow0 + cat[c] * 40 + col + 1);   // row 0, q side
          tc += xa.x * xb.x * d0;
          tc += xa.y * xb.y * d1;
        }
      }
      term += lg4 ? model[kAaWeights + c] * tc : tc;
      if (CAT && a.mode == MODE_EVALUATE) {       // one site per slot: its own log (no 1/4: a single category)
        term += __shfl_xor_sync(0xffffffffu, term, 1);
        term += __shfl_xor_sync(0xffffffffu, term, 2);
        if (tq == 0 && valid[c]) acc += (double)a.wgt[site0 + 8 * c] * log(fabs(term));
        term = 0.0;
      }
    }
    if (!CAT && a.mode == MODE_EVALUATE) {
      term += __shfl_xor_sync(0xffffffffu, term, 1);
      term += __shfl_xor_sync(0xffffffffu, term, 2);
      if (tq == 0 && valid[0]) acc += (double)a.wgt[site0] * (lg4 ? log(fabs(term)) : log(0.25 * fabs(term)));
    }
  }

  if (a.mode == MODE_EVALUATE) {
    const double tot = block_sum(acc, sh.red);
    if (tid == 0) { a.tile_out[2 * (size_t)blockIdx.x] = tot; a.tile_out[2 * (size_t)blockIdx.x + 1] = 0.0; }
  }
}

// ------------------------------------------------------------------------------------------------------
// Tip look-up table: for every (entry, partition, side) whose child is a tip,
//   lut[code][c][l] = sum_j EI[l][j] * d[c][side][j] * tipVector[code][j]        (23 codes x 4 x 20)
// -- the reference's umpX1 / umpX2 (avxLikelihood.c:1353-1389), built once per call for the whole descriptor.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
aa_lut_kernel(const LutArgs a)
{
  extern __shared__ double lsm[];                 // EI[4][400] | tip[4][460] | d[ncat][20]
  double *EI = lsm, *tip = lsm + 1600, *d = lsm + 1600 + 1840;
  const int side = blockIdx.x & 1;
  const int ep = blockIdx.x >> 1, pc = ep % a.nparts_cls, e = ep / a.nparts_cls;
  if (e < a.first) return;
  const int tc = a.entries[e].tipCase & 3;        // (scoring entries of exb_score_insertions carry kScoreBit)
  if (!(side == 0 ? tc <= 1 : tc == 0)) return;
  const int ncat = a.ncat, row = 23 * ncat * 20;
  const double *model = a.models + (size_t)pc * kAaModelDoubles;
  const double *drow = a.dtab + ((size_t)e * a.nparts_cls + pc) * a.dtab_row;
  for (int i = threadIdx.x; i < 1600; i += 256) EI[i] = model[kAaEI + i];
  for (int i = threadIdx.x; i < 1840; i += 256) tip[i] = model[kAaTip + i];
  for (int i = threadIdx.x; i < ncat * 20; i += 256) d[i] = drow[(i / 20) * 40 + side * 20 + i % 20];
  __syncthreads();
  double *out = a.lut + ((size_t)ep * 2 + side) * row;
  for (int i = threadIdx.x; i < row; i += 256) {
    const int code = i / (ncat * 20), c = (i % (ncat * 20)) / 20, l = i % 20;
    const int kc = a.cat ? 0 : c;                    // the eigen-system of the category (CAT: a single one)
    const double *E = EI + kc * 400, *tv = tip + kc * 460 + code * 20;
    double s = tv[0];                                // j = 0: EI[l][0] = 1, d[0] = 1
    #pragma unroll
    for (int j = 1; j < 20; j++) s += E[l * 20 + j] * (d[c * 20 + j] * tv[j]);
    out[i] = s;
  }
}

void launch_aa_lut(const LutArgs &a, cudaStream_t s)
{
  const int blocks = a.count * a.nparts_cls * 2;
  if (blocks <= 0) return;
  const size_t smem = (size_t)(1600 + 1840 + a.ncat * 20) * sizeof(double);
  static size_t configured[kMaxDevices];
  ensure_dynamic_smem(aa_lut_kernel, 64 * 1024, configured);
  aa_lut_kernel<<<blocks, 256, smem, s>>>(a);
}

__global__ void __launch_bounds__(kThreads, 2)
core_aa_kernel(const CoreArgs a)
{
  __shared__ union { CoreSmem g; CoreCatSmem c; } sh;
  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  double t1, t2;
  const double *model = a.models + (size_t)tile.part_cls * kAaModelDoubles;
  const double lz = a.lz[a.nbranches > 1 ? tile.part : 0];
  if (a.rate_cat) core_tile_cat<20>(tile, a.sum, a.wgt, model, a.rate_cat, a.cat_rates + (size_t)tile.part_cls * a.max_cats, a.max_cats, lz, sh.c, t1, t2);
  else            core_tile<20>(tile, a.sum, a.wgt, model, lz, sh.g, t1, t2);
  if (threadIdx.x == 0) { a.tile_out[2 * (size_t)blockIdx.x] = t1; a.tile_out[2 * (size_t)blockIdx.x + 1] = t2; }
}

template <bool CAT, bool LG4>
static void launch_aa_variant(const TraverseArgs &a, int ntiles, cudaStream_t s)
{
  const size_t smem = ((window_smem_bytes<(CAT ? kAaCatChunkSites : kAaChunkSites)>() + 127) & ~(size_t)127) + sizeof(AaConstsSmem);
  static size_t configured[kMaxDevices];
  ensure_dynamic_smem(traverse_aa_kernel<CAT, LG4>, std::min<size_t>(227 * 1024 - 16384, std::max<size_t>(smem, 64 * 1024)), configured);
  traverse_aa_kernel<CAT, LG4><<<ntiles, kAaThreads, smem, s>>>(a);
}

void launch_traverse_aa(const TraverseArgs &a, bool any_lg4, int ntiles, cudaStream_t s)
{
  if (ntiles <= 0) return;
  if (a.rate_cat)   launch_aa_variant<true, false>(a, ntiles, s);
  else if (any_lg4) launch_aa_variant<false, true>(a, ntiles, s);
  else              launch_aa_variant<false, false>(a, ntiles, s);
}
void launch_core_aa(const CoreArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles > 0) core_aa_kernel<<<ntiles, kThreads, 0, s>>>(a);
}

}  // namespace exb
