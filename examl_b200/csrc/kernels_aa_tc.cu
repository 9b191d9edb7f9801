// kernels_aa_tc.cu -- the 20-state (protein) GAMMA newview on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// north_star asks for the AA contraction on tcgen05 tiles.  tcgen05.mma has no FP64 kind, so this path is the
// engine's MIXED-PRECISION mode (exb_set_aa_precision(EXB_AA_TF32X3), off by default): the three 20x20 contractions
// of an entry
//     Y1 = (d_q (.) X1) * EI^T        Y2 = (d_r (.) X2) * EI^T        X3 = (Y1 (.) Y2) * EV
// run as 3xTF32 products -- every operand split into a TF32 "hi" and a TF32 "lo" part, hi*hi + hi*lo + lo*hi,
// FP32 accumulators in tensor memory -- with the site vectors renormalised by an exact power of two at every node so
// that FP32 holds their range.  tools/mixed_precision_study.py measured what that costs on the reference's own
// testData/140 and on the C4 shape: lnL within ~1e-8 relative of the FP64 result (the north_star bar is 1e-6), per-site
// log-likelihoods within ~2e-3.  CLVs are still stored as FP64 in the same arena with the reference's own scaling
// convention (2^256 multiplier, scaler counts), so evaluate / sum / core / exb_read_clv work unchanged on them.
//
// Replaces, in that mode, the per-site loops of newviewGTRGAMMAPROT_AVX (avxLikelihood.c:1312-1800).
//
// Mapping: a CTA of 128 threads works on 32 sites at a time; thread = (site, Gamma category) = TMEM lane = row of every
// MMA operand (the four rows of a site are adjacent lanes), so one tensor round serves all four categories.  Per entry:
//   inner child   rows  d (.) x  -> TF32 hi / lo -> tcgen05.st into TMEM;  one elected thread issues
//                 tcgen05.mma.kind::tf32 M=128 K=8:  A_hi x [EI_hi | EI_lo] (N=64)  +  A_lo x EI_hi (N=32), 3 k-steps each,
//                 tcgen05.commit -> mbarrier;  all rows tcgen05.ld their Y (hi*hi + lo*hi columns, hi*lo columns)
//   tip child     no tensor work: Y comes from the per-call look-up table P(t) * tipVector[code] (aa_lut_kernel)
//   parent        Z = Y1 (.) Y2 -> hi / lo -> TMEM, the same six MMAs with [EV_hi | EV_lo], tcgen05.ld
// B operands (EI^T, EV: constants of the partition) sit in shared memory in the canonical K-major layout.  The row just
// produced stays in registers as 20 floats + the site's exponent ("carry").  128 TMEM columns and < 128 registers per
// thread -> four CTAs per SM overlap each other's tensor and CUDA-core phases; a warp reads / writes 5 KB of contiguous
// CLV per entry.
#include "engine.cuh"
#include "stage.cuh"
#include <algorithm>

namespace exb {

namespace {

constexpr int S = 20, SPAN = 80;
constexpr int kTcThreads = 128;       // = rows of an MMA: 32 sites x 4 Gamma categories
constexpr int kTcSites = 32;          // sites per chunk
constexpr int kTcWindow = 96;
constexpr int kTcCols = 128;          // TMEM columns per CTA -> four CTAs per SM
// TMEM column map: A operand (hi at +0, lo at +24) | D accumulators (64 columns)
constexpr uint32_t kColA = 0, kColD = 64;
// instruction descriptors: kind::tf32, FP32 accumulate, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// B operand in shared memory, canonical K-major layout without swizzle: core matrix = 8 rows (n) x 16 bytes (4 k);
// N = 64 rows (0..31: hi part of the matrix, 32..63: lo part), K = 24
constexpr int kBWords = 6 * 8 * 32;                       // 6 k-chunks x 8 n-groups x 32 words
__device__ __forceinline__ int b_word(int n, int k) { return ((k >> 2) * 8 + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3); }

struct TcShared {
  uint32_t bEI[kBWords];     // B[n = l][k = j] = EI[l][j]   (hi | lo)
  uint32_t bEV[kBWords];     // B[n = m][k = l] = EV[l][m]
  uint64_t mbar_tma, mbar_mma;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t f2tf32(float x)       // round to nearest TF32 (low 13 mantissa bits zero)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));                   // rounded, not left to the tensor core's truncation
}
__device__ __forceinline__ double pow2d(int e) { return __hiloint2double((1023 + e) << 20, 0); }      // -1022 <= e <= 1023
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((uint32_t)(127 + e) << 23); }   // -126 <= e <= 127

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
  // start address | leading (K) byte offset = 8 n-groups * 128 B | stride (N) byte offset = 128 B | version 1 | no swizzle
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t id, uint32_t accumulate)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(id), "r"(accumulate), "r"(0u) : "memory");
}
// hi x [B_hi | B_lo] (N = 64) and lo x B_hi (N = 32) into the 64 columns at d: three k-steps each
__device__ __forceinline__ void mma_3xtf32(uint32_t d, uint32_t a_hi, uint32_t b_saddr)
{
  #pragma unroll
  for (int j = 0; j < 3; j++) mma_tf32(d, a_hi + 8u * j, smem_desc(b_saddr + (uint32_t)(2 * j) * 1024u), idesc(64), j > 0);
  #pragma unroll
  for (int j = 0; j < 3; j++) mma_tf32(d, a_hi + 24u + 8u * j, smem_desc(b_saddr + (uint32_t)(2 * j) * 1024u), idesc(32), 1u);
}
// 24 columns (20 values + 4 zeros) of this thread's row -> TMEM
__device__ __forceinline__ void tmem_store24(uint32_t taddr, const uint32_t (&v)[20])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr + 8u), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr + 16u), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
// 20 columns of this thread's row <- TMEM
__device__ __forceinline__ void tmem_load20(uint32_t taddr, float (&v)[20])
{
  uint32_t r[20];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]) : "r"(taddr + 16u) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  #pragma unroll
  for (int i = 0; i < 20; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

}  // namespace

__global__ void __launch_bounds__(kTcThreads, 4)
traverse_aa_tc_kernel(const TraverseArgs a)
{
  constexpr int TS = kTcSites, KW = kTcWindow;
  extern __shared__ __align__(128) unsigned char dyn[];
  const SEntry *ents = reinterpret_cast<const SEntry *>(dyn);
  const uint8_t *stips = dyn + (size_t)(KW + 1) * sizeof(SEntry);
  TcShared &sh = *reinterpret_cast<TcShared *>(dyn + ((window_smem_bytes<TS, KW>() + 127) & ~(size_t)127));

  const Tile tile = a.tiles[blockIdx.x];
  if (!a.mask[tile.part]) return;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int c = tid & 3;                        // this row's Gamma category; the 4 rows of a site are adjacent lanes
  const size_t N = a.N;
  const double *model = a.models + (size_t)tile.part_cls * kAaModelDoubles;

  // ---- one-time set-up: B operands (hi | lo) in shared memory, barriers, tensor memory ----
  for (int i = tid; i < 64 * 24; i += kTcThreads) {
    const int n = i / 24, k = i % 24, nn = n & 31;
    float ei = 0.f, ev = 0.f;
    if (nn < S && k < S) { ei = (float)model[kAaEI + nn * S + k]; ev = (float)model[kAaEV + k * S + nn]; }
    uint32_t h, l;
    split_tf32(ei, h, l); sh.bEI[b_word(n, k)] = n < 32 ? h : l;
    split_tf32(ev, h, l); sh.bEV[b_word(n, k)] = n < 32 ? h : l;
  }
  if (tid == 0) {
    mbar_init(&sh.mbar_tma, 1);
    mbar_init(&sh.mbar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(stage_smem_u32(&sh.tmem_base)), "r"(kTcCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the B operands are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = sh.tmem_base;
  const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);         // this warp's 32 lanes; thread = lane = MMA row
  const uint32_t sEI = stage_smem_u32(sh.bEI), sEV = stage_smem_u32(sh.bEV);

  const double *drow0 = a.dtab + (size_t)tile.part_cls * kAaDtabRow + c * 40;   // + e * dstride + side*20 + k
  const size_t dstride = (size_t)a.nparts_cls * kAaDtabRow;
  uint32_t tma_phase = 0, mma_phase = 0;

  // one tensor round: the 128 rows just stored at kColA times B (3xTF32) -> 20 values per row
  auto tensor_round = [&](uint32_t b_saddr, float (&out)[S]) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      mma_3xtf32(tbase + kColD, tbase + kColA, b_saddr);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(stage_smem_u32(&sh.mbar_mma)) : "memory");
    }
    mbar_wait(&sh.mbar_mma, mma_phase); mma_phase ^= 1u;
    tc_fence_after();
    float t[S];
    tmem_load20(trow + kColD, out); tmem_load20(trow + kColD + 32u, t);      // hi*hi + lo*hi  |  hi*lo
    #pragma unroll
    for (int k = 0; k < S; k++) out[k] += t[k];
  };

  for (unsigned chunk0 = 0; chunk0 < tile.nsites; chunk0 += TS) {
    const unsigned nchunk = min(tile.nsites - chunk0, (unsigned)TS);
    const bool valid = (unsigned)(tid >> 2) < nchunk;
    const int lsite = valid ? (tid >> 2) : 0;
    const size_t site = (size_t)tile.site0 + chunk0 + lsite;
    const size_t soff = site * SPAN + c * S;    // this row's 20 doubles inside a CLV
    float carry[S];                             // this row of the previous entry's CLV: carry * 2^cexp (cexp per site)
    int cexp = 0;
    #pragma unroll
    for (int i = 0; i < S; i++) carry[i] = 0.f;

    int w0 = a.first;
    do {
    const int w1 = min(w0 + KW, a.count);
    tma_phase = stage_window<TS, KW>(dyn, &sh.mbar_tma, tma_phase, a.entries, a.tips, a.mxtips, N, SPAN, (size_t)tile.site0 + chunk0,
                                     nchunk, w0, w1, a.first, a.count, false);
    for (int e = w0; e < w1; e++) {
      const SEntry en = ents[e - w0];
      const double *drow = drow0 + (size_t)e * dstride;
      const bool q_inner = en.tipCase == 2, r_inner = en.tipCase != 0;
      const bool q_carry = en.flags & kFlagCarryA, r_carry = en.flags & kFlagCarryB;
      const double *lutE = a.lut + ((size_t)e * a.nparts_cls + tile.part_cls) * 2 * kAaLutRow + c * S;

      // Y of one child: inner -> (d (.) x) in 3xTF32 through the tensor core; tip -> the per-call table
      // P(t) * tipVector[code] (aa_lut_kernel), no tensor work.  Returns the binary exponent of the site's vector.
      auto child = [&](bool inner, bool from_carry, long long off, const double *dside, const double *lut_side, float (&y)[S]) -> int {
        if (!inner) {
          const double *r = lut_side + (size_t)stips[off + lsite] * SPAN;
          #pragma unroll
          for (int k = 0; k < S; k += 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(r + k)); y[k] = (float)t.x; y[k + 1] = (float)t.y; }
          return 0;
        }
        float x[S];
        int ex = cexp;
        if (from_carry) {
          #pragma unroll
          for (int k = 0; k < S; k++) x[k] = carry[k];
        } else {                                 // read my 20 doubles; renormalise the SITE (4 adjacent lanes) to [0.5, 1)
          double v[S];
          const double *src = a.clv + off + soff;
          #pragma unroll
          for (int k = 0; k < S; k += 4) {
            const double4 t = *reinterpret_cast<const double4 *>(src + k);
            v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
          }
          unsigned int hm = 0;
          #pragma unroll
          for (int k = 0; k < S; k++) hm = max(hm, (unsigned int)__double2hiint(v[k]) & 0x7fffffffu);
          hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 1));
          hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 2));
          ex = hm ? (int)(hm >> 20) - 1022 : 0;                    // max |v| of the site in [2^(ex-1), 2^ex)
          const double sc = pow2d(-ex);
          #pragma unroll
          for (int k = 0; k < S; k++) x[k] = (float)(v[k] * sc);
        }
        uint32_t hi[S], lo[S];
        #pragma unroll
        for (int k = 0; k < S; k += 4) {
          const double2 d01 = __ldg(reinterpret_cast<const double2 *>(dside + k)), d23 = __ldg(reinterpret_cast<const double2 *>(dside + k + 2));
          const double4 d = make_double4(d01.x, d01.y, d23.x, d23.y);
          split_tf32(x[k] * (float)d.x, hi[k], lo[k]);         split_tf32(x[k + 1] * (float)d.y, hi[k + 1], lo[k + 1]);
          split_tf32(x[k + 2] * (float)d.z, hi[k + 2], lo[k + 2]); split_tf32(x[k + 3] * (float)d.w, hi[k + 3], lo[k + 3]);
        }
        tmem_store24(trow + kColA, hi); tmem_store24(trow + kColA + 24u, lo);
        tensor_round(sEI, y);
        return ex;
      };
      float y1[S], y2[S];
      const int eq = child(q_inner, q_carry, en.qoff, drow, lutE, y1);
      const int er = child(r_inner, r_carry, en.roff, drow + S, lutE + kAaLutRow, y2);

      // ---- Z = Y1 (.) Y2 -> TMEM, tensor round with EV: this row of the parent ----
      {
        uint32_t hi[S], lo[S];
        #pragma unroll
        for (int k = 0; k < S; k++) split_tf32(y1[k] * y2[k], hi[k], lo[k]);
        tmem_store24(trow + kColA, hi); tmem_store24(trow + kColA + 24u, lo);
      }
      tensor_round(sEV, carry);

      // ---- renormalise the site (exact power of two), apply the reference's scaling rule, store as FP64 ----
      unsigned int bm = 0;
      #pragma unroll
      for (int k = 0; k < S; k++) bm = max(bm, __float_as_uint(carry[k]) & 0x7fffffffu);
      bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
      bm = max(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
      const int k2 = bm >= 0x00800000u ? (int)(bm >> 23) - 126 : 0;       // max |x3| of the site in [2^(k2-1), 2^k2)
      const float fs = pow2f(-k2);
      int ex = eq + er + k2;
      #pragma unroll
      for (int k = 0; k < S; k++) carry[k] *= fs;
      unsigned int cnt = 0;
      if (r_inner && bm != 0 && ex <= -256) {        // all 80 values < 2^-256: the reference multiplies by 2^256 (avxLikelihood.c:1756-1787)
        ex += 256;
        if (c == 0 && valid) cnt = (unsigned int)a.wgt[site];
      }
      cexp = max(ex, -1000);
      if (valid) {
        const double sc = pow2d(cexp);
        double *dst = a.clv + en.poff + soff;
        #pragma unroll
        for (int k = 0; k < S; k += 4)
          *reinterpret_cast<double4 *>(dst + k) = make_double4((double)carry[k] * sc, (double)carry[k + 1] * sc, (double)carry[k + 2] * sc, (double)carry[k + 3] * sc);
      }
      if (r_inner && __any_sync(0xffffffffu, cnt != 0)) {
        const unsigned int tot = __reduce_add_sync(0xffffffffu, cnt);
        if ((tid & 31) == 0) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + tile.part], tot);
      }
    }
    } while ((w0 += KW) < a.count);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kTcCols) : "memory");
}

void launch_traverse_aa_tc(const TraverseArgs &a, int ntiles, cudaStream_t s)
{
  if (ntiles <= 0) return;
  const size_t smem = ((window_smem_bytes<kTcSites, kTcWindow>() + 127) & ~(size_t)127) + sizeof(TcShared) + 128;
  traverse_aa_tc_kernel<<<ntiles, kTcThreads, smem, s>>>(a);
}

}  // namespace exb
