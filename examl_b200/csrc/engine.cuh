// engine.cuh -- structures shared by the host engine (engine.cu) and the sm_100a kernels
// (kernels_dna.cu, kernels_aa.cu).  See DESIGN.md for the HBM layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exb {

constexpr double kZmin = 1.0E-15;                    // axml.h:110
constexpr double kTwoToThe256 =                      // axml.h:113-116
    115792089237316195423570985008687907853269984665640564039457584007913129639936.0;
constexpr double kMinLikelihood = 1.0 / kTwoToThe256;  // axml.h:118

enum Mode : int { MODE_NEWVIEW = 0, MODE_EVALUATE = 1, MODE_SUM = 2, MODE_CORE = 3, MODE_SCORE = 4 };

// Data-type class: all partitions with the same number of states share one site axis.
constexpr int CLS_DNA = 0;   // 4 states, 16 tip codes, span 16 doubles  (128 B per site and node)
constexpr int CLS_AA  = 1;   // 20 states, 23 tip codes, span 80 doubles (640 B per site and node)

// Per-partition model constants as laid out in HBM (doubles).
//   DNA: EIGN[4] rates[4] EI[16] EV[16] tip[16*4]            = 104 doubles
//   AA : one eigen-system PER GAMMA CATEGORY (LG4M / LG4X, axml.h:560-575; ordinary models store four copies):
//        EIGN[4][20] rates[4] weights[4] lg4-flag pad | EI[4][400] | EV[4][400] | tip[4][23*20]   = 5136 doubles
//        weights = category weights of LG4X, 1.0 otherwise (they cancel in the derivative ratios); the flag selects
//        log(|sum w_c t_c|) (LG4, evaluateGenericSpecial.c:1164-1287) vs log(0.25 |sum t_c|) in evaluate.
constexpr int kDnaModelDoubles = 4 + 4 + 16 + 16 + 64;
constexpr int kAaEign = 0, kAaRates = 80, kAaWeights = 84, kAaLg4 = 88, kAaEI = 96, kAaEV = 1696, kAaTip = 3296;
constexpr int kAaModelDoubles  = 5136;

// A work tile: a run of sites of ONE partition, processed by one CTA for the whole traversal.
struct Tile {
  uint32_t site0;      // first site on the class site axis
  uint32_t nsites;     // <= tile_sites
  int32_t  part;       // global partition index (mask / scaler / branch slot)
  int32_t  part_cls;   // index of the partition inside its class (model constants)
};

// traversalInfo without branch arrays (mirrors exb_entry)
//   tipCase: -1 branch entry | 0..2 newview | (0..2) + kScoreBit: scoring entry of exb_score_insertions (p = -(candidate + 1))
struct DevEntry { int tipCase, p, q, r; };
constexpr int kScoreBit = 4;
constexpr int kScratchNodes = 64;                    // EXB_SCRATCH_NODES: node numbers 2*mxtips-1 .. 2*mxtips-2+kScratchNodes

// P(t) eigen-factor table, built once per call by dtab_kernel:
//   DNA row (32 doubles) per (entry, class-partition): [c][side q/r][k] = exp(EIGN[k] * rate[c] * lz_side)
//   AA  row (160 doubles):                             [c][side][k],  k < 20
// Row 0 (the branch entry of evaluate) carries the diagonal table of calcDiagptable in its q side.
constexpr int kDnaDtabRow = 32;
constexpr int kAaDtabRow  = 160;

// CAT / PSR (one rate category per site, axml.h:349): the table has one [2 sides][S] block per category of the
// partition, row = max_cats * 2 * S, and a site picks its block through rate_cat[site].
struct DtabArgs {
  double       *dtab;        // [count][parts_in_class][row]
  const double *models;      // class models
  const int    *cls_part_global;   // [parts_in_class] -> global partition index (branch slot)
  const double *lzq, *lzr;   // [count][nbranches]
  const double *cat_rates;   // CAT: [parts_in_class][max_cats] perSiteRates; null under GAMMA
  int count, nparts_cls, nbranches, states, model_doubles, row, max_cats;
  int eign_cat_stride;       // doubles between the EIGN of consecutive categories (0: shared, DNA; 20: AA)
  int rates_off, lg4_off;    // offsets of gammaRates[4] / the LG4 flag inside a model block (lg4_off < 0: none)
};

// Everything a traversal kernel needs.  Passed by value (fits the 4 KB parameter space).
struct TraverseArgs {
  // static layout of this class
  const Tile   *tiles;
  double       *clv;        // [(mxtips-2)][N][span]   NOT const/restrict: read-after-write by the same thread
  const uint8_t*tips;       // [mxtips][N]
  const int    *wgt;        // [N]
  double       *sum;        // [N][span]
  const double *models;     // [parts_in_class][model doubles]
  size_t        N;          // padded sites on the class axis
  int           mxtips;
  int           nparts;     // global partition count P
  int           nbranches;
  // per call
  const DevEntry      *entries;   // [count]
  const double        *lzq;       // [count][nbranches]  log(max(z, zmin)) of the q branch (entry 0: the root branch)
  const double        *lzr;       // [count][nbranches]
  const unsigned char *mask;      // [P] executeModel
  int           count;
  int           first;      // first newview entry (0 for newview, 1 when entry 0 is a branch)
  int           mode;       // Mode
  unsigned int *scale_inc;  // [maxEntries][P]  per-entry scaler increments (integer atomics: deterministic)
  double       *tile_out;   // [ntiles][2]      per-tile partial sums (evaluate: lnL; core: d1,d2)
  const double *dtab;       // [count][nparts_cls][row]
  const double *lut;        // AA only, may be null: [count][nparts_cls][2 sides][23 codes][4 | max_cats][20] = P(t) * tipVector[code]
  const int    *rate_cat;   // CAT: [N] rate category of every site; null under GAMMA
  int           max_cats;   // CAT: categories per partition in dtab / lut rows
  int           dtab_row;   // doubles per (entry, partition) in dtab
  int           nparts_cls;
  int           tile_sites; // sites per tile (multiple of 32)
  // -S (SEV, subtree equality vectors at tile granularity; null / 0 when off): a (node, tile) pair whose sites are all
  // gaps in the whole subtree below the node owns no CLV block -- every such site equals the node's "gap column"
  const double *const *sev_table;   // [mxtips-2][sev_ntiles] block of the (node, tile) pair; null = all-gap
  const uint8_t *sev_tipgap;        // [mxtips][sev_ntiles]   1 = all sites of the tile are gaps in this taxon (unused by the kernels: tips always go through their codes)
  const double *sev_gapcol;         // [mxtips-2][nparts_cls][span] CLV of an all-gap site at every inner node
  int           sev_ntiles;         // tiles of the class
  int           sev_tile0;          // class-wide index of this launch's tile 0 (also indexes score_out)
  // exb_score_insertions (MODE_SCORE; null / 0 otherwise)
  long long     scratch_off;        // element offset of scratch node 0 relative to clv (node n: + (n - (2*mxtips-1)) * N * span)
  double2      *score_out;          // [ncand][score_stride] per (candidate, tile, warp): {sum of wgt*log, summed weights of scaled sites}
  int           score_stride;       // tiles of the class x warps per CTA
};

// DNA constants of a class that holds a single partition: passed by value so that every EI/EV
// element is an immediate constant-bank operand of its DFMA (no load instruction at all).
struct DnaConsts { double EI[16]; double EV[16]; };

struct CoreArgs {
  const Tile   *tiles;
  const double *sum;
  const int    *wgt;
  const double *models;
  const int    *rate_cat;    // CAT: [N]; null under GAMMA
  const double *cat_rates;   // CAT: [parts_in_class][max_cats]
  int           max_cats;
  const double *lz;          // [nbranches] trial log z
  const unsigned char *mask;
  int           nbranches;
  double       *tile_out;    // [ntiles][2]
};

constexpr int kMaxP2pRanks = 16;
// several GPUs on one node: the all-reduce is done by the kernels themselves over CUDA-IPC mapped peer buffers (p2p.cuh)
struct P2pArgs {
  int           nranks, rank;                    // nranks <= 1: no exchange
  double       *peer_buf[kMaxP2pRanks];          // receive buffers: [2 slots][nranks][stride] doubles (entry `rank` is mine)
  unsigned int *peer_flags[kMaxP2pRanks];        // [nranks] sequence flags in each receive buffer
};
struct FinishArgs {
  // scaler fix-up: gs[p] = gs[q] + gs[r] + inc  for every newview entry, in order
  unsigned int *gs;          // [P][gs_stride], gs_stride = 2*mxtips + kScratchNodes
  int           gs_stride;
  unsigned int *scale_inc;   // [maxEntries][P] (re-zeroed here)
  const DevEntry *entries;
  int           count, first;
  int           mxtips, nparts;
  const unsigned char *mask;
  const int    *part_has_sites;   // [P] local width > 0
  // reduction of tile partials
  const double *tile_out[2];      // per class
  const int    *part_tile_begin;  // [P] first tile of the partition inside its class list
  const int    *part_tile_count;  // [P]
  const int    *part_cls;         // [P] class of the partition
  int           mode;
  double        log_minlikelihood;
  double       *out;              // evaluate: [P]; core: [2][P]
  // single-rank fast path: results and a completion flag go straight to host-mapped pinned memory
  double       *out_host;         // may be null
  volatile unsigned int *done_flag_host;   // may be null
  unsigned int *ticket;           // block arrival counter (self-resetting)
  unsigned int  seq;              // value to publish in *done_flag_host
  int           gs_in_smem;       // scaler chain staged in shared memory (set by launch_finish)
  // several GPUs on one node: the all-reduce of evaluateGeneric / topLevelMakenewz is part of this kernel (p2p.cuh);
  // receive-buffer rows are 2*P doubles
  P2pArgs       x;
  unsigned int  xseq;             // exchange number (slot = xseq & 1)
};

// Short descriptors on a small alignment (the thousands of partial traversals of a tree search): the whole call travels
// in the kernel parameters -- descriptor, execute mask, the host-built d-table rows -- so no copy precedes the launch, and
// the last CTA of the (last) traverse launch does the finish work itself (finish.cuh): one launch per call.
constexpr int kSmallEntries = 8, kSmallParts = 8, kSmallDtabDoubles = 512, kSmallMaxTiles = 64;
struct SmallCall {
  int count, fold;                             // entries incl. the branch entry | this launch ends the call
  DevEntry entries[kSmallEntries];
  unsigned char mask[kSmallParts];
  double dtab[kSmallDtabDoubles];              // [count][nparts_cls][32] as stage_traversal lays it out
  FinishArgs fin;                              // entries / mask inside are ignored (the kernel's own copies are used)
};
struct NoSmallCall {};

// AA tip look-up table: what precomputeTips / umpX do in the reference (avxLikelihood.c:1353-1389, mic_native_aa.c:97)
constexpr int kAaLutRow = 23 * 80;
struct LutArgs {
  double *lut; const double *models; const double *dtab; const DevEntry *entries;
  int count, first, nparts_cls;
  int ncat;        // categories per table row: 4 (GAMMA) or max_cats (CAT)
  int dtab_row;
  int cat;         // CAT: every category uses the (single) eigen-system 0
};
void launch_aa_lut(const LutArgs &a, cudaStream_t s);
// ---- device-resident Newton-Raphson (kernels_nr.cu) ----
constexpr int kMaxBranchSlots = 256;     // NUM_BRANCHES, axml.h:134
struct NrState {
  double z[kMaxBranchSlots], z0[kMaxBranchSlots], zprev[kMaxBranchSlots], zstep[kMaxBranchSlots], lz[kMaxBranchSlots];
  double d1[kMaxBranchSlots], d2[kMaxBranchSlots];
  int budget[kMaxBranchSlots];
  unsigned char done[kMaxBranchSlots], curvatOK[kMaxBranchSlots];
  int all_done, iterations;
  unsigned int xseq;        // multi-rank: number of the last exchange used
  int peer_lost;            // multi-rank: a peer did not deliver within the watchdog time
};
constexpr int kNrInitSlots = 8;     // start values travel as kernel parameters up to this many branch slots / partitions
struct NrArgs {
  NrState *state;
  unsigned char *mask;            // [P] executeModel, updated by the state machine
  unsigned int *barrier;          // monotonic arrival counter (zeroed before a launch of more than one CTA)
  const Tile *tiles[2]; int ntiles[2];
  const double *sum[2]; const int *wgt[2]; const double *models[2];
  const int *rate_cat[2]; const double *cat_rates[2]; int max_cats;     // CAT (null / 0 under GAMMA)
  double *tile_out[2];
  double *part_out;               // [2][P]
  const int *part_has_sites, *part_tile_begin, *part_tile_count, *part_cls;
  int nparts, nbranches;
  // several GPUs on one node: one exchange per iteration inside the kernel (p2p.cuh); the first one is number xseq0 + 1,
  // the last one used is left in state->xseq
  P2pArgs x;
  unsigned int xseq0;
  // start-up without copies (nbranches, nparts <= kNrInitSlots): z0 / maxiter / executeModel as parameters instead of a
  // NrState and a mask uploaded beforehand
  int init_from_args, maxiter;
  double z0[kNrInitSlots];
  unsigned char mask0[kNrInitSlots];
  // results straight into pinned host memory + completion flag (the host polls; no device-to-host copy, no stream sync)
  NrState *result_host; volatile unsigned int *done_flag_host; unsigned int seq;
  // newton_raphson_small_kernel (one CTA, everything on chip): where each partition's rows sit on its class's site axis
  unsigned int small_site0[kNrInitSlots], small_width[kNrInitSlots];
  int small_cls[kNrInitSlots], small_pc[kNrInitSlots];        // class and class-partition index of every partition
};
constexpr size_t kNrStagedBytes = 192 * 1024;    // one-CTA Newton-Raphson: the sum table (+ weights) of all tiles in shared memory up to this size
int launch_newton_raphson(const NrArgs &a, int blocks, size_t staged_bytes, cudaStream_t s);   // staged_bytes > 0: the one-CTA kernel
int newton_raphson_max_blocks(int device);

// ---- per-site likelihood at an explicit rate (kernels_partial.cu; evaluatePartialGeneric) ----
constexpr int kPartialThreads = 128;
struct PartialPair  { double rate; uint32_t site; uint32_t pad; };        // site on the class axis
struct PartialChunk { uint32_t begin, count; int part_cls, branch_slot; };  // <= kPartialThreads pairs of ONE partition
struct PartialEntry { int tipCase, p, q, r; };                             // tips: node number; inner nodes: slot index
struct PartialArgs {
  const PartialChunk *chunks; int nchunks;
  const PartialPair  *pairs;
  const PartialEntry *entries; int count;
  const double *lzq, *lzr;     // [count][nbranches] log(max(z, zmin))
  const double *root_lz;       // [nbranches] log(z) of the root branch, NOT clamped (evaluatePartialGenericSpecial.c:232-234)
  int nbranches;
  const double *models; int model_doubles;
  const uint8_t *tips; size_t N;
  const int *wgt;
  double *scratch; int nslots;   // [grid][nslots][S][kPartialThreads]
  double log_minlikelihood;
  double *out;                   // [npairs], in pair order
};
void launch_partial(const PartialArgs &a, int states, int blocks, cudaStream_t s);
int partial_max_blocks(int states, int device);

void launch_dtab(const DtabArgs &a, cudaStream_t s);
void launch_traverse_dna(const TraverseArgs &a, const DnaConsts *single_partition_consts, int ntiles, cudaStream_t s,
                         const SmallCall *small = nullptr);
void launch_traverse_dna_cat(const TraverseArgs &a, const DnaConsts *single_partition_consts, int ntiles, cudaStream_t s);
void launch_core_dna(const CoreArgs &a, int ntiles, cudaStream_t s);
void launch_traverse_aa(const TraverseArgs &a, bool any_lg4, int ntiles, cudaStream_t s);
void launch_traverse_aa_tc(const TraverseArgs &a, int ntiles, cudaStream_t s);
void launch_core_aa(const CoreArgs &a, int ntiles, cudaStream_t s);
void launch_finish(const FinishArgs &a, cudaStream_t s);

// exb_score_insertions: scaler chain of the mixed descriptor (one block per partition), then the fixed-order reduction of
// the per-(candidate, tile, warp) partials into lnl[candidate][partition]
#ifndef EXB_AA_WARPS
#define EXB_AA_WARPS 12
#endif
constexpr int kDnaWarps = 8, kAaWarps = EXB_AA_WARPS;   // warps per CTA of the GAMMA traverse kernels (slots per tile in score_out)
constexpr int kAaChunkSites = 8 * kAaWarps, kAaCatChunkSites = 32 * kAaWarps;   // sites per chunk of the AA kernel (GAMMA: 8 per warp; PSR: 32)
struct ScoreFinishArgs {
  unsigned int *gs; int gs_stride, gs_in_smem;
  unsigned int *scale_inc;
  const DevEntry *entries; int count, mxtips, nparts, ncand;
  const unsigned char *mask; const int *part_has_sites;
  unsigned int *sc;                  // [P][ncand] gs[q] + gs[r] + gs[s] at the time of the scoring entry
  const double2 *score_out[2]; int score_stride[2], warps[2];
  const int *part_tile_begin, *part_tile_count, *part_cls;
  double log_minlikelihood;
  double *out;                       // [ncand][P]
};
void launch_score_finish(const ScoreFinishArgs &a, cudaStream_t s);
void launch_p2p_warmup(const P2pArgs &x, unsigned int xseq, int stride, cudaStream_t s);

// -S: gap columns of the inner nodes of a descriptor, in entry order (kernels_sev.cu)
struct GapcolArgs {
  double *gapcol;              // [mxtips-2][nparts_cls][span]
  const DevEntry *entries; int first, count, mxtips;
  const double *dtab; int dtab_row, nparts_cls;
  const double *models; int model_doubles, states;
  const int *gapw;             // [count][nparts_cls] summed pattern weights of the partition's all-gap tiles at entry's node
  const int *cls_part_global;  // [nparts_cls] -> global partition
  const unsigned char *mask;   // [P]
  unsigned int *scale_inc; int nparts;
};
void launch_gapcol(const GapcolArgs &a, cudaStream_t s);

// Dynamic shared memory above 48 KB is an opt-in per (kernel, device): remember what each device was given.
constexpr int kMaxDevices = 64;
template <class Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t (&done)[kMaxDevices])
{
  int dev = 0;
  cudaError_t r = cudaGetDevice(&dev);
  if (r != cudaSuccess) return r;
  size_t &have = done[dev >= 0 && dev < kMaxDevices ? dev : 0];
  if (bytes <= have && dev < kMaxDevices) return cudaSuccess;
  r = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (r == cudaSuccess) have = bytes;
  return r;      // a failure also surfaces as a launch error, which the engine reports
}

// ---- device helpers shared by the kernels --------------------------------------------------------------
#ifdef __CUDACC__
constexpr int kThreads = 256;

// Deterministic block-wide sum: fixed shuffle tree inside each warp, then warp 0 adds the warp
// totals in warp order.  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *smem32)
{
  #pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // smem32 may still be in use by a previous reduction
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < (int)(blockDim.x >> 5)) ? smem32[lane] : 0.0;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}
#endif

}  // namespace exb
