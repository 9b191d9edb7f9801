// engine.cu -- host side of the B200 likelihood engine and its C ABI (include/examl_b200.h).
//
// Owns the HBM-resident state that the reference keeps in pInfo (axml.h:533-629): CLVs of all
// inner nodes (xVector), tip codes (yVector), pattern weights (wgt), scaler counts (globalScaler),
// the makenewz sum table (sumBuffer) -- and turns each *Iterative call of the reference into
// one fused kernel launch per data-type class plus one small "finish" launch.
#include "../../include/examl_b200.h"
#include "engine.cuh"
#include "p2p.cuh"
#include "finish.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>
#include <vector>

namespace exb {

// ------------------------------------------------------------------------------------------------------
// finish kernel (finish.cuh): scaler bookkeeping, per-partition reduction, scaling undo, results to the host.
// One block per partition.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
finish_kernel(const FinishArgs a)
{
  extern __shared__ unsigned int gsm[];            // one scaler row (2*mxtips) when it fits
  __shared__ double red0[4], red1[4];
  finish_body(a, a.entries, a.mask, gsm, red0, red1, (int)blockIdx.x, (int)gridDim.x);
  if (a.done_flag_host) {
    // the last block to arrive publishes the completion flag after everybody's results are visible to the host
    __shared__ int last, peer_lost;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    finish_publish(a, &peer_lost);
  }
}

// first touch of every peer mapping: an all-reduce of one zero
__global__ void p2p_warmup_kernel(const P2pArgs x, unsigned int xseq, int stride)
{
  if ((int)threadIdx.x < x.nranks) x.peer_buf[threadIdx.x][p2p_row(x, xseq, x.rank, stride)] = 0.0;
  p2p_finish_exchange(x, xseq, 1, stride, nullptr, nullptr);
}
void launch_p2p_warmup(const P2pArgs &x, unsigned int xseq, int stride, cudaStream_t s)
{
  p2p_warmup_kernel<<<1, 32, 0, s>>>(x, xseq, stride);
}


// ------------------------------------------------------------------------------------------------------
// exb_score_insertions, after the traverse kernels.
// (1) score_chain_kernel, one warp per partition: the scaler bookkeeping of the mixed descriptor in entry order --
//     newview entries gs[p] = gs[q] + gs[r] + inc as in finish_kernel (scratch nodes have slots behind the tree's own),
//     scoring entries record sc[candidate] = gs[q] + gs[r] + gs[s] as it stands at that point of the walk.
// (2) score_reduce_kernel, one block per (candidate, partition): the per-(tile, warp) partials in fixed order, then
//     lnl = sum + (scaled-site weights + sc) * log(minlikelihood)   (evaluateGenericSpecial.c:830).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
score_chain_kernel(const ScoreFinishArgs a)
{
  extern __shared__ unsigned int gsm[];            // one scaler row (gs_stride) when it fits, else the row is used in place
  const int lane = threadIdx.x;
  const int s_node = a.entries[0].p;
  for (int part = blockIdx.x; part < a.nparts; part += gridDim.x) {
    unsigned int *sc = a.sc + (size_t)part * a.ncand;
    if (!(a.mask[part] && a.part_has_sites[part])) {
      for (int k = lane; k < a.ncand; k += 32) sc[k] = 0;
      continue;
    }
    unsigned int *row = a.gs + (size_t)part * a.gs_stride;
    const bool staged = a.gs_in_smem;
    unsigned int *gs = staged ? gsm : row;
    if (staged) { for (int i = lane; i < a.gs_stride; i += 32) gs[i] = row[i]; }
    __syncwarp();
    for (int e0 = 1; e0 < a.count; e0 += 32) {
      const int e = e0 + lane;
      unsigned int inc = 0; DevEntry en = {0, 0, 0, 0};
      if (e < a.count) {
        en = a.entries[e];
        if (!(en.tipCase & kScoreBit)) {
          inc = a.scale_inc[(size_t)e * a.nparts + part];
          if (inc) a.scale_inc[(size_t)e * a.nparts + part] = 0;
        }
      }
      const int n = min(32, a.count - e0);
      for (int j = 0; j < n; j++) {                 // lane j's entry, in order
        const int tj = __shfl_sync(0xffffffffu, en.tipCase, j), pj = __shfl_sync(0xffffffffu, en.p, j),
                  qj = __shfl_sync(0xffffffffu, en.q, j), rj = __shfl_sync(0xffffffffu, en.r, j);
        const unsigned int ij = __shfl_sync(0xffffffffu, inc, j);
        if (lane == 0) {
          if (tj & kScoreBit) sc[-pj - 1] = gs[qj] + gs[rj] + gs[s_node];
          else gs[pj] = gs[qj] + gs[rj] + ij;
        }
        __syncwarp();
      }
    }
    if (staged) { for (int i = lane; i < a.gs_stride; i += 32) row[i] = gs[i]; }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128)
score_reduce_kernel(const ScoreFinishArgs a)
{
  __shared__ double red0[4], red1[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.x;
  for (int part = blockIdx.y; part < a.nparts; part += gridDim.y) {
    const bool exec = a.mask[part] && a.part_has_sites[part];
    double s0 = 0.0, s1 = 0.0;
    if (exec) {
      const int c = a.part_cls[part];
      const double2 *v = a.score_out[c] + (size_t)k * a.score_stride[c] + (size_t)a.part_tile_begin[part] * a.warps[c];
      const int n = a.part_tile_count[part] * a.warps[c];
      for (int t = tid; t < n; t += 128) { const double2 x = v[t]; s0 += x.x; s1 += x.y; }
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
      }
    }
    if (lane == 0) { red0[warp] = s0; red1[warp] = s1; }
    __syncthreads();
    if (tid == 0) {
      s0 = ((red0[0] + red0[1]) + (red0[2] + red0[3]));
      s1 = ((red1[0] + red1[1]) + (red1[2] + red1[3]));
      a.out[(size_t)k * a.nparts + part] = exec ? s0 + (s1 + (double)a.sc[(size_t)part * a.ncand + k]) * a.log_minlikelihood : 0.0;
    }
    __syncthreads();
  }
}

void launch_score_finish(const ScoreFinishArgs &a0, cudaStream_t s)
{
  ScoreFinishArgs a = a0;
  const size_t smem = (size_t)a.gs_stride * sizeof(unsigned int);
  const int blocks = a.nparts < 592 ? a.nparts : 592;
  a.gs_in_smem = smem <= 40 * 1024;               // else the scaler row is walked in place
  // (a piece of a scan that holds newview entries only still needs its scaler bookkeeping)
  if (a.count > 1) score_chain_kernel<<<blocks, 32, a.gs_in_smem ? smem : 0, s>>>(a);
  if (a.ncand <= 0) return;
  dim3 grid((unsigned)a.ncand, (unsigned)(a.nparts < 4096 ? a.nparts : 4096));
  score_reduce_kernel<<<grid, 128, 0, s>>>(a);
}

void launch_finish(const FinishArgs &a0, cudaStream_t s)
{
  FinishArgs a = a0;
  int blocks = a.nparts;                           // one block per partition, grid-stride beyond 592
  if (blocks > 592) blocks = 592;
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)2 * a.mxtips * sizeof(unsigned int);          // one scaler row per block
  a.gs_in_smem = smem <= 40 * 1024;
  finish_kernel<<<blocks, 128, a.gs_in_smem ? smem : 0, s>>>(a);
}

}  // namespace exb

using namespace exb;

// ------------------------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CUDA_TRY(x)                                                                              \
  do {                                                                                           \
    cudaError_t err__ = (x);                                                                     \
    if (err__ != cudaSuccess)                                                                    \
      return fail(EXB_ERR_CUDA, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily with dlopen so that (a) a single-GPU process never needs it and (b) inside a
// Python process that already carries torch's bundled libnccl.so.2 the same library is reused.
// ------------------------------------------------------------------------------------------------------
namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclUniqueId { char internal[128]; };
static_assert(sizeof(NcclUniqueId) == EXB_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, NcclUniqueId, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
} g_nccl;
constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;       // ncclSum
constexpr int kNcclInt8 = 0;      // ncclInt8 / ncclChar

int load_nccl()
{
  if (g_nccl.handle) return EXB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl.handle; i++) g_nccl.handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.handle) return fail(EXB_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(NcclUniqueId *))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, NcclUniqueId, int))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.handle, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char *(*)(int))dlsym(g_nccl.handle, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return fail(EXB_ERR_NCCL, "libnccl lacks an expected symbol");
  return EXB_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------------
struct HostPart {
  int states = 0, cls = -1, part_cls = -1;
  size_t width = 0, site0 = 0;
  bool defined = false;
  std::vector<int> wgt;
  std::vector<unsigned char> y;        // [mxtips][width], dropped after commit
  std::vector<double> model;           // host cache of the uploaded constants
  bool model_set = false;
  // CAT / PSR
  int ncat = 0;
  std::vector<double> cat_rates;       // perSiteRates[ncat]
  std::vector<int> rate_category;      // [width]
  bool cats_set = false;
};

// -S (SEV at tile granularity): CLV blocks exist only for (node, tile) pairs that are not all-gap
struct SevClass {
  bool on = false;
  size_t block_doubles = 0;                 // capacity of one block: tile capacity x span
  int ntiles = 0;
  std::vector<uint8_t> tipgap, nodegap;      // [mxtips][ntiles], [mxtips-2][ntiles]: all sites of the tile are gaps (below the node)
  std::vector<double *> table_h;             // [mxtips-2][ntiles] block of the pair, null = none
  double **table_d = nullptr;
  double *gapcol = nullptr;                  // [mxtips-2][nparts][span]
  int *gapw_d = nullptr;                     // [max_entries][nparts] weights of the all-gap tiles per (entry, partition)
  std::vector<int> gapw_h, tile_wgt;
  std::vector<double *> free_blocks;
  std::vector<void *> slabs;
  size_t blocks_in_use = 0;
};

struct ClassState {
  SevClass sev;
  int S = 0, span = 0, codes = 0, model_doubles = 0;
  size_t N = 0;                        // padded number of sites
  int nparts = 0;
  std::vector<Tile> tiles_h;
  Tile *tiles = nullptr;
  double *clv = nullptr, *sum = nullptr, *models = nullptr, *tile_out = nullptr, *dtab = nullptr, *lut = nullptr;
  uint8_t *tips = nullptr;
  int *wgt = nullptr, *part_global = nullptr;   // class-partition -> global partition
  int dtab_row = 0;
  int *rate_cat = nullptr;             // CAT: [N] category of every site
  double *cat_rates = nullptr;         // CAT: [nparts][max_cats]
  std::vector<double> models_h;
  std::vector<int> cp_tile_begin, cp_tile_count, cp_global;   // per class-partition: its run of tiles, global index
  int dirty_lo = 1 << 30, dirty_hi = -1;   // range of class-partitions whose constants must be re-uploaded
  // exb_score_insertions: CLVs of the scratch nodes [scratch_nodes][N][span] and the per-(candidate, tile, warp) partials
  double *scratch = nullptr; int scratch_nodes = 0;
  double2 *score_out = nullptr; size_t score_out_cap = 0;
};

struct exb_engine {
  int device = 0, mxtips = 0, P = 0, nb = 1;
  int gs_stride = 0;                   // entries per partition row of gs
  int tile_sites = 256;
  bool tile_sites_set = false;         // exb_set_tile_sites was called: no automatic balancing
  bool committed = false;
  std::vector<HostPart> parts;
  ClassState cls[2];
  bool cat = false;                    // CAT / PSR: one rate category per site (span = S)
  bool sev = false;                    // -S: no CLV storage for all-gap (node, tile) pairs (exb_set_sev)
  bool aa_tc = false;                  // AA GAMMA newview on tcgen05 in 3xTF32 (mixed precision, exb_set_aa_precision)
  int max_cats = 4;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_a = nullptr, ev_b = nullptr;   // call start / end, traverse start / end
  // per-call staging (pinned host + device mirror): entries | lzq | lzr | mask | lz(core)
  size_t stage_bytes = 0, stage_used_bytes = 0;
  unsigned char *stage_h = nullptr, *stage_d = nullptr;
  size_t off_entries = 0, off_lzq = 0, off_lzr = 0, off_mask = 0, off_corelz = 0;   // capacity layout (exb_core uses the tail)
  // per traversal call: the descriptor is packed compactly and only the bytes in use are copied
  size_t o_entries = 0, o_lzq = 0, o_lzr = 0, o_mask = 0, o_dtab[2] = {0, 0};
  bool host_dtab[2] = {false, false};  // short descriptors: the exponentials are computed on the host and ride along
  int max_entries = 0;
  // partition tables
  int *part_has_sites = nullptr, *part_tile_begin = nullptr, *part_tile_count = nullptr, *part_cls = nullptr;
  unsigned int *gs = nullptr, *scale_inc = nullptr;
  double *out_d = nullptr, *out_h = nullptr;     // [2*P]
  long launches = 0;
  float last_ms = 0.f;
  float phase_ms[3] = {0.f, 0.f, 0.f};   // tables (d-table, tip look-up) | traverse kernels | finish kernel of the last traversal call
  bool phases_valid = false;
  bool timing = true;                    // CUDA events around the kernels of every call (exb_last_kernel_ms)
  unsigned int *ticket = nullptr;        // finish-kernel arrival counter
  volatile unsigned int *done_flag = nullptr;   // pinned, written by the finish kernel on the single-rank path
  unsigned int seq = 0;
  unsigned char *stage_use = nullptr;    // what the kernels of the current call read: stage_d, or stage_h (zero copy)
  int total_tiles = 0;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // all GPUs on one node: the all-reduce is fused into finish_kernel over CUDA-IPC mapped peer buffers (NVLink)
  bool p2p = false;
  unsigned char *xbuf = nullptr;                 // my receive buffer: [2][nranks][2P] doubles, then [nranks] flags
  double *peer_buf[kMaxP2pRanks] = {};
  unsigned int *peer_flags[kMaxP2pRanks] = {};
  unsigned int xseq = 0;
  // device-resident Newton-Raphson
  NrState *nr_state_d = nullptr, *nr_state_h = nullptr;
  unsigned int *nr_barrier = nullptr;
  unsigned char *nr_mask = nullptr;
  int nr_max_blocks = 0;
  // exb_score_insertions: scaler sums per (partition, candidate), results [ncand][P] on the device and in pinned memory
  unsigned int *score_sc = nullptr; size_t score_sc_cap = 0;
  double *score_d = nullptr; size_t score_d_cap = 0;
  double *score_h = nullptr; size_t score_h_cap = 0;
  // exb_evaluate_partial: growable device buffers (descriptor + pairs | results | slot stacks)
  unsigned char *pp_in = nullptr; size_t pp_in_cap = 0;
  double *pp_out = nullptr; size_t pp_out_cap = 0;
  double *pp_scratch = nullptr; size_t pp_scratch_cap = 0;
  int pp_max_blocks[2] = {0, 0};
};

constexpr int kMaxParamPartitions = 16;
constexpr int kNrSingleCtaTiles = 4;             // device Newton-Raphson: up to this many tiles one CTA does the whole loop
// -S: CLV blocks are carved from slabs of ~256 MB that are allocated on demand
static size_t kSevSlabBlocks(size_t block_doubles) { return std::max<size_t>(1, ((size_t)256 << 20) / (block_doubles * sizeof(double))); }
constexpr size_t kHostDtabBytes = 16 * 1024;     // per class
constexpr int kHostDtabMaxExps = 256;            // ~25 ns each on the host: cheaper than a kernel launch below this

static int class_of(int states) { return states == 4 ? CLS_DNA : (states == 20 ? CLS_AA : -1); }

extern "C" const char *exb_last_error(void) { return g_err.c_str(); }

extern "C" int exb_create(exb_engine **out, int device, int mxtips, int numPartitions, int numBranches)
{
  if (!out || mxtips < 3 || numPartitions < 1 || numBranches < 1)
    return fail(EXB_ERR_ARG, "exb_create: bad arguments (mxtips=%d partitions=%d branches=%d)", mxtips, numPartitions, numBranches);
  if (numBranches > 1 && numBranches != numPartitions)
    return fail(EXB_ERR_ARG, "exb_create: numBranches must be 1 or equal to the number of partitions (makenewzGenericSpecial.c:1195)");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0)
    return fail(EXB_ERR_CUDA, "exb_create: no CUDA device available (%s); examl_b200 has no CPU fallback",
                err == cudaSuccess ? "device count is 0" : cudaGetErrorString(err));
  if (device < 0 || device >= ndev) return fail(EXB_ERR_ARG, "exb_create: device %d out of range (0..%d)", device, ndev - 1);
  CUDA_TRY(cudaSetDevice(device));
  exb_engine *e = new exb_engine();
  e->device = device; e->mxtips = mxtips; e->P = numPartitions; e->nb = numBranches;
  e->parts.resize(numPartitions);
  e->cls[CLS_DNA].S = 4;  e->cls[CLS_DNA].span = 16; e->cls[CLS_DNA].codes = 16; e->cls[CLS_DNA].model_doubles = kDnaModelDoubles;
  e->cls[CLS_AA].S = 20;  e->cls[CLS_AA].span = 80;  e->cls[CLS_AA].codes = 23;  e->cls[CLS_AA].model_doubles = kAaModelDoubles;
  e->cls[CLS_DNA].dtab_row = kDnaDtabRow; e->cls[CLS_AA].dtab_row = kAaDtabRow;
  { const char *v = getenv("EXB_AA_PRECISION"); e->aa_tc = v && strcmp(v, "tf32x3") == 0; }
  CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&e->ev0));
  CUDA_TRY(cudaEventCreate(&e->ev1));
  CUDA_TRY(cudaEventCreate(&e->ev_a));
  CUDA_TRY(cudaEventCreate(&e->ev_b));
  *out = e;
  return EXB_OK;
}

extern "C" void exb_destroy(exb_engine *e)
{
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  if (e->p2p)
    for (int r = 0; r < e->nranks; r++)
      if (r != e->rank && e->peer_buf[r]) cudaIpcCloseMemHandle(e->peer_buf[r]);
  cudaFree(e->xbuf);
  if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    if (k.sev.on) {
      for (void *sl : k.sev.slabs) cudaFree(sl);
      k.clv = nullptr;
      cudaFree(k.sev.table_d); cudaFree(k.sev.gapcol); cudaFree(k.sev.gapw_d);
    }
    cudaFree(k.tiles); cudaFree(k.clv); cudaFree(k.sum); cudaFree(k.models); cudaFree(k.tile_out);
    cudaFree(k.tips); cudaFree(k.wgt); cudaFree(k.dtab); cudaFree(k.part_global); cudaFree(k.lut);
    cudaFree(k.rate_cat); cudaFree(k.cat_rates);
    cudaFree(k.scratch); cudaFree(k.score_out);
  }
  cudaFree(e->score_sc); cudaFree(e->score_d); cudaFreeHost(e->score_h);
  cudaFree(e->stage_d); cudaFreeHost(e->stage_h);
  cudaFree(e->part_has_sites); cudaFree(e->part_tile_begin); cudaFree(e->part_tile_count); cudaFree(e->part_cls);
  cudaFree(e->gs); cudaFree(e->scale_inc); cudaFree(e->out_d); cudaFreeHost(e->out_h);
  cudaFree(e->nr_state_d); cudaFreeHost(e->nr_state_h); cudaFree(e->nr_barrier); cudaFree(e->nr_mask);
  cudaFree(e->ticket); cudaFreeHost((void *)e->done_flag);
  cudaFree(e->pp_in); cudaFree(e->pp_out); cudaFree(e->pp_scratch);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->ev_a) cudaEventDestroy(e->ev_a);
  if (e->ev_b) cudaEventDestroy(e->ev_b);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

extern "C" int exb_set_rate_model(exb_engine *e, int rateHetModel, int maxCategories)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_set_rate_model: must be called before exb_commit");
  for (const HostPart &p : e->parts)
    if (p.defined) return fail(EXB_ERR_STATE, "exb_set_rate_model: must be called before exb_set_partition");
  if (rateHetModel == 1) { e->cat = false; e->max_cats = 4; }
  else if (rateHetModel == 0) {
    if (maxCategories < 1 || maxCategories > 64) return fail(EXB_ERR_ARG, "exb_set_rate_model: %d rate categories (1..64 supported)", maxCategories);
    e->cat = true; e->max_cats = maxCategories;
  } else return fail(EXB_ERR_ARG, "exb_set_rate_model: rateHetModel %d is neither CAT (0) nor GAMMA (1)", rateHetModel);
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    k.span = (e->cat ? 1 : 4) * k.S;
    k.dtab_row = (e->cat ? e->max_cats : 4) * 2 * k.S;
  }
  return EXB_OK;
}

extern "C" int exb_set_sev(exb_engine *e, int on)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_set_sev: must be called before exb_commit");
  e->sev = on != 0;
  return EXB_OK;
}

extern "C" int exb_sev_stats(exb_engine *e, size_t *blocks_in_use, size_t *blocks_dense, size_t *bytes_allocated)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "exb_sev_stats: engine not committed");
  size_t used = 0, dense = 0, bytes = 0;
  for (int c = 0; c < 2; c++) {
    const ClassState &k = e->cls[c];
    if (!k.sev.on) continue;
    used += k.sev.blocks_in_use;
    dense += (size_t)(e->mxtips - 2) * k.sev.ntiles;
    bytes += k.sev.slabs.size() * (size_t)kSevSlabBlocks(k.sev.block_doubles) * k.sev.block_doubles * sizeof(double);
  }
  if (blocks_in_use) *blocks_in_use = used;
  if (blocks_dense) *blocks_dense = dense;
  if (bytes_allocated) *bytes_allocated = bytes;
  return EXB_OK;
}

extern "C" int exb_set_aa_precision(exb_engine *e, int mode)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_set_aa_precision: must be called before exb_commit");
  if (mode != EXB_AA_FP64 && mode != EXB_AA_TF32X3) return fail(EXB_ERR_ARG, "exb_set_aa_precision: unknown mode %d", mode);
  e->aa_tc = mode == EXB_AA_TF32X3;
  return EXB_OK;
}

extern "C" int exb_set_categories(exb_engine *e, int model, int numberOfCategories, const double *perSiteRates, const int *rateCategory)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "exb_set_categories: call exb_commit first");
  if (!e->cat) return fail(EXB_ERR_STATE, "exb_set_categories: the engine runs the GAMMA model (exb_set_rate_model)");
  if (model < 0 || model >= e->P) return fail(EXB_ERR_ARG, "exb_set_categories: partition %d out of range", model);
  if (numberOfCategories < 1 || numberOfCategories > e->max_cats)
    return fail(EXB_ERR_ARG, "exb_set_categories: %d categories, engine was set up for at most %d", numberOfCategories, e->max_cats);
  HostPart &p = e->parts[model];
  ClassState &k = e->cls[p.cls];
  for (size_t i = 0; i < p.width; i++)
    if (rateCategory[i] < 0 || rateCategory[i] >= numberOfCategories)
      return fail(EXB_ERR_ARG, "exb_set_categories: site %zu has category %d of %d", i, rateCategory[i], numberOfCategories);
  const bool same_rates = p.cats_set && p.ncat == numberOfCategories &&
                          memcmp(p.cat_rates.data(), perSiteRates, (size_t)numberOfCategories * sizeof(double)) == 0;
  const bool same_cats = p.cats_set && (p.width == 0 || memcmp(p.rate_category.data(), rateCategory, p.width * sizeof(int)) == 0);
  if (same_rates && same_cats) return EXB_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  if (!same_rates) {
    p.ncat = numberOfCategories;
    p.cat_rates.assign(e->max_cats, 0.0);
    memcpy(p.cat_rates.data(), perSiteRates, (size_t)numberOfCategories * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(k.cat_rates + (size_t)p.part_cls * e->max_cats, p.cat_rates.data(), (size_t)e->max_cats * sizeof(double),
                             cudaMemcpyHostToDevice, e->stream));
  }
  if (!same_cats && p.width) {
    p.rate_category.assign(rateCategory, rateCategory + p.width);
    CUDA_TRY(cudaMemcpyAsync(k.rate_cat + p.site0, p.rate_category.data(), p.width * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  }
  p.cats_set = true;
  return EXB_OK;
}

extern "C" int exb_set_tile_sites(exb_engine *e, int sites)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_set_tile_sites: must be called before exb_commit");
  if (sites < 32 || sites % 32) return fail(EXB_ERR_ARG, "exb_set_tile_sites: %d is not a positive multiple of 32", sites);
  e->tile_sites = sites;
  e->tile_sites_set = true;
  return EXB_OK;
}

extern "C" int exb_set_partition(exb_engine *e, int model, int states, size_t width,
                                 const int *wgt, const unsigned char *yResource)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_set_partition: engine already committed");
  if (model < 0 || model >= e->P) return fail(EXB_ERR_ARG, "exb_set_partition: partition %d out of range", model);
  const int c = class_of(states);
  if (c < 0)
    return fail(EXB_ERR_ARG, "exb_set_partition: %d-state data is not implemented on the B200 back-end "
                "(DNA = 4 and AA = 20 under GAMMA are)", states);
  if (width > 0 && (!wgt || !yResource)) return fail(EXB_ERR_ARG, "exb_set_partition: null data for a non-empty partition");
  HostPart &p = e->parts[model];
  p.states = states; p.cls = c; p.width = width; p.defined = true;
  p.wgt.assign(wgt, wgt + width);
  p.y.assign(yResource, yResource + (size_t)e->mxtips * width);
  const unsigned char maxcode = (unsigned char)(e->cls[c].codes - 1);
  for (unsigned char v : p.y)
    if (v > maxcode) return fail(EXB_ERR_ARG, "exb_set_partition: tip code %d out of range for %d-state data", (int)v, states);
  p.model.assign(e->cls[c].model_doubles, 0.0);
  return EXB_OK;
}

extern "C" int exb_commit(exb_engine *e)
{
  if (!e || e->committed) return fail(EXB_ERR_STATE, "exb_commit: already committed");
  CUDA_TRY(cudaSetDevice(e->device));
  for (int m = 0; m < e->P; m++)
    if (!e->parts[m].defined) return fail(EXB_ERR_STATE, "exb_commit: partition %d was never described", m);

  // Descriptor capacity: td[0].ti holds mxtips entries (axml.c:595) + the branch entry; the mixed descriptors of
  // exb_score_insertions (a scratch CLV and a scoring entry per candidate) are longer -- three times that, unless the
  // per-entry tables (d-table rows, branch slots) of a many-partition run would then take more than ~1 GB
  {
    size_t per_entry = sizeof(DevEntry) + (size_t)2 * e->nb * sizeof(double);
    for (const HostPart &p : e->parts) per_entry += (size_t)(p.cls == CLS_AA ? 2 * 23 * 80 + 160 : 32) * (e->cat ? (e->max_cats + 3) / 4 : 1) * sizeof(double) + sizeof(unsigned int);
    e->max_entries = e->mxtips + 1;
    if (per_entry * (size_t)(3 * e->mxtips + 1) <= ((size_t)1 << 30)) e->max_entries = 3 * e->mxtips + 1;
  }
  std::vector<int> has_sites(e->P), tbegin(e->P, 0), tcount(e->P, 0), pcls(e->P, 0);
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    size_t n = 0;
    int idx = 0;
    for (int m = 0; m < e->P; m++) {
      HostPart &p = e->parts[m];
      if (p.cls != c) continue;
      p.part_cls = idx++;
      p.site0 = n;
      pcls[m] = c;
      has_sites[m] = p.width > 0;
      tbegin[m] = (int)k.tiles_h.size();
      // AA tiles are whole 96-site chunks (12 warps x 8 sites)
      // (CAT: 384-site chunks, 12 warps x 32 sites)
      const size_t aachunk = e->cat ? kAaCatChunkSites : (e->aa_tc ? 32 : kAaChunkSites);      // (tcgen05 path: 32 sites x 4 categories = the MMA's M)
      size_t tsz = c == CLS_AA ? (size_t)((e->tile_sites + aachunk - 1) / aachunk) * aachunk : (size_t)e->tile_sites;
      // tcgen05 path: one 32-site chunk per CTA unless told otherwise -- four co-resident CTAs per SM want many small tiles
      if (c == CLS_AA && e->aa_tc && !e->cat && !e->tile_sites_set) tsz = 32;
      // AA GAMMA (one CTA per SM): ONE wave of equal tiles.  With fixed 288-site tiles C4 (200 k sites) is 695 tiles = 4.7
      // waves on 148 SMs, i.e. 15 chunk-times where 14.1 would do; one tile per SM is 14 full chunks + 16 sites whose two
      // warps sit on different schedulers (idle warps of a partial chunk skip the tensor work).  Measured on C4, same box:
      // 31.36 vs 31.85 ms (tools/r2_call14.sh; an L2 prefetch of the next entry's loaded children one entry ahead, tried in
      // the same call, cost 0.3 ms and was dropped).  Not under -S, whose saving is per (node, tile) and wants small tiles.
      if (c == CLS_AA && !e->aa_tc && !e->cat && !e->tile_sites_set && !e->sev) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        const size_t per_sm = (p.width + (size_t)(sms > 0 ? sms : 148) - 1) / (size_t)(sms > 0 ? sms : 148);
        const size_t want = (per_sm + 15) / 16 * 16;             // tip rows are staged with 16-byte granularity
        tsz = std::max<size_t>(want, aachunk);
      }
      // One big DNA partition on a few waves of CTAs (strong scaling: 125 k sites per GPU are 1.65 waves of 256-site
      // tiles): cut it into k full waves of equal tiles instead, so that the last wave is not a third empty.
      if (c == CLS_DNA && !e->cat && !e->tile_sites_set && e->P == 1) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
        const size_t slots = (size_t)2 * (sms > 0 ? sms : 148);           // 2 CTAs per SM at 4 sites per thread
        const size_t waves = (p.width + slots * tsz - 1) / (slots * tsz);
        if (waves >= 1 && waves <= 6) {
          const size_t want = (p.width + waves * slots - 1) / (waves * slots);
          tsz = std::min<size_t>(tsz, std::max<size_t>(144, (want + 15) / 16 * 16));     // > 128 keeps 4 sites per thread
        }
      }
      for (size_t s = 0; s < p.width; s += tsz) {
        Tile t;
        t.site0 = (uint32_t)(n + s);
        t.nsites = (uint32_t)std::min<size_t>(tsz, p.width - s);
        t.part = m; t.part_cls = p.part_cls;
        k.tiles_h.push_back(t);
      }
      tcount[m] = (int)k.tiles_h.size() - tbegin[m];
      k.cp_tile_begin.push_back(tbegin[m]);
      k.cp_tile_count.push_back(tcount[m]);
      k.cp_global.push_back(m);
      n += (p.width + 31) / 32 * 32;            // every partition starts on a 32-site boundary
    }
    k.nparts = idx;
    k.N = n;
    if (n >= ((size_t)1 << 32)) return fail(EXB_ERR_ARG, "exb_commit: more than 2^32 sites per rank in one class");
    if (idx == 0) continue;
    k.models_h.assign((size_t)idx * k.model_doubles, 0.0);
    CUDA_TRY(cudaMalloc(&k.models, k.models_h.size() * sizeof(double)));
    {
      std::vector<int> pg;
      for (int m = 0; m < e->P; m++) if (e->parts[m].cls == c) pg.push_back(m);
      CUDA_TRY(cudaMalloc(&k.part_global, pg.size() * sizeof(int)));
      CUDA_TRY(cudaMemcpy(k.part_global, pg.data(), pg.size() * sizeof(int), cudaMemcpyHostToDevice));
      const size_t dbytes = (size_t)e->max_entries * idx * k.dtab_row * sizeof(double);
      CUDA_TRY(cudaMalloc(&k.dtab, dbytes));
      if (e->cat) {
        CUDA_TRY(cudaMalloc(&k.cat_rates, (size_t)idx * e->max_cats * sizeof(double)));
        CUDA_TRY(cudaMemset(k.cat_rates, 0, (size_t)idx * e->max_cats * sizeof(double)));
      }
      // AA tip look-up table (optional: skipped when it would not fit comfortably, e.g. thousands of AA partitions)
      const size_t lbytes = (size_t)e->max_entries * idx * 2 * 23 * 20 * (e->cat ? e->max_cats : 4) * sizeof(double);
      if (c == CLS_AA && lbytes <= ((size_t)1 << 30) && !getenv("EXB_AA_NO_LUT")) CUDA_TRY(cudaMalloc(&k.lut, lbytes));
    }
    if (n == 0) continue;
    // host images of tips / weights on the class site axis (padding: gap code, weight 0)
    std::vector<uint8_t> tips((size_t)e->mxtips * n, (uint8_t)(k.codes - 1));
    std::vector<int> wgt(n, 0);
    for (int m = 0; m < e->P; m++) {
      HostPart &p = e->parts[m];
      if (p.cls != c || p.width == 0) continue;
      memcpy(&wgt[p.site0], p.wgt.data(), p.width * sizeof(int));
      for (int t = 0; t < e->mxtips; t++)
        memcpy(&tips[(size_t)t * n + p.site0], &p.y[(size_t)t * p.width], p.width);
      std::vector<unsigned char>().swap(p.y);
    }
    CUDA_TRY(cudaMalloc(&k.tips, tips.size()));
    CUDA_TRY(cudaMemcpy(k.tips, tips.data(), tips.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&k.wgt, n * sizeof(int)));
    CUDA_TRY(cudaMemcpy(k.wgt, wgt.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    if (e->cat) {
      CUDA_TRY(cudaMalloc(&k.rate_cat, n * sizeof(int)));
      CUDA_TRY(cudaMemset(k.rate_cat, 0, n * sizeof(int)));
    }
    if (e->sev && !e->cat && !(c == CLS_AA && e->aa_tc)) {
      // -S: no dense arena.  Which (taxon, tile) pairs are all-gap is fixed by the alignment; blocks are handed out per call.
      SevClass &v = k.sev;
      v.on = true;
      v.ntiles = (int)k.tiles_h.size();
      size_t cap = 0;
      for (const Tile &t : k.tiles_h) cap = std::max<size_t>(cap, t.nsites);
      v.block_doubles = cap * k.span;
      v.tipgap.assign((size_t)e->mxtips * v.ntiles, 1);
      v.nodegap.assign((size_t)(e->mxtips - 2) * v.ntiles, 0);
      v.table_h.assign((size_t)(e->mxtips - 2) * v.ntiles, nullptr);
      v.tile_wgt.assign(v.ntiles, 0);
      const uint8_t gapcode = (uint8_t)(k.codes - 1);
      for (int t = 0; t < v.ntiles; t++) {
        const Tile &tl = k.tiles_h[t];
        for (uint32_t i = 0; i < tl.nsites; i++) v.tile_wgt[t] += wgt[tl.site0 + i];
        for (int tip = 0; tip < e->mxtips; tip++) {
          const uint8_t *row = &tips[(size_t)tip * n + tl.site0];
          for (uint32_t i = 0; i < tl.nsites; i++) if (row[i] != gapcode) { v.tipgap[(size_t)tip * v.ntiles + t] = 0; break; }
        }
      }
      CUDA_TRY(cudaMalloc(&v.table_d, v.table_h.size() * sizeof(double *)));
      CUDA_TRY(cudaMemset(v.table_d, 0, v.table_h.size() * sizeof(double *)));
      CUDA_TRY(cudaMalloc(&v.gapcol, (size_t)(e->mxtips - 2) * k.nparts * k.span * sizeof(double)));
      CUDA_TRY(cudaMemset(v.gapcol, 0, (size_t)(e->mxtips - 2) * k.nparts * k.span * sizeof(double)));
      v.gapw_h.assign((size_t)e->max_entries * k.nparts, 0);
      CUDA_TRY(cudaMalloc(&v.gapw_d, v.gapw_h.size() * sizeof(int)));
      CUDA_TRY(cudaMemset(v.gapw_d, 0, v.gapw_h.size() * sizeof(int)));
      void *slab = nullptr;                        // the first slab doubles as the base all block offsets are relative to
      const size_t nblk = kSevSlabBlocks(v.block_doubles);
      if (cudaMalloc(&slab, nblk * v.block_doubles * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        return fail(EXB_ERR_CUDA, "exb_commit: cannot allocate the first CLV slab (-S)");
      }
      v.slabs.push_back(slab);
      for (size_t b = nblk; b-- > 0;) v.free_blocks.push_back((double *)slab + b * v.block_doubles);
      k.clv = (double *)slab;
    } else {
    const size_t clv_bytes = (size_t)(e->mxtips - 2) * n * k.span * sizeof(double);
    if (cudaMalloc(&k.clv, clv_bytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(EXB_ERR_CUDA, "exb_commit: cannot allocate %.2f GB of HBM for the CLVs of %zu %d-state sites x %d inner nodes",
                  clv_bytes / 1e9, n, k.S, e->mxtips - 2);
    }
    CUDA_TRY(cudaMemset(k.clv, 0, clv_bytes));
    }
    CUDA_TRY(cudaMalloc(&k.sum, n * k.span * sizeof(double)));
    CUDA_TRY(cudaMemset(k.sum, 0, n * k.span * sizeof(double)));
    CUDA_TRY(cudaMalloc(&k.tiles, k.tiles_h.size() * sizeof(Tile)));
    CUDA_TRY(cudaMemcpy(k.tiles, k.tiles_h.data(), k.tiles_h.size() * sizeof(Tile), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&k.tile_out, k.tiles_h.size() * 2 * sizeof(double)));
    CUDA_TRY(cudaMemset(k.tile_out, 0, k.tiles_h.size() * 2 * sizeof(double)));
  }

  auto upload_ints = [&](int **dst, const std::vector<int> &src) -> cudaError_t {
    cudaError_t r = cudaMalloc(dst, src.size() * sizeof(int));
    if (r != cudaSuccess) return r;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(int), cudaMemcpyHostToDevice);
  };
  CUDA_TRY(upload_ints(&e->part_has_sites, has_sites));
  CUDA_TRY(upload_ints(&e->part_tile_begin, tbegin));
  CUDA_TRY(upload_ints(&e->part_tile_count, tcount));
  CUDA_TRY(upload_ints(&e->part_cls, pcls));

  e->gs_stride = 2 * e->mxtips + kScratchNodes;   // globalScaler has 2*mxtips entries (axml.c:1966-1971); + the scratch nodes of exb_score_insertions
  CUDA_TRY(cudaMalloc(&e->gs, (size_t)e->P * e->gs_stride * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(e->gs, 0, (size_t)e->P * e->gs_stride * sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&e->scale_inc, (size_t)e->max_entries * e->P * sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(e->scale_inc, 0, (size_t)e->max_entries * e->P * sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&e->out_d, (size_t)2 * e->P * sizeof(double)));
  CUDA_TRY(cudaMemset(e->out_d, 0, (size_t)2 * e->P * sizeof(double)));
  CUDA_TRY(cudaMallocHost(&e->out_h, (size_t)2 * e->P * sizeof(double)));

  CUDA_TRY(cudaMalloc(&e->ticket, sizeof(unsigned int)));
  CUDA_TRY(cudaMemset(e->ticket, 0, sizeof(unsigned int)));
  CUDA_TRY(cudaMallocHost((void **)&e->done_flag, sizeof(unsigned int)));
  *e->done_flag = 0;
  e->total_tiles = (int)(e->cls[0].tiles_h.size() + e->cls[1].tiles_h.size());
  CUDA_TRY(cudaMalloc(&e->nr_state_d, sizeof(NrState)));
  CUDA_TRY(cudaMallocHost(&e->nr_state_h, sizeof(NrState)));
  CUDA_TRY(cudaMalloc(&e->nr_barrier, sizeof(unsigned int)));
  CUDA_TRY(cudaMalloc(&e->nr_mask, (size_t)e->P));
  e->nr_max_blocks = newton_raphson_max_blocks(e->device);

  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t off = 0;
  e->off_entries = off; off = align16(off + (size_t)e->max_entries * sizeof(DevEntry));
  e->off_lzq = off;     off = align16(off + (size_t)e->max_entries * e->nb * sizeof(double));
  e->off_lzr = off;     off = align16(off + (size_t)e->max_entries * e->nb * sizeof(double));
  e->off_corelz = off;  off = align16(off + (size_t)e->nb * sizeof(double));
  e->off_mask = off;    off = align16(off + (size_t)e->P);
  off += 2 * (kHostDtabBytes + 128);              // room for host-built d-tables of short descriptors
  e->stage_bytes = off;
  CUDA_TRY(cudaMallocHost(&e->stage_h, off));
  CUDA_TRY(cudaMalloc(&e->stage_d, off));
  e->committed = true;
  return EXB_OK;
}

// Store a partition's constants in its class layout and mark them for upload when they changed.
static int commit_model(exb_engine *e, HostPart &p, ClassState &k, const std::vector<double> &m)
{
  if (p.model_set && memcmp(m.data(), p.model.data(), m.size() * sizeof(double)) == 0) return EXB_OK;
  p.model = m;
  p.model_set = true;
  memcpy(&k.models_h[(size_t)p.part_cls * k.model_doubles], m.data(), m.size() * sizeof(double));
  k.dirty_lo = std::min(k.dirty_lo, p.part_cls);
  k.dirty_hi = std::max(k.dirty_hi, p.part_cls);
  return EXB_OK;
}

extern "C" int exb_set_model(exb_engine *e, int model, const double *EIGN, const double *EV,
                             const double *EI, const double *tipVector, const double *gammaRates)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "exb_set_model: call exb_commit first");
  if (model < 0 || model >= e->P) return fail(EXB_ERR_ARG, "exb_set_model: partition %d out of range", model);
  HostPart &p = e->parts[model];
  ClassState &k = e->cls[p.cls];
  const int S = k.S;
  std::vector<double> m((size_t)k.model_doubles, 0.0);
  if (p.cls == CLS_DNA) {
    double *w = m.data();
    memcpy(w, EIGN, S * sizeof(double)); w += S;
    memcpy(w, gammaRates, 4 * sizeof(double)); w += 4;
    memcpy(w, EI, (size_t)S * S * sizeof(double)); w += S * S;
    memcpy(w, EV, (size_t)S * S * sizeof(double)); w += S * S;
    memcpy(w, tipVector, (size_t)k.codes * S * sizeof(double));
  } else {                                           // AA: the same eigen-system for all four categories
    for (int c = 0; c < 4; c++) {
      memcpy(&m[kAaEign + c * 20], EIGN, 20 * sizeof(double));
      memcpy(&m[kAaEI + c * 400], EI, 400 * sizeof(double));
      memcpy(&m[kAaEV + c * 400], EV, 400 * sizeof(double));
      memcpy(&m[kAaTip + c * 460], tipVector, 460 * sizeof(double));
      m[kAaWeights + c] = 1.0;
    }
    memcpy(&m[kAaRates], gammaRates, 4 * sizeof(double));
    m[kAaLg4] = 0.0;
  }
  return commit_model(e, p, k, m);
}

extern "C" int exb_set_model_lg4(exb_engine *e, int model, const double *EIGN, const double *EV, const double *EI,
                                 const double *tipVector, const double *gammaRates, const double *weights)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "exb_set_model_lg4: call exb_commit first");
  if (model < 0 || model >= e->P) return fail(EXB_ERR_ARG, "exb_set_model_lg4: partition %d out of range", model);
  HostPart &p = e->parts[model];
  if (p.cls != CLS_AA) return fail(EXB_ERR_ARG, "exb_set_model_lg4: partition %d is not protein data", model);
  if (e->cat) return fail(EXB_ERR_ARG, "exb_set_model_lg4: LG4M / LG4X are GAMMA-only models");
  ClassState &k = e->cls[p.cls];
  std::vector<double> m((size_t)k.model_doubles, 0.0);
  memcpy(&m[kAaEign], EIGN, 80 * sizeof(double));
  memcpy(&m[kAaRates], gammaRates, 4 * sizeof(double));
  memcpy(&m[kAaWeights], weights, 4 * sizeof(double));
  m[kAaLg4] = 1.0;
  memcpy(&m[kAaEI], EI, 1600 * sizeof(double));
  memcpy(&m[kAaEV], EV, 1600 * sizeof(double));
  memcpy(&m[kAaTip], tipVector, 1840 * sizeof(double));
  return commit_model(e, p, k, m);
}

// Upload whatever the host changed since the last call.
static int flush_models(exb_engine *e)
{
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    if (k.dirty_hi < k.dirty_lo) continue;
    const size_t off = (size_t)k.dirty_lo * k.model_doubles, cnt = (size_t)(k.dirty_hi - k.dirty_lo + 1) * k.model_doubles;
    // pageable source: the copy is staged by the driver before the call returns, so models_h may change afterwards
    CUDA_TRY(cudaMemcpyAsync(k.models + off, &k.models_h[off], cnt * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    k.dirty_lo = 1 << 30; k.dirty_hi = -1;
  }
  return EXB_OK;
}

static int check_models(exb_engine *e, const unsigned char *mask)
{
  for (int m = 0; m < e->P; m++)
    if ((!mask || mask[m]) && e->parts[m].width > 0 && !e->parts[m].model_set)
      return fail(EXB_ERR_STATE, "partition %d has no model constants (exb_set_model)", m);
  for (int m = 0; m < e->P && e->cat; m++)
    if ((!mask || mask[m]) && e->parts[m].width > 0 && !e->parts[m].cats_set)
      return fail(EXB_ERR_STATE, "partition %d has no rate categories (exb_set_categories)", m);
  return EXB_OK;
}

struct exb_engine;
static void set_finish_outputs(exb_engine *e, FinishArgs &f);

// Validate and stage a traversal descriptor; one H2D copy of exactly the bytes in use.
static int stage_traversal(exb_engine *e, const exb_entry *ti, int count, int first, const double *qz, const double *rz,
                           const unsigned char *mask, int ncand = -1, int *max_scratch = nullptr, bool copy = true)
{
  if (count < first || count > e->max_entries)
    return fail(EXB_ERR_ARG, "traversal of %d entries (max %d)", count, e->max_entries);
  // exb_score_insertions (ncand >= 0): scoring entries, and newviews may use the scratch nodes behind the tree's own
  const int lo = e->mxtips + 1, hi_tree = 2 * e->mxtips - 2, hi = ncand >= 0 ? hi_tree + kScratchNodes : hi_tree;
  int scratch_top = -1;
  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t off = 0;
  e->o_entries = off; off = align16(off + (size_t)count * sizeof(DevEntry));
  e->o_lzq = off;     off = align16(off + (size_t)count * e->nb * sizeof(double));
  e->o_lzr = off;     off = align16(off + (size_t)count * e->nb * sizeof(double));
  e->o_mask = off;    off = align16(off + (size_t)e->P);
  DevEntry *de = (DevEntry *)(e->stage_h + e->o_entries);
  double *lq = (double *)(e->stage_h + e->o_lzq), *lr = (double *)(e->stage_h + e->o_lzr);
  const double logzmin = log(kZmin);
  for (int i = 0; i < count; i++) {
    const exb_entry &t = ti[i];
    if (i < first) {
      if (t.pNumber < 1 || t.pNumber > hi_tree || t.qNumber < 1 || t.qNumber > hi_tree)
        return fail(EXB_ERR_ARG, "branch entry has node numbers %d,%d outside 1..%d", t.pNumber, t.qNumber, hi_tree);
      de[i] = DevEntry{-1, t.pNumber, t.qNumber, 0};
    } else {
      const bool score = ncand >= 0 && t.tipCase >= 0 && (t.tipCase & EXB_SCORE);
      bool ok = score ? (t.pNumber >= 0 && t.pNumber < ncand) : (t.pNumber >= lo && t.pNumber <= hi);
      switch (score ? (t.tipCase & ~EXB_SCORE) : t.tipCase) {
      case EXB_TIP_TIP:     ok = ok && t.qNumber >= 1 && t.qNumber < lo && t.rNumber >= 1 && t.rNumber < lo; break;
      case EXB_TIP_INNER:   ok = ok && t.qNumber >= 1 && t.qNumber < lo && t.rNumber >= lo && t.rNumber <= hi; break;
      case EXB_INNER_INNER: ok = ok && t.qNumber >= lo && t.qNumber <= hi && t.rNumber >= lo && t.rNumber <= hi; break;
      default: ok = false;
      }
      if (!ok) return fail(EXB_ERR_ARG, "traversal entry %d is inconsistent (tipCase=%d p=%d q=%d r=%d, mxtips=%d)",
                           i, t.tipCase, t.pNumber, t.qNumber, t.rNumber, e->mxtips);
      de[i] = score ? DevEntry{(t.tipCase & 3) | kScoreBit, -t.pNumber - 1, t.qNumber, t.rNumber}
                    : DevEntry{t.tipCase, t.pNumber, t.qNumber, t.rNumber};
      if (!score) scratch_top = std::max(scratch_top, t.pNumber - hi_tree - 1);
      scratch_top = std::max(scratch_top, std::max(t.qNumber, t.rNumber) - hi_tree - 1);
    }
    for (int b = 0; b < e->nb; b++) {
      const double zq = qz[(size_t)i * e->nb + b];
      lq[(size_t)i * e->nb + b] = (zq > kZmin) ? log(zq) : logzmin;        // newviewGenericSpecial.c:983-984
      const double zr = (i < first) ? 1.0 : rz[(size_t)i * e->nb + b];
      lr[(size_t)i * e->nb + b] = (zr > kZmin) ? log(zr) : logzmin;
    }
  }
  unsigned char *mk = e->stage_h + e->o_mask;
  for (int m = 0; m < e->P; m++) mk[m] = mask ? (mask[m] != 0) : 1;
  if (max_scratch) *max_scratch = scratch_top;

  // Short descriptors (the partial traversals of a tree search): the few exponentials of the d-table are computed
  // here and travel with the descriptor -- one kernel launch less on the critical path of the call.
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    e->host_dtab[c] = false;
    if (k.tiles_h.empty() || k.nparts == 0) continue;
    const int R = e->cat ? e->max_cats : 4;
    const size_t doubles = (size_t)count * k.nparts * k.dtab_row;
    const size_t exps = (size_t)count * k.nparts * R * 2 * (k.S - 1);
    if (exps > (size_t)kHostDtabMaxExps || doubles * sizeof(double) > kHostDtabBytes) continue;
    e->host_dtab[c] = true;
    off = (off + 127) & ~(size_t)127;               // the kernels read the table with 256-bit loads
    e->o_dtab[c] = off;
    double *dt = (double *)(e->stage_h + off);
    off = align16(off + doubles * sizeof(double));
    for (int i = 0; i < count; i++)
      for (int pc = 0; pc < k.nparts; pc++) {
        const double *model = &k.models_h[(size_t)pc * k.model_doubles];     // EIGN[S] rates[4] ...
        const HostPart &hp = e->parts[k.cp_global[pc]];
        const int slot = e->nb > 1 ? k.cp_global[pc] : 0;
        double *row = dt + ((size_t)i * k.nparts + pc) * k.dtab_row;
        const int rates_off = c == CLS_AA ? kAaRates : k.S, estride = c == CLS_AA ? 20 : 0;
        const bool lg4 = c == CLS_AA && model[kAaLg4] != 0.0;
        for (int cc = 0; cc < R; cc++) {
          const double rate = e->cat ? (cc < (int)hp.cat_rates.size() ? hp.cat_rates[cc] : 0.0) : model[rates_off + cc];
          const double *eign = model + (e->cat ? 0 : cc) * estride;
          for (int side = 0; side < 2; side++) {
            const double lz = side ? lr[(size_t)i * e->nb + slot] : lq[(size_t)i * e->nb + slot];
            double *d = row + ((size_t)cc * 2 + side) * k.S;
            d[0] = 1.0;
            for (int kk = 1; kk < k.S; kk++)               // same association as makeP / makeP_FlexLG4
              d[kk] = lg4 ? exp(rate * eign[kk] * lz) : exp(rate * (eign[kk] * lz));
          }
        }
      }
  }
  e->stage_use = e->stage_d;
  e->stage_used_bytes = off;
  if (copy) CUDA_TRY(cudaMemcpyAsync(e->stage_d, e->stage_h, off, cudaMemcpyHostToDevice, e->stream));
  return EXB_OK;
}

// -S: walk the descriptor on the host and bring the block table up to date: node p is all-gap in a tile iff both children
// are (x3_gap = x1_gap & x2_gap, newviewGenericSpecial.c:1182); all-gap pairs give their block back, the others get one.
// Also sums, per (entry, partition), the pattern weights of the all-gap tiles (scaling events of all-gap sites).
static int sev_update(exb_engine *e, const exb_entry *ti, int count, int first, const unsigned char *mask)
{
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    SevClass &v = k.sev;
    if (!v.on) continue;
    const int nt = v.ntiles;
    std::fill(v.gapw_h.begin(), v.gapw_h.begin() + (size_t)count * k.nparts, 0);
    for (int i = first; i < count; i++) {
      const int p = ti[i].pNumber, q = ti[i].qNumber, r = ti[i].rNumber;
      const size_t row = (size_t)(p - e->mxtips - 1) * nt;
      const uint8_t *gq = q <= e->mxtips ? &v.tipgap[(size_t)(q - 1) * nt] : &v.nodegap[(size_t)(q - e->mxtips - 1) * nt];
      const uint8_t *gr = r <= e->mxtips ? &v.tipgap[(size_t)(r - 1) * nt] : &v.nodegap[(size_t)(r - e->mxtips - 1) * nt];
      bool dirty = false;
      for (int t = 0; t < nt; t++) {
        const Tile &tl = k.tiles_h[t];
        if (mask && !mask[tl.part]) continue;        // a masked partition is not recomputed: its CLVs and gap state stay
        const uint8_t g = gq[t] & gr[t];
        v.nodegap[row + t] = g;
        double *&blk = v.table_h[row + t];
        if (g) {
          v.gapw_h[(size_t)i * k.nparts + tl.part_cls] += v.tile_wgt[t];
          if (blk) { v.free_blocks.push_back(blk); blk = nullptr; v.blocks_in_use--; dirty = true; }
        } else if (!blk) {
          if (v.free_blocks.empty()) {
            void *slab = nullptr;
            const size_t nblk = kSevSlabBlocks(v.block_doubles);
            if (cudaMalloc(&slab, nblk * v.block_doubles * sizeof(double)) != cudaSuccess) {
              cudaGetLastError();
              return fail(EXB_ERR_CUDA, "-S: out of HBM after %zu CLV blocks of %.1f KB (%zu slabs)", v.blocks_in_use,
                          v.block_doubles * sizeof(double) / 1024.0, v.slabs.size());
            }
            v.slabs.push_back(slab);
            for (size_t b = nblk; b-- > 0;) v.free_blocks.push_back((double *)slab + b * v.block_doubles);
          }
          blk = v.free_blocks.back(); v.free_blocks.pop_back(); v.blocks_in_use++; dirty = true;
        }
      }
      if (dirty) CUDA_TRY(cudaMemcpyAsync(v.table_d + row, &v.table_h[row], (size_t)nt * sizeof(double *), cudaMemcpyHostToDevice, e->stream));
    }
    if (count > first)
      CUDA_TRY(cudaMemcpyAsync(v.gapw_d, v.gapw_h.data(), (size_t)count * k.nparts * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  }
  return EXB_OK;
}

static int score_prepare(exb_engine *e, int ncand, int max_scratch);

static int run_traversal(exb_engine *e, const exb_entry *ti, int count, int first, const double *qz, const double *rz,
                         const unsigned char *mask, int mode, int ncand = -1)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (!ti || !qz || (count > first && !rz)) return fail(EXB_ERR_ARG, "null traversal arguments");
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = check_models(e, mask);
  if (rc) return rc;
  // the previous call has been synchronised, so the pinned staging area is free
  int max_scratch = -1;
  rc = stage_traversal(e, ti, count, first, qz, rz, mask, ncand, &max_scratch, false);
  if (rc) return rc;
  // Short descriptor on a small DNA alignment: everything travels in the kernel parameters and the traverse kernel
  // finishes the call itself (SmallCall, engine.cuh) -- no copy, one launch
  static const bool small_ok = !getenv("EXB_NO_SMALL_CALLS");
  const ClassState &kd = e->cls[CLS_DNA];
  const bool small = small_ok && mode != MODE_SCORE && !e->cat && !e->sev && e->cls[CLS_AA].tiles_h.empty() && !kd.tiles_h.empty() &&
                     kd.nparts <= kMaxParamPartitions && e->P <= kSmallParts && count <= kSmallEntries && e->host_dtab[CLS_DNA] &&
                     (size_t)count * kd.nparts * kDnaDtabRow <= (size_t)kSmallDtabDoubles && e->total_tiles <= kSmallMaxTiles &&
                     !(e->comm && !e->p2p) && 2 * (size_t)e->mxtips * sizeof(unsigned int) <= 40 * 1024;
  if (small) e->stage_use = e->stage_h;              // (only a finish kernel that has to run on its own reads it, zero-copy)
  else CUDA_TRY(cudaMemcpyAsync(e->stage_d, e->stage_h, e->stage_used_bytes, cudaMemcpyHostToDevice, e->stream));
  if (mode == MODE_SCORE) { rc = score_prepare(e, ncand, max_scratch); if (rc) return rc; }
  rc = flush_models(e);
  if (rc) return rc;
  if (e->sev) { rc = sev_update(e, ti, count, first, mask); if (rc) return rc; }
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
  const double *dtab_use[2] = {nullptr, nullptr};
  for (int c = 0; c < 2; c++) {                    // tables first: d-table (makeP exponentials), AA tip look-up table
    ClassState &k = e->cls[c];
    if (k.tiles_h.empty()) continue;
    dtab_use[c] = k.dtab;
    if (e->host_dtab[c]) dtab_use[c] = (const double *)(e->stage_use + e->o_dtab[c]);
    else {
      DtabArgs d;
      d.dtab = k.dtab; d.models = k.models; d.cls_part_global = k.part_global;
      d.lzq = (const double *)(e->stage_use + e->o_lzq); d.lzr = (const double *)(e->stage_use + e->o_lzr);
      d.count = count; d.nparts_cls = k.nparts; d.nbranches = e->nb; d.states = k.S;
      d.model_doubles = k.model_doubles; d.row = k.dtab_row;
      d.cat_rates = e->cat ? k.cat_rates : nullptr; d.max_cats = e->max_cats;
      d.eign_cat_stride = c == CLS_AA ? 20 : 0; d.rates_off = c == CLS_AA ? kAaRates : k.S; d.lg4_off = c == CLS_AA ? kAaLg4 : -1;
      launch_dtab(d, e->stream);
      e->launches++;
    }
    if (k.sev.on) {                                 // -S: gap columns of the descriptor's nodes (needs the d-table)
      GapcolArgs g;
      g.gapcol = k.sev.gapcol; g.entries = (const DevEntry *)(e->stage_use + e->o_entries);
      g.first = first; g.count = count; g.mxtips = e->mxtips;
      g.dtab = dtab_use[c]; g.dtab_row = k.dtab_row; g.nparts_cls = k.nparts;
      g.models = k.models; g.model_doubles = k.model_doubles; g.states = k.S;
      g.gapw = k.sev.gapw_d; g.cls_part_global = k.part_global; g.mask = e->stage_use + e->o_mask;
      g.scale_inc = e->scale_inc; g.nparts = e->P;
      launch_gapcol(g, e->stream);
      if (count > first) e->launches++;
    }
    if (c == CLS_AA && k.lut) {
      LutArgs l;
      l.lut = k.lut; l.models = k.models; l.dtab = dtab_use[c];
      l.entries = (const DevEntry *)(e->stage_use + e->o_entries);
      l.count = count; l.first = first; l.nparts_cls = k.nparts;
      l.ncat = e->cat ? e->max_cats : 4; l.dtab_row = k.dtab_row; l.cat = e->cat ? 1 : 0;
      launch_aa_lut(l, e->stream);
      e->launches++;
    }
  }
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev_a, e->stream));
  FinishArgs f;
  memset(&f, 0, sizeof f);
  bool folded = false;
  if (mode != MODE_SCORE) {                        // (a scoring call ends in its own two kernels)
    f.gs = e->gs; f.gs_stride = e->gs_stride; f.scale_inc = e->scale_inc;
    f.entries = (const DevEntry *)(e->stage_use + e->o_entries);
    f.count = count; f.first = first; f.mxtips = e->mxtips; f.nparts = e->P;
    f.mask = e->stage_use + e->o_mask; f.part_has_sites = e->part_has_sites;
    f.tile_out[0] = e->cls[0].tile_out; f.tile_out[1] = e->cls[1].tile_out;
    f.part_tile_begin = e->part_tile_begin; f.part_tile_count = e->part_tile_count; f.part_cls = e->part_cls;
    f.mode = mode; f.log_minlikelihood = log(kMinLikelihood); f.out = e->out_d;
    set_finish_outputs(e, f);
  }
  SmallCall sc;
  int last_launched = -1;
  if (small) {
    const unsigned char *mk = e->stage_h + e->o_mask;
    for (int pc = 0; pc < kd.nparts; pc++)
      if (kd.cp_tile_count[pc] != 0 && mk[kd.cp_global[pc]]) last_launched = pc;
    sc.count = count; sc.fold = 0;
    memcpy(sc.entries, e->stage_h + e->o_entries, (size_t)count * sizeof(DevEntry));
    memset(sc.mask, 0, sizeof sc.mask);
    memcpy(sc.mask, mk, (size_t)e->P);
    memcpy(sc.dtab, e->stage_h + e->o_dtab[CLS_DNA], (size_t)count * kd.nparts * kDnaDtabRow * sizeof(double));
    sc.fin = f;
    sc.fin.gs_in_smem = 1;
  }
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    if (k.tiles_h.empty()) continue;
    TraverseArgs a;
    a.dtab = dtab_use[c]; a.lut = k.lut; a.nparts_cls = k.nparts; a.tile_sites = e->tile_sites;
    a.rate_cat = e->cat ? k.rate_cat : nullptr; a.max_cats = e->max_cats; a.dtab_row = k.dtab_row;
    a.tiles = k.tiles; a.clv = k.clv; a.tips = k.tips; a.wgt = k.wgt; a.sum = k.sum; a.models = k.models;
    a.N = k.N; a.mxtips = e->mxtips; a.nparts = e->P; a.nbranches = e->nb;
    a.entries = (const DevEntry *)(e->stage_use + e->o_entries);
    a.lzq = (const double *)(e->stage_use + e->o_lzq);
    a.lzr = (const double *)(e->stage_use + e->o_lzr);
    a.mask = e->stage_use + e->o_mask;
    a.count = count; a.first = first; a.mode = mode;
    a.scale_inc = e->scale_inc; a.tile_out = k.tile_out;
    a.sev_table = k.sev.on ? k.sev.table_d : nullptr; a.sev_tipgap = nullptr; a.sev_gapcol = k.sev.gapcol;
    a.sev_ntiles = k.sev.ntiles; a.sev_tile0 = 0;
    a.scratch_off = mode == MODE_SCORE && k.scratch ? (long long)(k.scratch - k.clv) : 0;
    a.score_out = mode == MODE_SCORE ? k.score_out : nullptr;
    a.score_stride = (int)k.tiles_h.size() * (c == CLS_DNA ? kDnaWarps : kAaWarps);
    if (c == CLS_DNA && k.nparts <= kMaxParamPartitions) {
      // few partitions: one launch per partition, its EI/EV travel as kernel parameters (constant bank)
      for (int pc = 0; pc < k.nparts; pc++) {
        if (k.cp_tile_count[pc] == 0 || !(e->stage_h + e->o_mask)[k.cp_global[pc]]) continue;   // empty or masked
        DnaConsts kc;
        memcpy(kc.EI, &k.models_h[(size_t)pc * k.model_doubles + 8], sizeof kc.EI);
        memcpy(kc.EV, &k.models_h[(size_t)pc * k.model_doubles + 24], sizeof kc.EV);
        TraverseArgs ap = a;
        ap.tiles = k.tiles + k.cp_tile_begin[pc];
        ap.sev_tile0 = k.cp_tile_begin[pc];
        ap.tile_out = k.tile_out + 2 * (size_t)k.cp_tile_begin[pc];
        if (small) {                                // the last launch of the call also finishes it
          sc.fold = pc == last_launched;
          folded = folded || sc.fold;
        }
        if (e->cat) launch_traverse_dna_cat(ap, &kc, k.cp_tile_count[pc], e->stream);
        else        launch_traverse_dna(ap, &kc, k.cp_tile_count[pc], e->stream, small ? &sc : nullptr);
        if (pc) e->launches++;
      }
    } else if (c == CLS_DNA) {
      if (e->cat) launch_traverse_dna_cat(a, nullptr, (int)k.tiles_h.size(), e->stream);
      else        launch_traverse_dna(a, nullptr, (int)k.tiles_h.size(), e->stream);   // many partitions: constants via shared memory
    } else {
      bool any_lg4 = false;                        // per-category eigen-systems anywhere in the class?
      for (int pc = 0; pc < k.nparts; pc++) any_lg4 = any_lg4 || k.models_h[(size_t)pc * k.model_doubles + kAaLg4] != 0.0;
      if (e->aa_tc && !e->cat && !any_lg4 && k.lut) {
        // mixed-precision mode: the newview entries on tcgen05 (3xTF32), the branch tail (evaluate / sum table) by the
        // FP64 kernel on the CLVs just written
        if (count > first) { launch_traverse_aa_tc(a, (int)k.tiles_h.size(), e->stream); e->launches++; }
        if (mode != MODE_NEWVIEW) {
          TraverseArgs tail = a;
          tail.first = count;
          launch_traverse_aa(tail, any_lg4, (int)k.tiles_h.size(), e->stream);
        } else e->launches--;
      } else
        launch_traverse_aa(a, any_lg4, (int)k.tiles_h.size(), e->stream);
    }
    e->launches++;
  }
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev_b, e->stream));
  e->phases_valid = e->timing;
  if (mode == MODE_SCORE) {
    ScoreFinishArgs sf;
    sf.gs = e->gs; sf.gs_stride = e->gs_stride; sf.scale_inc = e->scale_inc;
    sf.entries = (const DevEntry *)(e->stage_use + e->o_entries);
    sf.count = count; sf.mxtips = e->mxtips; sf.nparts = e->P; sf.ncand = ncand;
    sf.mask = e->stage_use + e->o_mask; sf.part_has_sites = e->part_has_sites;
    sf.sc = e->score_sc;
    for (int c = 0; c < 2; c++) {
      sf.score_out[c] = e->cls[c].score_out;
      sf.warps[c] = c == CLS_DNA ? kDnaWarps : kAaWarps;
      sf.score_stride[c] = (int)e->cls[c].tiles_h.size() * sf.warps[c];
    }
    sf.part_tile_begin = e->part_tile_begin; sf.part_tile_count = e->part_tile_count; sf.part_cls = e->part_cls;
    sf.log_minlikelihood = log(kMinLikelihood); sf.out = e->score_d;
    launch_score_finish(sf, e->stream);
    e->launches += 2;
    if (e->timing) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaGetLastError());
    return EXB_OK;
  }
  if (!folded) {
    launch_finish(f, e->stream);
    e->launches++;
  }
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
  CUDA_TRY(cudaGetLastError());
  return EXB_OK;
}

static P2pArgs p2p_args(const exb_engine *e)
{
  P2pArgs x;
  memset(&x, 0, sizeof x);
  x.nranks = e->p2p ? e->nranks : 1; x.rank = e->rank;
  for (int r = 0; r < e->nranks && e->p2p; r++) { x.peer_buf[r] = e->peer_buf[r]; x.peer_flags[r] = e->peer_flags[r]; }
  return x;
}

// Single rank: the finish kernel writes results and a completion flag into pinned host memory; the host polls.
static void set_finish_outputs(exb_engine *e, FinishArgs &f)
{
  f.ticket = e->ticket;
  f.x.nranks = 1; f.x.rank = 0; f.xseq = 0;
  if (e->comm && !e->p2p) { f.out_host = nullptr; f.done_flag_host = nullptr; f.seq = 0; return; }   // NCCL after the kernel
  if (++e->seq == 0xFFFFFFFFu) e->seq = 1;          // 0xFFFFFFFF is the kernel's "peer lost" marker
  f.out_host = e->out_h; f.done_flag_host = e->done_flag; f.seq = e->seq;
  if (e->p2p) {
    f.x = p2p_args(e);
    if (f.mode == MODE_EVALUATE || f.mode == MODE_CORE) e->xseq += 1;
    f.xseq = e->xseq;
  }
}

static int wait_done(exb_engine *e)
{
  if (e->comm && !e->p2p) { CUDA_TRY(cudaStreamSynchronize(e->stream)); }
  else {
    unsigned long spins = 0;
    for (;;) {
      // acquire: the results in out_h were written before the flag (the kernel orders them with __threadfence_system)
      const unsigned int flag = __atomic_load_n((const unsigned int *)e->done_flag, __ATOMIC_ACQUIRE);
      if (flag == e->seq) break;
      if (flag == 0xFFFFFFFFu)
        return fail(EXB_ERR_NCCL, "a peer rank did not deliver its partial sums within 20 s (all-reduce inside finish_kernel)");
      if ((++spins & 0xFFFF) == 0) {                 // a faulting kernel must not hang the host
        const cudaError_t q = cudaStreamQuery(e->stream);
        if (q != cudaSuccess && q != cudaErrorNotReady)
          return fail(EXB_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(q));
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#elif defined(__aarch64__)
      asm volatile("yield" ::: "memory");
#endif
    }
  }
  if (e->timing) { CUDA_TRY(cudaEventSynchronize(e->ev1)); cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1); }
  return EXB_OK;
}

// all-reduce (if a communicator exists), copy n doubles of out_d to the pinned mirror, wait.
static int collect(exb_engine *e, size_t n)
{
  if (!e->comm || e->p2p) return wait_done(e);       // (globally reduced) results already sit in out_h
  if (e->comm) {
    const int r = g_nccl.AllReduce(e->out_d, e->out_d, n, kNcclFloat64, kNcclSum, e->comm, e->stream);
    if (r != 0) return fail(EXB_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  }
  if (n) CUDA_TRY(cudaMemcpyAsync(e->out_h, e->out_d, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  return wait_done(e);
}

extern "C" int exb_newview(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                           const unsigned char *executeModel)
{
  if (count == 0) return EXB_OK;
  int rc = run_traversal(e, ti, count, 0, qz, rz, executeModel, MODE_NEWVIEW);
  if (rc) return rc;
  return wait_done(e);                            // the staging buffers are reused by the next call
}

extern "C" int exb_evaluate(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                            const unsigned char *executeModel, double *perPartitionLH)
{
  if (count < 1 || !perPartitionLH) return fail(EXB_ERR_ARG, "exb_evaluate: needs the branch entry and an output array");
  int rc = run_traversal(e, ti, count, 1, qz, rz, executeModel, MODE_EVALUATE);
  if (rc) return rc;
  rc = collect(e, (size_t)e->P);
  if (rc) return rc;
  for (int m = 0; m < e->P; m++)
    if (!executeModel || executeModel[m]) perPartitionLH[m] = e->out_h[m];   // masked partitions keep their value
  return EXB_OK;
}

extern "C" int exb_makenewz_prepare(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                                    const unsigned char *executeModel)
{
  if (count < 1) return fail(EXB_ERR_ARG, "exb_makenewz_prepare: needs the branch entry");
  int rc = run_traversal(e, ti, count, 1, qz, rz, executeModel, MODE_SUM);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (e->timing) cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1);
  return EXB_OK;
}

extern "C" int exb_core(exb_engine *e, const double *lz, const unsigned char *executeModel,
                        double *dlnLdlz, double *d2lnLdlz2)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (!lz || !dlnLdlz || !d2lnLdlz2) return fail(EXB_ERR_ARG, "exb_core: null argument");
  CUDA_TRY(cudaSetDevice(e->device));
  memcpy(e->stage_h + e->off_corelz, lz, (size_t)e->nb * sizeof(double));
  unsigned char *mk = e->stage_h + e->off_mask;
  for (int m = 0; m < e->P; m++) mk[m] = executeModel ? (executeModel[m] != 0) : 1;
  // core-lz and mask are adjacent at the tail of the staging area: one small copy
  e->stage_use = e->stage_d;
  CUDA_TRY(cudaMemcpyAsync(e->stage_d + e->off_corelz, e->stage_h + e->off_corelz, e->off_mask + (size_t)e->P - e->off_corelz,
                           cudaMemcpyHostToDevice, e->stream));
  int rc = flush_models(e);
  if (rc) return rc;
  e->phases_valid = false;
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    if (k.tiles_h.empty()) continue;
    CoreArgs a;
    a.tiles = k.tiles; a.sum = k.sum; a.wgt = k.wgt; a.models = k.models;
    a.rate_cat = e->cat ? k.rate_cat : nullptr; a.cat_rates = k.cat_rates; a.max_cats = e->max_cats;
    a.lz = (const double *)(e->stage_use + e->off_corelz);
    a.mask = e->stage_use + e->off_mask; a.nbranches = e->nb; a.tile_out = k.tile_out;
    if (c == CLS_DNA) launch_core_dna(a, (int)k.tiles_h.size(), e->stream);
    else              launch_core_aa(a, (int)k.tiles_h.size(), e->stream);
    e->launches++;
  }
  FinishArgs f;
  memset(&f, 0, sizeof f);
  f.gs = e->gs; f.gs_stride = e->gs_stride; f.scale_inc = e->scale_inc; f.entries = (const DevEntry *)(e->stage_use + e->off_entries);
  f.count = 0; f.first = 0; f.mxtips = e->mxtips; f.nparts = e->P;
  f.mask = e->stage_use + e->off_mask; f.part_has_sites = e->part_has_sites;
  f.tile_out[0] = e->cls[0].tile_out; f.tile_out[1] = e->cls[1].tile_out;
  f.part_tile_begin = e->part_tile_begin; f.part_tile_count = e->part_tile_count; f.part_cls = e->part_cls;
  f.mode = MODE_CORE; f.log_minlikelihood = 0.0; f.out = e->out_d;
  set_finish_outputs(e, f);
  launch_finish(f, e->stream);
  e->launches++;
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
  CUDA_TRY(cudaGetLastError());
  rc = collect(e, (size_t)2 * e->P);
  if (rc) return rc;
  // per-partition sums -> branch slots, in partition order (makenewzGenericSpecial.c:921-946, 1075-1076)
  for (int b = 0; b < e->nb; b++) { dlnLdlz[b] = 0.0; d2lnLdlz2[b] = 0.0; }
  for (int m = 0; m < e->P; m++) {
    if (executeModel && !executeModel[m]) continue;
    const int slot = e->nb > 1 ? m : 0;
    dlnLdlz[slot] += e->out_h[m];
    d2lnLdlz2[slot] += e->out_h[e->P + m];
  }
  return EXB_OK;
}

extern "C" int exb_makenewz(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                            const unsigned char *executeModel, const double *z0, int maxiter,
                            double *result, int *curvatOK, int *iterations)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (e->comm && !e->p2p)
    return fail(EXB_ERR_STATE, "exb_makenewz: with several ranks the device-resident Newton-Raphson needs the in-kernel all-reduce "
                "(all ranks on one node with peer access, exb_comm_is_p2p); use exb_makenewz_prepare + exb_core");
  if (count < 1 || !z0 || !result) return fail(EXB_ERR_ARG, "exb_makenewz: bad arguments");
  if (e->nb > kMaxBranchSlots) return fail(EXB_ERR_ARG, "exb_makenewz: more than %d branch slots", kMaxBranchSlots);
  // 1. newviews + sum table (the first-iteration work of topLevelMakenewz); the mask of that pass is the one the
  //    state machine derives in its first round: executeModel && !curvatOK, with curvatOK == FALSE for every slot
  int rc = run_traversal(e, ti, count, 1, qz, rz, executeModel, MODE_SUM);
  if (rc) return rc;
  // 2. initial state: as kernel parameters when it is small (no upload at all), else a NrState + mask upload
  NrState *h = e->nr_state_h;
  NrArgs a;
  memset(&a, 0, sizeof a);
  const bool from_args = e->nb <= kNrInitSlots && e->P <= kNrInitSlots;
  if (from_args) {
    a.init_from_args = 1; a.maxiter = maxiter;
    for (int b = 0; b < e->nb; b++) a.z0[b] = z0[b];
    for (int m = 0; m < e->P; m++) a.mask0[m] = (e->stage_h + e->o_mask)[m];
  } else {
    memset(h, 0, sizeof *h);
    for (int b = 0; b < e->nb; b++) { h->z[b] = z0[b]; h->z0[b] = z0[b]; h->budget[b] = maxiter; h->done[b] = 0; h->curvatOK[b] = 1; }
    CUDA_TRY(cudaMemcpyAsync(e->nr_state_d, h, sizeof *h, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->nr_mask, e->stage_h + e->o_mask, (size_t)e->P, cudaMemcpyHostToDevice, e->stream));
  }
  a.state = e->nr_state_d; a.mask = e->nr_mask; a.barrier = e->nr_barrier;
  int total_tiles = 0;
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    a.tiles[c] = k.tiles; a.ntiles[c] = (int)k.tiles_h.size(); a.sum[c] = k.sum; a.wgt[c] = k.wgt;
    a.models[c] = k.models; a.tile_out[c] = k.tile_out;
    a.rate_cat[c] = e->cat ? k.rate_cat : nullptr; a.cat_rates[c] = k.cat_rates;
    total_tiles += a.ntiles[c];
  }
  a.part_out = e->out_d;
  a.part_has_sites = e->part_has_sites; a.part_tile_begin = e->part_tile_begin; a.part_tile_count = e->part_tile_count;
  a.part_cls = e->part_cls; a.nparts = e->P; a.nbranches = e->nb; a.max_cats = e->max_cats;
  a.x = p2p_args(e); a.xseq0 = e->xseq;
  // a handful of tiles: ONE CTA walks them (no grid barrier, no cooperative launch) -- on such alignments the latency of
  // the call is everything; otherwise persistent CTAs with grid barriers
  // ... and when (GAMMA) the sum table of those tiles fits in shared memory it is staged there once for all iterations
  size_t staged = 0;
  if (!e->cat && from_args && total_tiles <= 16) {
    for (int m = 0; m < e->P; m++) {
      const HostPart &p = e->parts[m];
      a.small_site0[m] = (unsigned int)p.site0; a.small_width[m] = (unsigned int)p.width;
      a.small_cls[m] = p.cls; a.small_pc[m] = p.part_cls;
      staged += ((p.width * (e->cls[p.cls].span + 1) * sizeof(double) + 15) & ~(size_t)15) + ((p.width * sizeof(int) + 15) & ~(size_t)15);
    }
    if (staged > kNrStagedBytes) staged = 0;
    else if (staged == 0) staged = 16;              // (no local sites at all: still the one-CTA kernel)
  }
  const int blocks = (staged || total_tiles <= kNrSingleCtaTiles) ? 1 : std::max(1, std::min(total_tiles, e->nr_max_blocks));
  if (blocks > 1) CUDA_TRY(cudaMemsetAsync(e->nr_barrier, 0, sizeof(unsigned int), e->stream));
  // results: into pinned memory + the completion flag the host polls (the same flag the finish kernels raise)
  if (++e->seq == 0xFFFFFFFFu) e->seq = 1;
  a.result_host = h; a.done_flag_host = e->done_flag; a.seq = e->seq;
  const int lr = launch_newton_raphson(a, blocks, staged, e->stream);
  if (lr != 0) return fail(EXB_ERR_CUDA, "launch of the Newton-Raphson kernel failed: %s", cudaGetErrorString((cudaError_t)lr));
  e->launches++;
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
  {
    unsigned long spins = 0;
    while (__atomic_load_n((const unsigned int *)e->done_flag, __ATOMIC_ACQUIRE) != e->seq) {
      if ((++spins & 0xFFFF) == 0) {                 // a faulting kernel must not hang the host
        const cudaError_t q = cudaStreamQuery(e->stream);
        if (q != cudaSuccess && q != cudaErrorNotReady) return fail(EXB_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(q));
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#elif defined(__aarch64__)
      asm volatile("yield" ::: "memory");
#endif
    }
  }
  if (e->timing) { CUDA_TRY(cudaEventSynchronize(e->ev1)); cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1); }
  if (e->p2p) {
    if (h->peer_lost) return fail(EXB_ERR_NCCL, "a peer rank did not deliver its derivative sums within 20 s (all-reduce inside newton_raphson_kernel)");
    e->xseq = h->xseq;                              // every rank ran the same number of exchanges
  }
  for (int b = 0; b < e->nb; b++) {
    result[b] = h->z[b];
    if (curvatOK) curvatOK[b] = h->curvatOK[b];
  }
  if (iterations) *iterations = h->iterations;
  return EXB_OK;
}

// ---- batched scoring of lazy SPR insertions (testInsertBIG searchAlgo.c:682) ----
extern "C" int exb_max_entries(exb_engine *e) { return e && e->committed ? e->max_entries : 0; }

extern "C" int exb_can_score_insertions(exb_engine *e)
{
  if (!e || !e->committed || e->cat || e->sev) return 0;
  if (e->aa_tc && !e->cls[CLS_AA].tiles_h.empty()) return 0;
  // LG4M / LG4X partitions: the scoring tail has the per-category weights, but no parity test covers it yet -- refuse rather
  // than answer unverified (the caller then scores candidate by candidate)
  const ClassState &k = e->cls[CLS_AA];
  for (int pc = 0; pc < k.nparts; pc++)
    if (k.models_h[(size_t)pc * k.model_doubles + kAaLg4] != 0.0) return 0;
  return 1;
}

// Buffers of a scoring call: scratch CLVs up to node index max_scratch (grown with their contents: a scan may be cut into
// several calls and deeper levels can appear later), per-(candidate, tile, warp) partials, scaler sums, results.
static int score_prepare(exb_engine *e, int ncand, int max_scratch)
{
  for (int c = 0; c < 2; c++) {
    ClassState &k = e->cls[c];
    if (k.tiles_h.empty()) continue;
    if (max_scratch >= k.scratch_nodes) {
      const int want = std::min(kScratchNodes, std::max(max_scratch + 1, k.scratch_nodes + 4));
      const size_t node_bytes = k.N * k.span * sizeof(double);
      double *grown = nullptr;
      if (cudaMalloc(&grown, (size_t)want * node_bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(EXB_ERR_CUDA, "exb_score_insertions: cannot allocate %d scratch CLVs of %.2f GB", want, node_bytes / 1e9);
      }
      if (k.scratch) {
        CUDA_TRY(cudaMemcpyAsync(grown, k.scratch, (size_t)k.scratch_nodes * node_bytes, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        cudaFree(k.scratch);
      }
      k.scratch = grown; k.scratch_nodes = want;
    }
    const size_t need = (size_t)ncand * k.tiles_h.size() * (c == CLS_DNA ? kDnaWarps : kAaWarps) * sizeof(double2);
    if (need > k.score_out_cap) {
      cudaFree(k.score_out); k.score_out = nullptr; k.score_out_cap = 0;
      if (cudaMalloc(&k.score_out, need + need / 4) != cudaSuccess) { cudaGetLastError(); return fail(EXB_ERR_CUDA, "exb_score_insertions: out of HBM for the partial sums"); }
      k.score_out_cap = need + need / 4;
    }
  }
  const size_t nres = (size_t)ncand * e->P;
  if (nres * sizeof(unsigned int) > e->score_sc_cap) {
    cudaFree(e->score_sc); e->score_sc = nullptr; e->score_sc_cap = 0;
    CUDA_TRY(cudaMalloc(&e->score_sc, 2 * nres * sizeof(unsigned int)));
    e->score_sc_cap = 2 * nres * sizeof(unsigned int);
  }
  if (nres * sizeof(double) > e->score_d_cap) {
    cudaFree(e->score_d); e->score_d = nullptr; e->score_d_cap = 0;
    CUDA_TRY(cudaMalloc(&e->score_d, 2 * nres * sizeof(double)));
    e->score_d_cap = 2 * nres * sizeof(double);
  }
  if (nres * sizeof(double) > e->score_h_cap) {
    cudaFreeHost(e->score_h); e->score_h = nullptr; e->score_h_cap = 0;
    CUDA_TRY(cudaMallocHost(&e->score_h, 2 * nres * sizeof(double)));
    e->score_h_cap = 2 * nres * sizeof(double);
  }
  return EXB_OK;
}

extern "C" int exb_score_insertions(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                                    const unsigned char *executeModel, int ncand, double *lnl)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (!exb_can_score_insertions(e))
    return fail(EXB_ERR_STATE, "exb_score_insertions: not available under PSR, -S, the tf32x3 protein mode or with LG4M / LG4X partitions (score the candidates one by one)");
  if (count < 1 || ncand < 0 || (ncand > 0 && !lnl)) return fail(EXB_ERR_ARG, "exb_score_insertions: bad arguments");
  int rc = run_traversal(e, ti, count, 1, qz, rz, executeModel, MODE_SCORE, ncand);
  if (rc) return rc;
  const size_t n = (size_t)ncand * e->P;
  if (n && e->comm) {
    const int r = g_nccl.AllReduce(e->score_d, e->score_d, n, kNcclFloat64, kNcclSum, e->comm, e->stream);
    if (r != 0) return fail(EXB_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  }
  if (n) CUDA_TRY(cudaMemcpyAsync(e->score_h, e->score_d, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (e->timing) cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1);
  if (n) memcpy(lnl, e->score_h, n * sizeof(double));
  return EXB_OK;
}

// ---- per-site likelihood at explicit rates (evaluatePartialGeneric, evaluatePartialGenericSpecial.c:259) ----
template <class T>
static cudaError_t grow(T **p, size_t *cap, size_t need)
{
  if (need <= *cap) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  const size_t want = need + need / 4;
  cudaError_t r = cudaMalloc((void **)p, want);
  if (r == cudaSuccess) *cap = want;
  return r;
}

extern "C" int exb_evaluate_partial(exb_engine *e, const exb_entry *ti, int count, const double *qz, const double *rz,
                                    size_t n, const int *models, const int *sites, const double *rates, double *out)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (!e->cat) return fail(EXB_ERR_STATE, "exb_evaluate_partial: only under the CAT / PSR rate model (evaluatePartialGenericSpecial.c:326)");
  if (!ti || !qz || !rz || count < 1) return fail(EXB_ERR_ARG, "exb_evaluate_partial: null traversal arguments");
  if (n == 0) return EXB_OK;
  if (!models || !sites || !rates || !out) return fail(EXB_ERR_ARG, "exb_evaluate_partial: null pair arrays");
  if (n >= ((size_t)1 << 31)) return fail(EXB_ERR_ARG, "exb_evaluate_partial: more than 2^31 pairs in one call");
  if (ti[0].pNumber < 1 || ti[0].pNumber > e->mxtips || ti[0].qNumber <= e->mxtips)
    return fail(EXB_ERR_ARG, "exb_evaluate_partial: the root branch must join a tip (pNumber=%d) and an inner node (qNumber=%d), "
                "evaluatePartialGenericSpecial.c:210", ti[0].pNumber, ti[0].qNumber);
  CUDA_TRY(cudaSetDevice(e->device));
  int rc = stage_traversal(e, ti, count, 1, qz, rz, nullptr);      // validation, lzq / lzr on the device
  if (rc) return rc;
  rc = flush_models(e);
  if (rc) return rc;

  // slot program: a post-order traversal keeps few vectors alive; give every inner node a slot for its lifetime
  std::vector<PartialEntry> prog((size_t)count);
  std::vector<int> slot_of((size_t)2 * e->mxtips, -1), free_slots;
  int nslots = 0;
  for (int k = 1; k < count; k++) {
    const exb_entry &t = ti[k];
    int q = t.qNumber, r = t.rNumber;
    if (t.tipCase == EXB_INNER_INNER) {
      if (slot_of[q] < 0) return fail(EXB_ERR_ARG, "exb_evaluate_partial: entry %d reads inner node %d which this traversal does not compute "
                                      "(a FULL traversal is required, evaluatePartialGenericSpecial.c:214-226)", k, q);
      const int sq = slot_of[q]; slot_of[q] = -1; free_slots.push_back(sq); q = sq;
    }
    if (t.tipCase != EXB_TIP_TIP) {
      if (slot_of[r] < 0) return fail(EXB_ERR_ARG, "exb_evaluate_partial: entry %d reads inner node %d which this traversal does not compute "
                                      "(a FULL traversal is required, evaluatePartialGenericSpecial.c:214-226)", k, r);
      const int sr = slot_of[r]; slot_of[r] = -1; free_slots.push_back(sr); r = sr;
    }
    int sp;                                        // may reuse a child's slot: the thread holds both children in registers
    if (!free_slots.empty()) { sp = free_slots.back(); free_slots.pop_back(); } else sp = nslots++;
    slot_of[t.pNumber] = sp;
    prog[k] = PartialEntry{t.tipCase, sp, q, r};
  }
  if (slot_of[ti[0].qNumber] < 0)
    return fail(EXB_ERR_ARG, "exb_evaluate_partial: the traversal does not compute the root's inner node %d", ti[0].qNumber);
  prog[0] = PartialEntry{-1, ti[0].pNumber, slot_of[ti[0].qNumber], 0};
  if (nslots < 1) nslots = 1;

  // group the pairs by class and partition (counting sort, stable), cut each partition's run into chunks
  std::vector<size_t> pstart((size_t)e->P + 1, 0);
  for (size_t j = 0; j < n; j++) {
    const int m = models[j];
    if (m < 0 || m >= e->P) return fail(EXB_ERR_ARG, "exb_evaluate_partial: pair %zu names partition %d", j, m);
    if (sites[j] < 0 || (size_t)sites[j] >= e->parts[m].width)
      return fail(EXB_ERR_ARG, "exb_evaluate_partial: pair %zu names site %d of partition %d (width %zu)", j, sites[j], m, e->parts[m].width);
    if (!e->parts[m].model_set) return fail(EXB_ERR_STATE, "partition %d has no model constants (exb_set_model)", m);
    pstart[m + 1]++;
  }
  // order: class 0 partitions first, then class 1, each in partition order
  std::vector<int> porder;
  for (int c = 0; c < 2; c++) for (int m = 0; m < e->P; m++) if (e->parts[m].cls == c) porder.push_back(m);
  std::vector<size_t> pbase((size_t)e->P, 0);
  size_t run = 0, cls_begin[3] = {0, 0, 0};
  {
    int ci = 0;
    for (int m : porder) {
      while (ci < e->parts[m].cls) cls_begin[++ci] = run;
      pbase[m] = run; run += pstart[m + 1];
    }
    while (ci < 2) cls_begin[++ci] = run;
  }
  std::vector<PartialPair> pairs(n);
  std::vector<uint32_t> origin(n);
  {
    std::vector<size_t> fill = pbase;
    for (size_t j = 0; j < n; j++) {
      const HostPart &p = e->parts[models[j]];
      const size_t at = fill[models[j]]++;
      pairs[at] = PartialPair{rates[j], (uint32_t)(p.site0 + (size_t)sites[j]), 0u};
      origin[at] = (uint32_t)j;
    }
  }
  std::vector<PartialChunk> chunks[2];
  for (int m : porder) {
    const HostPart &p = e->parts[m];
    const size_t cnt = pstart[m + 1];
    for (size_t o = 0; o < cnt; o += kPartialThreads)
      chunks[p.cls].push_back(PartialChunk{(uint32_t)(pbase[m] + o - cls_begin[p.cls]), (uint32_t)std::min<size_t>(kPartialThreads, cnt - o),
                                           p.part_cls, e->nb > 1 ? m : 0});
  }
  std::vector<double> root_lz((size_t)e->nb);
  for (int b = 0; b < e->nb; b++) root_lz[b] = log(qz[b]);        // not clamped, :232-234

  // device image: prog | root_lz | pairs | chunks[0] | chunks[1]
  auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  size_t off = 0;
  const size_t o_prog = off;  off = align16(off + prog.size() * sizeof(PartialEntry));
  const size_t o_root = off;  off = align16(off + root_lz.size() * sizeof(double));
  const size_t o_pairs = off; off = align16(off + pairs.size() * sizeof(PartialPair));
  size_t o_chunks[2];
  for (int c = 0; c < 2; c++) { o_chunks[c] = off; off = align16(off + chunks[c].size() * sizeof(PartialChunk)); }
  CUDA_TRY(grow(&e->pp_in, &e->pp_in_cap, off));
  CUDA_TRY(grow(&e->pp_out, &e->pp_out_cap, n * sizeof(double)));
  CUDA_TRY(cudaMemcpyAsync(e->pp_in + o_prog, prog.data(), prog.size() * sizeof(PartialEntry), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->pp_in + o_root, root_lz.data(), root_lz.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  CUDA_TRY(cudaMemcpyAsync(e->pp_in + o_pairs, pairs.data(), pairs.size() * sizeof(PartialPair), cudaMemcpyHostToDevice, e->stream));
  for (int c = 0; c < 2; c++)
    if (!chunks[c].empty())
      CUDA_TRY(cudaMemcpyAsync(e->pp_in + o_chunks[c], chunks[c].data(), chunks[c].size() * sizeof(PartialChunk), cudaMemcpyHostToDevice, e->stream));
  int blocks[2] = {0, 0};
  size_t scratch_need = 0;
  for (int c = 0; c < 2; c++) {
    if (chunks[c].empty()) continue;
    if (!e->pp_max_blocks[c]) e->pp_max_blocks[c] = partial_max_blocks(e->cls[c].S, e->device);
    blocks[c] = (int)std::min<size_t>(chunks[c].size(), (size_t)e->pp_max_blocks[c]);
    scratch_need = std::max(scratch_need, (size_t)blocks[c] * nslots * e->cls[c].S * kPartialThreads * sizeof(double));
  }
  CUDA_TRY(grow(&e->pp_scratch, &e->pp_scratch_cap, scratch_need));
  e->phases_valid = false;
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
  for (int c = 0; c < 2; c++) {
    if (chunks[c].empty()) continue;
    ClassState &k = e->cls[c];
    PartialArgs a;
    a.chunks = (const PartialChunk *)(e->pp_in + o_chunks[c]); a.nchunks = (int)chunks[c].size();
    a.pairs = (const PartialPair *)(e->pp_in + o_pairs) + cls_begin[c];
    a.entries = (const PartialEntry *)(e->pp_in + o_prog); a.count = count;
    a.lzq = (const double *)(e->stage_d + e->o_lzq); a.lzr = (const double *)(e->stage_d + e->o_lzr);
    a.root_lz = (const double *)(e->pp_in + o_root); a.nbranches = e->nb;
    a.models = k.models; a.model_doubles = k.model_doubles; a.tips = k.tips; a.N = k.N; a.wgt = k.wgt;
    a.scratch = e->pp_scratch; a.nslots = nslots; a.log_minlikelihood = log(kMinLikelihood);
    a.out = e->pp_out + cls_begin[c];
    launch_partial(a, k.S, blocks[c], e->stream);       // the two classes share the scratch area: same stream, in order
    e->launches++;
  }
  if (e->timing) CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
  CUDA_TRY(cudaGetLastError());
  std::vector<double> res(n);
  CUDA_TRY(cudaMemcpyAsync(res.data(), e->pp_out, n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (e->timing) cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1);
  for (size_t j = 0; j < n; j++) out[origin[j]] = res[j];
  return EXB_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------------------------
extern "C" int exb_comm_unique_id(void *id_out)
{
  int rc = load_nccl();
  if (rc) return rc;
  NcclUniqueId id;
  const int r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail(EXB_ERR_NCCL, "ncclGetUniqueId failed (%d)", r);
  memcpy(id_out, &id, sizeof id);
  return EXB_OK;
}

// One node, every GPU can map every other GPU's memory: exchange CUDA-IPC handles of the receive buffers (through
// NCCL, once) so that finish_kernel can do the all-reduce itself.  Any rank that cannot -> everybody stays on NCCL.
static int setup_p2p(exb_engine *e)
{
  const int n = e->nranks;
  // default on one node; EXB_P2P=0 keeps ncclAllReduce after the kernel
  const char *want = getenv("EXB_P2P");
  if (n > kMaxP2pRanks || (want && atoi(want) == 0) || !g_nccl.AllGather) return EXB_OK;
  struct Card { cudaIpcMemHandle_t h; int ok; int device; char host[64]; };
  const size_t data_bytes = (size_t)2 * n * 2 * e->P * sizeof(double);
  const size_t flag_off = (data_bytes + 255) & ~(size_t)255;
  Card mine;
  memset(&mine, 0, sizeof mine);
  mine.device = e->device;
  gethostname(mine.host, sizeof mine.host - 1);
  mine.ok = cudaMalloc(&e->xbuf, flag_off + (size_t)n * sizeof(unsigned int)) == cudaSuccess &&
            cudaMemset(e->xbuf, 0, flag_off + (size_t)n * sizeof(unsigned int)) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.h, e->xbuf) == cudaSuccess;
  cudaGetLastError();
  std::vector<Card> cards((size_t)n);
  Card *cards_d = nullptr;
  CUDA_TRY(cudaMalloc(&cards_d, (size_t)n * sizeof(Card)));
  CUDA_TRY(cudaMemcpyAsync(cards_d + e->rank, &mine, sizeof mine, cudaMemcpyHostToDevice, e->stream));
  int r = g_nccl.AllGather(cards_d + e->rank, cards_d, sizeof(Card), kNcclInt8, e->comm, e->stream);
  if (r != 0) { cudaFree(cards_d); return fail(EXB_ERR_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); }
  CUDA_TRY(cudaMemcpyAsync(cards.data(), cards_d, (size_t)n * sizeof(Card), cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  bool all = true;
  for (int i = 0; i < n; i++) all = all && cards[i].ok && strncmp(cards[i].host, mine.host, sizeof mine.host) == 0;
  double bad = 0.0;
  if (all) {
    for (int i = 0; i < n && bad == 0.0; i++) {
      void *ptr = e->xbuf;
      if (i != e->rank && cudaIpcOpenMemHandle(&ptr, cards[i].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); bad = 1.0; break; }
      e->peer_buf[i] = (double *)ptr;
      e->peer_flags[i] = (unsigned int *)((unsigned char *)ptr + flag_off);
    }
  } else bad = 1.0;
  // agree: one failure anywhere keeps everybody on the NCCL path
  double *bad_d = (double *)cards_d;
  CUDA_TRY(cudaMemcpyAsync(bad_d, &bad, sizeof bad, cudaMemcpyHostToDevice, e->stream));
  r = g_nccl.AllReduce(bad_d, bad_d, 1, kNcclFloat64, kNcclSum, e->comm, e->stream);
  if (r != 0) { cudaFree(cards_d); return fail(EXB_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); }
  CUDA_TRY(cudaMemcpyAsync(&bad, bad_d, sizeof bad, cudaMemcpyDeviceToHost, e->stream));
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  cudaFree(cards_d);
  if (bad != 0.0) {
    for (int i = 0; i < n; i++) {
      if (i != e->rank && e->peer_buf[i]) cudaIpcCloseMemHandle(e->peer_buf[i]);
      e->peer_buf[i] = nullptr; e->peer_flags[i] = nullptr;
    }
    return EXB_OK;
  }
  e->p2p = true;
  // one exchange now: maps the peers' pages (IPC mappings are established lazily) outside anybody's timed region
  e->xseq = 1;
  launch_p2p_warmup(p2p_args(e), e->xseq, 2 * e->P, e->stream);
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  return EXB_OK;
}

extern "C" int exb_comm_init(exb_engine *e, int rank, int nranks, const void *id)
{
  if (!e) return fail(EXB_ERR_ARG, "exb_comm_init: null engine");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(EXB_ERR_ARG, "exb_comm_init: rank %d of %d", rank, nranks);
  e->nranks = nranks;
  if (nranks == 1) return EXB_OK;
  int rc = load_nccl();
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(e->device));
  NcclUniqueId uid;
  memcpy(&uid, id, sizeof uid);
  const int r = g_nccl.CommInitRank(&e->comm, nranks, uid, rank);
  if (r != 0) return fail(EXB_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  e->rank = rank;
  return setup_p2p(e);
}

extern "C" int exb_comm_is_p2p(exb_engine *e) { return e && e->p2p ? 1 : 0; }

// ---- introspection ---------------------------------------------------------------------------------------
extern "C" int exb_read_clv(exb_engine *e, int model, int nodeNumber, double *out)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (model < 0 || model >= e->P || nodeNumber <= e->mxtips || nodeNumber > 2 * e->mxtips - 2)
    return fail(EXB_ERR_ARG, "exb_read_clv: bad partition/node");
  HostPart &p = e->parts[model];
  ClassState &k = e->cls[p.cls];
  if (p.width == 0) return EXB_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  if (k.sev.on) {                                  // -S: tile by tile -- a block, or the node's gap column for every site
    const size_t row = (size_t)(nodeNumber - e->mxtips - 1);
    std::vector<double> gap((size_t)k.span);
    CUDA_TRY(cudaMemcpy(gap.data(), k.sev.gapcol + (row * k.nparts + p.part_cls) * k.span, k.span * sizeof(double), cudaMemcpyDeviceToHost));
    for (int t = k.cp_tile_begin[p.part_cls]; t < k.cp_tile_begin[p.part_cls] + k.cp_tile_count[p.part_cls]; t++) {
      const Tile &tl = k.tiles_h[t];
      double *dst = out + (size_t)(tl.site0 - p.site0) * k.span;
      const double *blk = k.sev.table_h[row * k.sev.ntiles + t];
      if (blk) CUDA_TRY(cudaMemcpy(dst, blk, (size_t)tl.nsites * k.span * sizeof(double), cudaMemcpyDeviceToHost));
      else for (uint32_t i = 0; i < tl.nsites; i++) memcpy(dst + (size_t)i * k.span, gap.data(), k.span * sizeof(double));
    }
    return EXB_OK;
  }
  const double *src = k.clv + ((size_t)(nodeNumber - e->mxtips - 1) * k.N + p.site0) * k.span;
  CUDA_TRY(cudaMemcpy(out, src, p.width * k.span * sizeof(double), cudaMemcpyDeviceToHost));
  return EXB_OK;
}

extern "C" int exb_read_scaler(exb_engine *e, int model, int nodeNumber, unsigned int *out)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (model < 0 || model >= e->P || nodeNumber < 1 || nodeNumber > 2 * e->mxtips - 2)
    return fail(EXB_ERR_ARG, "exb_read_scaler: bad partition/node");
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaMemcpy(out, e->gs + (size_t)model * e->gs_stride + nodeNumber, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  return EXB_OK;
}

extern "C" int exb_read_sumtable(exb_engine *e, int model, double *out)
{
  if (!e || !e->committed) return fail(EXB_ERR_STATE, "engine not committed");
  if (model < 0 || model >= e->P) return fail(EXB_ERR_ARG, "exb_read_sumtable: bad partition");
  HostPart &p = e->parts[model];
  ClassState &k = e->cls[p.cls];
  if (p.width == 0) return EXB_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  CUDA_TRY(cudaMemcpy(out, k.sum + p.site0 * k.span, p.width * k.span * sizeof(double), cudaMemcpyDeviceToHost));
  return EXB_OK;
}

extern "C" int exb_last_kernel_ms(exb_engine *e, float *ms)
{
  if (!e || !ms) return fail(EXB_ERR_ARG, "exb_last_kernel_ms: null argument");
  *ms = e->last_ms;
  return EXB_OK;
}

extern "C" int exb_last_phase_ms(exb_engine *e, float *ms3)
{
  if (!e || !ms3) return fail(EXB_ERR_ARG, "exb_last_phase_ms: null argument");
  if (!e->phases_valid) return fail(EXB_ERR_STATE, "exb_last_phase_ms: the last call was not a timed traversal call");
  CUDA_TRY(cudaEventSynchronize(e->ev1));
  cudaEventElapsedTime(&ms3[0], e->ev0, e->ev_a);
  cudaEventElapsedTime(&ms3[1], e->ev_a, e->ev_b);
  cudaEventElapsedTime(&ms3[2], e->ev_b, e->ev1);
  return EXB_OK;
}

extern "C" long exb_launch_count(exb_engine *e) { return e ? e->launches : 0; }

extern "C" int exb_set_timing(exb_engine *e, int on)
{
  if (!e) return fail(EXB_ERR_ARG, "exb_set_timing: null engine");
  e->timing = on != 0;
  return EXB_OK;
}
