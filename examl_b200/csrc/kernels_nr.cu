// kernels_nr.cu -- the Newton-Raphson loop of makenewzGeneric resident on the device.
//
// The reference iterates on the host: per iteration one pass over the sum table (execCore,
// makenewzGenericSpecial.c:849), one all-reduce, one scalar update of z (topLevelMakenewz :1133-1349).  On a GPU each
// of those iterations costs a launch, a device-to-host copy and a synchronisation -- more than the 128 B/site pass
// itself.  Here ONE cooperative kernel runs the whole loop: persistent CTAs sweep the tiles (core_tile), a grid
// barrier, block 0 reduces tile partials per partition in fixed order and advances the per-branch-slot state machine
// (identical to the host code: curvature guard z = 0.37 z + 0.63, step z *= exp(-d1/d2) with the 0.25 zprev + 0.75
// cap, |z - zprev| <= zstep, the -20 give-up), another barrier, until every slot has converged.
// Several ranks on one node: after the per-partition reduction block 0 stores this rank's 2P sums into every peer's
// receive buffer (CUDA-IPC over NVLink, p2p.cuh), waits for the peers' rows and adds them in rank order -- every rank
// then advances an identical state machine with identical inputs, so all ranks leave the loop in the same iteration.
// No host round trip per iteration on any rank (the reference: one MPI_Allreduce per iteration, :1241-1248).
#include "engine.cuh"
#include "core_tile.cuh"
#include "p2p.cuh"

namespace exb {

namespace {

constexpr double kZmaxD = 1.0 - 1.0E-6;          // axml.h:111
constexpr int kSmallP = 8;                       // up to this many partitions block 0 reduces the tile partials itself

// monotonic-counter grid barrier; all CTAs are co-resident (cooperative launch).  A single CTA needs no counter.
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &generation)
{
  __syncthreads();
  if (gridDim.x == 1) return;
  if (threadIdx.x == 0) {
    generation += 1;
    const unsigned int target = generation * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

// Working copy of the state machine in block 0's shared memory (the global NrState is only the way in and out)
struct NrLocal {
  double z[kMaxBranchSlots], z0[kMaxBranchSlots], zprev[kMaxBranchSlots], zstep[kMaxBranchSlots], lz[kMaxBranchSlots];
  int budget[kMaxBranchSlots];
  unsigned char done[kMaxBranchSlots], curvatOK[kMaxBranchSlots];
  int all_done, iterations, peer_lost;
  unsigned int xseq;
};

}  // namespace

// Structure of one iteration (block 0 owns the state machine; its scalars live in shared memory):
//     [barrier A]  everybody sees lz / mask / all_done published by block 0
//     sweep        every CTA: core_tile over its share of the tiles -> tile_out
//     [barrier B]
//     reduce       tile partials -> partitions, fixed order (block 0 alone for <= kSmallP partitions, else spread over the
//                  CTAs + [barrier C])
//     block 0      (exchange with the peer ranks,) partitions -> slots, Newton-Raphson update, the prelude of the next
//                  iteration (topLevelMakenewz's loop top), publish
// A launch of ONE CTA (a few tiles: the alignments on which the per-call latency is everything) has no grid barrier at all.
// STAGED (one CTA, GAMMA, the sum table of all tiles fits in shared memory): the rows are brought on chip once and every
// iteration reads them there -- the loop then costs microseconds per iteration instead of an L2 round trip per pass.
template <bool STAGED>
__global__ void __launch_bounds__(kThreads, STAGED ? 1 : 2)
newton_raphson_kernel(const NrArgs a)
{
  __shared__ union { CoreSmem g; CoreCatSmem c; double red[32]; } shu;
  __shared__ NrLocal L;
  extern __shared__ __align__(16) unsigned char nr_dyn[];     // STAGED: per tile [nsites][span] doubles, then [nsites] ints
  __shared__ double smodel[STAGED ? kNrInitSlots : 1][88];    // STAGED (<= kNrInitSlots partitions): head of every model block
  CoreSmem &sh = shu.g;
  NrState *st = a.state;
  unsigned int gen = 0;
  const int tid = threadIdx.x;
  const int nb = a.nbranches, P = a.nparts;
  const bool lead = blockIdx.x == 0;

  // top of the do-loop in topLevelMakenewz (makenewzGenericSpecial.c:1161-1230) for the slots still running, then publish
  auto prelude_and_publish = [&]() {
    for (int i = tid; i < nb; i += kThreads) {
      if (!L.done[i] && L.curvatOK[i]) {
        L.curvatOK[i] = 0;
        L.zprev[i] = L.z[i];
        L.zstep[i] = (1.0 - kZmaxD) * L.z[i] + kZmin;
      }
      if (!L.done[i] && !L.curvatOK[i]) {
        double z = L.z[i];
        if (z < kZmin) z = kZmin; else if (z > kZmaxD) z = kZmaxD;
        L.z[i] = z;
        L.lz[i] = log(z);
      }
      st->lz[i] = L.lz[i];
    }
    __syncthreads();
    for (int m = tid; m < P; m += kThreads) {
      if (nb > 1) { if (a.mask[m]) a.mask[m] = !L.curvatOK[m]; }
      else a.mask[m] = !L.curvatOK[0];
    }
    if (tid == 0) st->all_done = L.all_done;
  };

  if (STAGED) {
    size_t off = 0;
    for (int c = 0; c < 2; c++)
      for (int t = 0; t < a.ntiles[c]; t++) {
        const Tile tile = a.tiles[c][t];
        const int span = c == CLS_DNA ? 16 : 80;
        const double2 *src = reinterpret_cast<const double2 *>(a.sum[c] + (size_t)tile.site0 * span);
        double2 *dst = reinterpret_cast<double2 *>(nr_dyn + off);
        const int n2 = (int)tile.nsites * span / 2;
        for (int i = tid; i < n2; i += kThreads) dst[i] = src[i];
        off += (size_t)tile.nsites * span * sizeof(double);
        int *wd = reinterpret_cast<int *>(nr_dyn + off);
        for (int i = tid; i < (int)tile.nsites; i += kThreads) wd[i] = a.wgt[c][tile.site0 + i];
        off += (((size_t)tile.nsites * sizeof(int)) + 15) & ~(size_t)15;
      }
    // the few model constants core_tile reads (EIGN, rates, LG4X weights: the head of a model block), per partition
    for (int c = 0; c < 2; c++)
      for (int t = 0; t < a.ntiles[c]; t++) {
        const Tile tile = a.tiles[c][t];
        const double *m = a.models[c] + (size_t)tile.part_cls * (c == CLS_DNA ? kDnaModelDoubles : kAaModelDoubles);
        for (int i = tid; i < (c == CLS_DNA ? 8 : 88); i += kThreads) smodel[tile.part][i] = m[i];
      }
    // (the first grid_barrier below orders these writes before the first sweep)
  }
  if (lead) {
    for (int i = tid; i < nb; i += kThreads) {
      if (a.init_from_args) {
        L.z[i] = a.z0[i]; L.z0[i] = a.z0[i]; L.budget[i] = a.maxiter; L.done[i] = 0; L.curvatOK[i] = 1;
        L.zprev[i] = 0.0; L.zstep[i] = 0.0; L.lz[i] = 0.0;
      } else {
        L.z[i] = st->z[i]; L.z0[i] = st->z0[i]; L.zprev[i] = st->zprev[i]; L.zstep[i] = st->zstep[i]; L.lz[i] = st->lz[i];
        L.budget[i] = st->budget[i]; L.done[i] = st->done[i]; L.curvatOK[i] = st->curvatOK[i];
      }
    }
    if (a.init_from_args) for (int m = tid; m < P; m += kThreads) a.mask[m] = a.mask0[m];
    if (tid == 0) { L.all_done = 0; L.iterations = 0; L.peer_lost = 0; L.xseq = a.xseq0; }
    __syncthreads();
    prelude_and_publish();
  }

  for (;;) {
    grid_barrier(a.barrier, gen);                                        // A
    if (*(volatile int *)&st->all_done) break;

    // ---- derivatives: every CTA sweeps its share of the tiles of both classes ----
    size_t soff = 0;
    for (int c = 0; c < 2; c++) {
      const int nt = a.ntiles[c];
      for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const Tile tile = a.tiles[c][t];
        const size_t rows = soff, wrow = soff + (size_t)tile.nsites * (c == CLS_DNA ? 16 : 80) * sizeof(double);
        if (STAGED) soff = wrow + ((((size_t)tile.nsites * sizeof(int)) + 15) & ~(size_t)15);
        if (!a.mask[tile.part]) continue;
        double t1, t2;
        const double lz = STAGED ? L.lz[nb > 1 ? tile.part : 0] : st->lz[nb > 1 ? tile.part : 0];
        if (STAGED) {
          const double *srows = reinterpret_cast<const double *>(nr_dyn + rows);
          const int *swgt = reinterpret_cast<const int *>(nr_dyn + wrow);
          if (c == CLS_DNA) core_tile<4, true>(tile, srows, swgt, smodel[tile.part], lz, sh, t1, t2);
          else              core_tile<20, true>(tile, srows, swgt, smodel[tile.part], lz, sh, t1, t2);
        } else if (a.rate_cat[c]) {
          const double *rates = a.cat_rates[c] + (size_t)tile.part_cls * a.max_cats;
          if (c == CLS_DNA) core_tile_cat<4>(tile, a.sum[c], a.wgt[c], a.models[c] + (size_t)tile.part_cls * kDnaModelDoubles, a.rate_cat[c], rates, a.max_cats, lz, shu.c, t1, t2);
          else              core_tile_cat<20>(tile, a.sum[c], a.wgt[c], a.models[c] + (size_t)tile.part_cls * kAaModelDoubles, a.rate_cat[c], rates, a.max_cats, lz, shu.c, t1, t2);
        } else if (c == CLS_DNA) core_tile<4>(tile, a.sum[c], a.wgt[c], a.models[c] + (size_t)tile.part_cls * kDnaModelDoubles, lz, sh, t1, t2);
        else                     core_tile<20>(tile, a.sum[c], a.wgt[c], a.models[c] + (size_t)tile.part_cls * kAaModelDoubles, lz, sh, t1, t2);
        if (tid == 0) { a.tile_out[c][2 * (size_t)t] = t1; a.tile_out[c][2 * (size_t)t + 1] = t2; }
      }
    }
    grid_barrier(a.barrier, gen);                                        // B

    // ---- tile partials -> partitions (fixed order) ----
    const bool small = P <= kSmallP;
    if (!small || lead) {
      for (int part = small ? 0 : (int)blockIdx.x; part < P; part += small ? 1 : (int)gridDim.x) {
        double s0 = 0.0, s1 = 0.0;
        if (a.mask[part] && a.part_has_sites[part]) {
          const double2 *to = reinterpret_cast<const double2 *>(a.tile_out[a.part_cls[part]] + 2 * (size_t)a.part_tile_begin[part]);
          const int n = a.part_tile_count[part];
          int t = tid;
          for (; t + 3 * kThreads < n; t += 4 * kThreads) {
            const double2 v0 = to[t], v1 = to[t + kThreads], v2 = to[t + 2 * kThreads], v3 = to[t + 3 * kThreads];
            s0 += v0.x; s1 += v0.y; s0 += v1.x; s1 += v1.y; s0 += v2.x; s1 += v2.y; s0 += v3.x; s1 += v3.y;
          }
          for (; t < n; t += kThreads) { const double2 v = to[t]; s0 += v.x; s1 += v.y; }
        }
        s0 = block_sum(s0, shu.g.red);
        s1 = block_sum(s1, shu.g.red);
        if (tid == 0) { a.part_out[part] = s0; a.part_out[P + part] = s1; }
      }
    }
    if (!small) grid_barrier(a.barrier, gen);                            // C
    else __syncthreads();

    // ---- block 0: (all-reduce over the ranks,) partitions -> slots, NR update, next prelude ----
    if (lead) {
      bool peers_ok = true;
      if (a.x.nranks > 1) {
        const unsigned int xs = a.xseq0 + 1u + (unsigned int)L.iterations;
        const size_t row = p2p_row(a.x, xs, a.x.rank, 2 * P);
        for (int r = 0; r < a.x.nranks; r++)
          for (int i = tid; i < 2 * P; i += kThreads) a.x.peer_buf[r][row + i] = a.part_out[i];
        peers_ok = p2p_finish_exchange(a.x, xs, 2 * P, 2 * P, a.part_out, nullptr);
        if (tid == 0) L.xseq = xs;
      }
      if (tid == 0 && !peers_ok) { L.peer_lost = 1; L.iterations += 1; L.all_done = 1; }
      if (tid == 0 && peers_ok) {
        // partitions -> slots in partition order (makenewzGenericSpecial.c:921-946)
        int all = 1;
        for (int i = 0; i < nb; i++) {
          double d1 = 0.0, d2 = 0.0;
          if (nb > 1) { if (a.mask[i]) { d1 += a.part_out[i]; d2 += a.part_out[P + i]; } }
          else for (int m = 0; m < P; m++) { if (!a.mask[m]) continue; d1 += a.part_out[m]; d2 += a.part_out[P + m]; }
          if (!L.done[i] && !L.curvatOK[i]) {
            if (d2 >= 0.0 && L.z[i] < kZmaxD)
              L.zprev[i] = L.z[i] = 0.37 * L.z[i] + 0.63;
            else
              L.curvatOK[i] = 1;
          }
          if (L.curvatOK[i] && !L.done[i]) {
            double z = L.z[i];
            if (d2 < 0.0) {
              const double step = -d1 / d2;
              const double cap = 0.25 * L.zprev[i] + 0.75;
              if (step < 100) {
                z *= exp(step);
                if (z < kZmin) z = kZmin;
                if (z > cap) z = cap;
              } else
                z = cap;
            }
            if (z > kZmaxD) z = kZmaxD;
            L.budget[i] -= 1;
            if (fabs(z - L.zprev[i]) > L.zstep[i]) {
              if (L.budget[i] < -20) { z = L.z0[i]; L.done[i] = 1; }
            } else
              L.done[i] = 1;
            L.z[i] = z;
          }
          all = all && L.done[i];
        }
        L.iterations += 1;
        L.all_done = all;
      }
      __syncthreads();
      prelude_and_publish();          // harmless after the last iteration: every slot is done
    }
  }

  // ---- results: the global state (tests, multi-CTA callers) and, when asked for, pinned host memory + flag ----
  if (lead) {
    for (int i = tid; i < nb; i += kThreads) {
      st->z[i] = L.z[i]; st->curvatOK[i] = L.curvatOK[i];
      if (a.result_host) { a.result_host->z[i] = L.z[i]; a.result_host->curvatOK[i] = L.curvatOK[i]; }
    }
    if (tid == 0) {
      st->iterations = L.iterations; st->xseq = L.xseq; st->peer_lost = L.peer_lost;
      if (a.result_host) { a.result_host->iterations = L.iterations; a.result_host->xseq = L.xseq; a.result_host->peer_lost = L.peer_lost; }
    }
    if (a.done_flag_host) {
      __threadfence_system();
      __syncthreads();
      if (tid == 0) *a.done_flag_host = a.seq;
    }
  }
}

int launch_newton_raphson(const NrArgs &a, int blocks, size_t staged_bytes, cudaStream_t s)
{
  if (blocks == 1 && staged_bytes) {               // one CTA, sum table in shared memory
    static size_t configured[kMaxDevices];
    const cudaError_t r = ensure_dynamic_smem(newton_raphson_kernel<true>, kNrStagedBytes, configured);
    if (r != cudaSuccess) return (int)r;
    newton_raphson_kernel<true><<<1, kThreads, staged_bytes, s>>>(a);
    return (int)cudaGetLastError();
  }
  if (blocks == 1) {                               // no grid barrier inside: an ordinary launch
    newton_raphson_kernel<false><<<1, kThreads, 0, s>>>(a);
    return (int)cudaGetLastError();
  }
  void *params[] = {(void *)&a};
  return (int)cudaLaunchCooperativeKernel((const void *)newton_raphson_kernel<false>, dim3(blocks), dim3(kThreads), params, 0, s);
}

int newton_raphson_max_blocks(int device)
{
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, newton_raphson_kernel<false>, kThreads, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

}  // namespace exb
