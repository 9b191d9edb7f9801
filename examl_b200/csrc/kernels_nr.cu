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
__global__ void __launch_bounds__(kThreads, 2)
newton_raphson_kernel(const NrArgs a)
{
  __shared__ union { CoreSmem g; CoreCatSmem c; double red[32]; } shu;
  __shared__ NrLocal L;
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
    for (int c = 0; c < 2; c++) {
      const int nt = a.ntiles[c];
      for (int t = blockIdx.x; t < nt; t += gridDim.x) {
        const Tile tile = a.tiles[c][t];
        if (!a.mask[tile.part]) continue;
        double t1, t2;
        const double lz = st->lz[nb > 1 ? tile.part : 0];
        if (a.rate_cat[c]) {
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


// ------------------------------------------------------------------------------------------------------
// The same loop for a handful of tiles (alignments on which the latency of a call is everything): ONE CTA, an ordinary
// launch, and nothing but shared memory inside the loop.  Prologue: the sum-table rows and weights of every partition
// (contiguous on the class's site axis), the heads of the model blocks (EIGN, rates, LG4X weights), the partition tables
// and the start values (kernel parameters) are brought on chip in one wave of independent loads.  Per iteration: the
// exponential tables of all partitions, one sweep per partition over its rows (same per-site arithmetic as core_tile),
// one block reduction per partition, the state machine -- a few microseconds, no global round trip.  GAMMA, at most
// kNrInitSlots partitions / branch slots.  Several ranks: the per-iteration exchange goes through part_out as above.
// ------------------------------------------------------------------------------------------------------
namespace {
struct SmallTables { double d0[kNrInitSlots][80], d1[kNrInitSlots][80], d2[kNrInitSlots][80], wc[kNrInitSlots][4]; };

// first and second derivative sums of one partition over its staged rows; result valid in thread 0.
// Thread = site (no shuffles, one division per site and pass); rows are staged with a stride of SPAN + 1 doubles so that
// the threads of a warp read different banks.
template <int S>
__device__ __forceinline__ void small_sweep(const double *rows, const int *wgt, unsigned int nsites, const double *d0,
                                            const double *d1, const double *d2, const double *wcat, double *red,
                                            double &out1, double &out2)
{
  constexpr int SPAN = 4 * S, STRIDE = SPAN + 1;
  double acc1 = 0.0, acc2 = 0.0;
  for (unsigned int i = threadIdx.x; i < nsites; i += kThreads) {
    const double *p = rows + (size_t)i * STRIDE;
    double L = 0.0, L1 = 0.0, L2 = 0.0;
    #pragma unroll 1
    for (int c = 0; c < 4; c++) {
      double l = 0.0, l1 = 0.0, l2 = 0.0;
      #pragma unroll 4
      for (int k = 0; k < S; k++) {
        const double t = d0[c * S + k] * p[c * S + k];
        l += t; l1 += t * d1[c * S + k]; l2 += t * d2[c * S + k];
      }
      if (S == 20) { const double wc = wcat[c]; l *= wc; l1 *= wc; l2 *= wc; }     // LG4X category weights (1.0 otherwise)
      L += l; L1 += l1; L2 += l2;
    }
    const double inv = 1.0 / fabs(L);
    const double g = L1 * inv, h = L2 * inv, w = (double)wgt[i];
    acc1 += w * g;
    acc2 += w * (h - g * g);
  }
  out1 = block_sum(acc1, red);
  out2 = block_sum(acc2, red);
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
newton_raphson_small_kernel(const NrArgs a)
{
  extern __shared__ __align__(16) unsigned char nr_dyn[];     // per partition: [width][span + 1] doubles, then [width] ints (16-byte padded)
  __shared__ SmallTables T;
  __shared__ double red[32];
  __shared__ double smodel[kNrInitSlots][88];
  __shared__ double s_part[2 * kNrInitSlots];
  __shared__ size_t s_rows[kNrInitSlots], s_wgt[kNrInitSlots];
  __shared__ int s_cls[kNrInitSlots], s_has[kNrInitSlots];
  __shared__ unsigned char s_mask[kNrInitSlots];
  __shared__ struct {
    double z[kNrInitSlots], z0[kNrInitSlots], zprev[kNrInitSlots], zstep[kNrInitSlots], lz[kNrInitSlots];
    int budget[kNrInitSlots];
    unsigned char done[kNrInitSlots], curvatOK[kNrInitSlots];
    int all_done, iterations, peer_lost;
    unsigned int xseq;
  } L;
  const int tid = threadIdx.x;
  const int nb = a.nbranches, P = a.nparts;

  // ---- prologue ----
  if (tid == 0) {
    size_t off = 0;
    for (int m = 0; m < P; m++) {
      const int c = a.small_cls[m];
      s_cls[m] = c; s_has[m] = a.small_width[m] != 0;
      s_rows[m] = off; off += (((size_t)a.small_width[m] * (c == CLS_DNA ? 17 : 81) * sizeof(double)) + 15) & ~(size_t)15;
      s_wgt[m] = off;  off += ((size_t)a.small_width[m] * sizeof(int) + 15) & ~(size_t)15;
    }
    L.all_done = 0; L.iterations = 0; L.peer_lost = 0; L.xseq = a.xseq0;
  }
  if (tid < nb) {
    L.z[tid] = a.z0[tid]; L.z0[tid] = a.z0[tid]; L.budget[tid] = a.maxiter; L.done[tid] = 0; L.curvatOK[tid] = 1;
    L.zprev[tid] = 0.0; L.zstep[tid] = 0.0; L.lz[tid] = 0.0;
  }
  if (tid < P) s_mask[tid] = a.mask0[tid];
  __syncthreads();
  for (int m = 0; m < P; m++) {
    // (all addresses come from the kernel parameters: one wave of independent loads)
    const int c = a.small_cls[m], span = c == CLS_DNA ? 16 : 80;
    const double2 *src = reinterpret_cast<const double2 *>(a.sum[c] + (size_t)a.small_site0[m] * span);
    double *dst = reinterpret_cast<double *>(nr_dyn + s_rows[m]);
    const int n2 = (int)a.small_width[m] * span / 2;
    #pragma unroll 8
    for (int i = tid; i < n2; i += kThreads) {
      const double2 v = src[i];
      const int e = 2 * i, site = e / span, k = e % span;            // row stride span + 1: bank-conflict-free thread-per-site reads
      dst[(size_t)site * (span + 1) + k] = v.x;
      dst[(size_t)site * (span + 1) + k + 1] = v.y;
    }
    int *wd = reinterpret_cast<int *>(nr_dyn + s_wgt[m]);
    for (int i = tid; i < (int)a.small_width[m]; i += kThreads) wd[i] = a.wgt[c][a.small_site0[m] + i];
    const double *mb = a.models[c] + (size_t)a.small_pc[m] * (c == CLS_DNA ? kDnaModelDoubles : kAaModelDoubles);
    for (int i = tid; i < (c == CLS_DNA ? 8 : 88); i += kThreads) smodel[m][i] = mb[i];
  }
  __syncthreads();
  // d1 = EIGN * rate, d2 = d1^2 (coreGTRGAMMA makenewzGenericSpecial.c:2330-2345) and the category weights: once
  for (int i = tid; i < P * 80; i += kThreads) {
    const int m = i / 80, w = i % 80, S = s_cls[m] == CLS_DNA ? 4 : 20;
    if (w >= 4 * S) continue;
    const int cc = w / S, k = w % S;
    const double ki = S == 20 ? m_rate<20>(smodel[m], cc) : m_rate<4>(smodel[m], cc);
    const double ek = S == 20 ? m_eign<20>(smodel[m], cc, k) : m_eign<4>(smodel[m], cc, k);
    T.d1[m][w] = k ? ek * ki : 0.0;
    T.d2[m][w] = k ? ek * ek * (ki * ki) : 0.0;
    if (w < 4) T.wc[m][w] = S == 20 ? m_weight<20>(smodel[m], w) : 1.0;
  }

  // top of the do-loop in topLevelMakenewz (makenewzGenericSpecial.c:1161-1230)
  auto prelude = [&]() {
    if (tid < nb) {
      const int i = tid;
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
    }
    __syncthreads();
    if (tid < P) {
      if (nb > 1) { if (s_mask[tid]) s_mask[tid] = !L.curvatOK[tid]; }
      else s_mask[tid] = !L.curvatOK[0];
    }
    __syncthreads();
  };
  __syncthreads();
  prelude();

  while (!L.all_done) {
    // exponentials of every executed partition at its slot's trial lz: d0 = exp(EIGN * rate * lz), same association as core_tile
    for (int i = tid; i < P * 80; i += kThreads) {
      const int m = i / 80, w = i % 80, S = s_cls[m] == CLS_DNA ? 4 : 20;
      if (w >= 4 * S || !s_mask[m]) continue;
      const int cc = w / S, k = w % S;
      const double ki = S == 20 ? m_rate<20>(smodel[m], cc) : m_rate<4>(smodel[m], cc);
      const double ek = S == 20 ? m_eign<20>(smodel[m], cc, k) : m_eign<4>(smodel[m], cc, k);
      T.d0[m][w] = k ? exp(ek * ki * L.lz[nb > 1 ? m : 0]) : 1.0;
    }
    __syncthreads();
    for (int m = 0; m < P; m++) {
      double s0 = 0.0, s1 = 0.0;
      if (s_mask[m] && s_has[m]) {                   // block-uniform
        const double *rows = reinterpret_cast<const double *>(nr_dyn + s_rows[m]);
        const int *wg = reinterpret_cast<const int *>(nr_dyn + s_wgt[m]);
        if (s_cls[m] == CLS_DNA) small_sweep<4>(rows, wg, a.small_width[m], T.d0[m], T.d1[m], T.d2[m], T.wc[m], red, s0, s1);
        else                     small_sweep<20>(rows, wg, a.small_width[m], T.d0[m], T.d1[m], T.d2[m], T.wc[m], red, s0, s1);
      }
      if (tid == 0) { s_part[m] = s0; s_part[P + m] = s1; }
    }
    __syncthreads();
    bool peers_ok = true;
    if (a.x.nranks > 1) {
      const unsigned int xs = a.xseq0 + 1u + (unsigned int)L.iterations;
      const size_t row = p2p_row(a.x, xs, a.x.rank, 2 * P);
      for (int r = 0; r < a.x.nranks; r++)
        for (int i = tid; i < 2 * P; i += kThreads) a.x.peer_buf[r][row + i] = s_part[i];
      peers_ok = p2p_finish_exchange(a.x, xs, 2 * P, 2 * P, a.part_out, nullptr);
      for (int i = tid; i < 2 * P; i += kThreads) s_part[i] = a.part_out[i];
      if (tid == 0) L.xseq = xs;
      __syncthreads();
    }
    if (tid == 0 && !peers_ok) { L.peer_lost = 1; L.iterations += 1; L.all_done = 1; }
    if (tid == 0 && peers_ok) {
      int all = 1;
      for (int i = 0; i < nb; i++) {                 // partitions -> slots in partition order (makenewzGenericSpecial.c:921-946)
        double d1 = 0.0, d2 = 0.0;
        if (nb > 1) { if (s_mask[i]) { d1 += s_part[i]; d2 += s_part[P + i]; } }
        else for (int m = 0; m < P; m++) { if (!s_mask[m]) continue; d1 += s_part[m]; d2 += s_part[P + m]; }
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
    prelude();
  }

  // ---- results into pinned host memory, then the flag ----
  if (tid < nb) { a.result_host->z[tid] = L.z[tid]; a.result_host->curvatOK[tid] = L.curvatOK[tid]; }
  if (tid == 0) { a.result_host->iterations = L.iterations; a.result_host->xseq = L.xseq; a.result_host->peer_lost = L.peer_lost; }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) *a.done_flag_host = a.seq;
}

int launch_newton_raphson(const NrArgs &a, int blocks, size_t staged_bytes, cudaStream_t s)
{
  if (blocks == 1 && staged_bytes) {               // one CTA, everything in shared memory
    static size_t configured[kMaxDevices];
    const cudaError_t r = ensure_dynamic_smem(newton_raphson_small_kernel, kNrStagedBytes, configured);
    if (r != cudaSuccess) return (int)r;
    newton_raphson_small_kernel<<<1, kThreads, staged_bytes, s>>>(a);
    return (int)cudaGetLastError();
  }
  if (blocks == 1) {                               // no grid barrier inside: an ordinary launch
    newton_raphson_kernel<<<1, kThreads, 0, s>>>(a);
    return (int)cudaGetLastError();
  }
  void *params[] = {(void *)&a};
  return (int)cudaLaunchCooperativeKernel((const void *)newton_raphson_kernel, dim3(blocks), dim3(kThreads), params, 0, s);
}

int newton_raphson_max_blocks(int device)
{
  int per_sm = 0, sms = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, newton_raphson_kernel, kThreads, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return per_sm * sms;
}

}  // namespace exb
