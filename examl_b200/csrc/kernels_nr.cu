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

// monotonic-counter grid barrier; all CTAs are co-resident (cooperative launch)
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &generation)
{
  __syncthreads();
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

}  // namespace

__global__ void __launch_bounds__(kThreads, 2)
newton_raphson_kernel(const NrArgs a)
{
  __shared__ union { CoreSmem g; CoreCatSmem c; double red[32]; } shu;
  CoreSmem &sh = shu.g;
  NrState *st = a.state;
  unsigned int gen = 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nbranches, P = a.nparts;

  // ---- state-machine prelude (block 0): corresponds to the top of the do-loop in topLevelMakenewz ----
  for (;;) {
    if (blockIdx.x == 0) {
      for (int i = tid; i < nb; i += kThreads) {
        if (!st->done[i] && st->curvatOK[i]) {
          st->curvatOK[i] = 0;
          st->zprev[i] = st->z[i];
          st->zstep[i] = (1.0 - kZmaxD) * st->z[i] + kZmin;
        }
        if (!st->done[i] && !st->curvatOK[i]) {
          double z = st->z[i];
          if (z < kZmin) z = kZmin; else if (z > kZmaxD) z = kZmaxD;
          st->z[i] = z;
          st->lz[i] = log(z);
        }
      }
      __syncthreads();
      for (int m = tid; m < P; m += kThreads) {
        if (nb > 1) { if (a.mask[m]) a.mask[m] = !st->curvatOK[m]; }
        else a.mask[m] = !st->curvatOK[0];
      }
    }
    grid_barrier(a.barrier, gen);

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
    grid_barrier(a.barrier, gen);

    // ---- tile partials -> partitions (fixed order); the partitions are spread over all CTAs ----
    for (int part = blockIdx.x; part < P; part += gridDim.x) {
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
    grid_barrier(a.barrier, gen);

    // ---- block 0: (all-reduce over the ranks,) partitions -> slots, NR update ----
    if (blockIdx.x == 0) {
      bool peers_ok = true;
      if (a.x.nranks > 1) {
        const unsigned int xs = a.xseq0 + 1u + (unsigned int)st->iterations;
        const size_t row = p2p_row(a.x, xs, a.x.rank, 2 * P);
        for (int r = 0; r < a.x.nranks; r++)
          for (int i = tid; i < 2 * P; i += kThreads) a.x.peer_buf[r][row + i] = a.part_out[i];
        peers_ok = p2p_finish_exchange(a.x, xs, 2 * P, 2 * P, a.part_out, nullptr);
        if (tid == 0) st->xseq = xs;
      }
      if (tid == 0 && !peers_ok) { st->peer_lost = 1; st->iterations += 1; st->all_done = 1; }
      if (tid == 0 && peers_ok) {
        // partitions -> slots in partition order (makenewzGenericSpecial.c:921-946)
        for (int i = 0; i < nb; i++) { st->d1[i] = 0.0; st->d2[i] = 0.0; }
        for (int m = 0; m < P; m++) {
          if (!a.mask[m]) continue;
          const int slot = nb > 1 ? m : 0;
          st->d1[slot] += a.part_out[m];
          st->d2[slot] += a.part_out[P + m];
        }
        int all = 1;
        for (int i = 0; i < nb; i++) {
          if (!st->done[i] && !st->curvatOK[i]) {
            if (st->d2[i] >= 0.0 && st->z[i] < kZmaxD)
              st->zprev[i] = st->z[i] = 0.37 * st->z[i] + 0.63;
            else
              st->curvatOK[i] = 1;
          }
          if (st->curvatOK[i] && !st->done[i]) {
            double z = st->z[i];
            if (st->d2[i] < 0.0) {
              const double step = -st->d1[i] / st->d2[i];
              const double cap = 0.25 * st->zprev[i] + 0.75;
              if (step < 100) {
                z *= exp(step);
                if (z < kZmin) z = kZmin;
                if (z > cap) z = cap;
              } else
                z = cap;
            }
            if (z > kZmaxD) z = kZmaxD;
            st->budget[i] -= 1;
            if (fabs(z - st->zprev[i]) > st->zstep[i]) {
              if (st->budget[i] < -20) { z = st->z0[i]; st->done[i] = 1; }
            } else
              st->done[i] = 1;
            st->z[i] = z;
          }
          all = all && st->done[i];
        }
        st->iterations += 1;
        st->all_done = all;
      }
    }
    grid_barrier(a.barrier, gen);
    if (*(volatile int *)&st->all_done) break;
  }
}

int launch_newton_raphson(const NrArgs &a, int blocks, cudaStream_t s)
{
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
