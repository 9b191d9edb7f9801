// finish.cuh -- what ends every traversal call: (1) scaler bookkeeping globalScaler[p] = globalScaler[q] + globalScaler[r]
// + increment (newviewGenericSpecial.c:1503-1506) for every newview entry in order; (2) deterministic reduction of the
// per-tile partial sums of each partition; (3) evaluate only: undo the scaling, + (gs[p] + gs[q]) * log(minlikelihood)
// (evaluateGenericSpecial.c:830); (4) results + completion flag into pinned host memory (single rank), or the all-reduce
// over the peer ranks first (p2p.cuh).
// Used by finish_kernel (engine.cu: one block per partition, its own launch) and, for the short descriptors of a tree
// search on a small alignment, by the last CTA of the traverse kernel itself (kernels_dna.cu, SMALL instantiations): one
// launch per call instead of two.
#pragma once
#include "engine.cuh"
#include "p2p.cuh"

namespace exb {

// Partitions part0, part0 + stride, ...  All threads of the CTA must call (barriers inside); the first 128 work.
// gsm: shared memory for one scaler row (2 * mxtips words) when a.gs_in_smem; red0 / red1: 4 doubles of shared memory each.
// entries / mask: the call's descriptor and execute mask (a.entries / a.mask, or the fused kernel's shared-memory copies)
__device__ __forceinline__ void finish_body(const FinishArgs &a, const DevEntry *entries, const unsigned char *mask,
                                            unsigned int *gsm, double *red0, double *red1, int part0, int part_stride)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool act = tid < 128;
  for (int part = part0; part < a.nparts; part += part_stride) {
    const bool exec = mask[part] && a.part_has_sites[part];
    unsigned int *gs = a.gs + (size_t)part * a.gs_stride;
    // ---- (1) scaler chain, sequential in the entries: warp 0, in shared memory when the row fits ----
    if (exec && a.count > a.first && warp == 0) {
      if (a.gs_in_smem) {
        for (int i = lane; i < 2 * a.mxtips; i += 32) gsm[i] = gs[i];
        __syncwarp();
        for (int e0 = a.first; e0 < a.count; e0 += 32) {
          const int e = e0 + lane;
          unsigned int inc = 0; DevEntry en = {0, 0, 0, 0};
          if (e < a.count) {
            en = entries[e];
            inc = a.scale_inc[(size_t)e * a.nparts + part];
            if (inc) a.scale_inc[(size_t)e * a.nparts + part] = 0;
          }
          const int n = min(32, a.count - e0);
          for (int j = 0; j < n; j++) {               // lane j's entry, in order
            const int pj = __shfl_sync(0xffffffffu, en.p, j), qj = __shfl_sync(0xffffffffu, en.q, j),
                      rj = __shfl_sync(0xffffffffu, en.r, j);
            const unsigned int ij = __shfl_sync(0xffffffffu, inc, j);
            if (lane == 0) gsm[pj] = gsm[qj] + gsm[rj] + ij;
            __syncwarp();
          }
        }
        for (int i = lane; i < 2 * a.mxtips; i += 32) gs[i] = gsm[i];
      } else if (lane == 0) {
        for (int e = a.first; e < a.count; e++) {
          const DevEntry en = entries[e];
          unsigned int *inc = &a.scale_inc[(size_t)e * a.nparts + part];
          gs[en.p] = gs[en.q] + gs[en.r] + *inc;
          *inc = 0;
        }
      }
    }
    __syncthreads();                                 // gs[] of this partition is final; red[] is free
    if (a.mode == MODE_NEWVIEW || a.mode == MODE_SUM) continue;
    // ---- (2) tile partials -> partition, fixed order: thread-strided sums (loads in flight), warp tree, 4 warps ----
    double s0 = 0.0, s1 = 0.0;
    if (exec && act) {
      const double2 *to = reinterpret_cast<const double2 *>((a.part_cls[part] ? a.tile_out[1] : a.tile_out[0]) +
                                                             2 * (size_t)a.part_tile_begin[part]);
      const int n = a.part_tile_count[part];
      int t = tid;
      for (; t + 3 * 128 < n; t += 4 * 128) {
        const double2 v0 = to[t], v1 = to[t + 128], v2 = to[t + 256], v3 = to[t + 384];
        s0 += v0.x; s1 += v0.y; s0 += v1.x; s1 += v1.y; s0 += v2.x; s1 += v2.y; s0 += v3.x; s1 += v3.y;
      }
      for (; t < n; t += 128) { const double2 v = to[t]; s0 += v.x; s1 += v.y; }
      #pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_down_sync(0xffffffffu, s0, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
      }
    }
    if (lane == 0 && act) { red0[warp] = s0; red1[warp] = s1; }
    __syncthreads();
    if (tid == 0) {
      s0 = ((red0[0] + red0[1]) + (red0[2] + red0[3]));
      s1 = ((red1[0] + red1[1]) + (red1[2] + red1[3]));
      if (a.mode == MODE_EVALUATE && exec) {
        const DevEntry br = entries[0];
        s0 += (double)(gs[br.p] + gs[br.q]) * a.log_minlikelihood;   // evaluateGenericSpecial.c:830
      }
      if (a.x.nranks > 1) {        // my contribution goes straight into every rank's receive buffer
        const size_t at = p2p_row(a.x, a.xseq, a.x.rank, 2 * a.nparts) + part;
        for (int r = 0; r < a.x.nranks; r++) {
          a.x.peer_buf[r][at] = s0;
          if (a.mode == MODE_CORE) a.x.peer_buf[r][at + a.nparts] = s1;
        }
      } else {
        a.out[part] = s0;
        if (a.out_host) a.out_host[part] = s0;
        if (a.mode == MODE_CORE) {
          a.out[a.nparts + part] = s1;
          if (a.out_host) a.out_host[a.nparts + part] = s1;
        }
      }
    }
    __syncthreads();
  }
}

// The CTA that knows everybody's results are written (the last to arrive, or the only one): exchange with the peer ranks
// when there are any, then publish the completion flag the host polls.  All threads of that CTA call.
__device__ __forceinline__ void finish_publish(const FinishArgs &a, int *peer_lost_smem)
{
  const int tid = threadIdx.x;
  if (tid == 0) *peer_lost_smem = 0;
  __syncthreads();
  if (a.x.nranks > 1 && a.mode != MODE_NEWVIEW && a.mode != MODE_SUM) {
    const bool ok = p2p_finish_exchange(a.x, a.xseq, (a.mode == MODE_CORE ? 2 : 1) * a.nparts, 2 * a.nparts, a.out, a.out_host);
    if (tid == 0 && !ok) *peer_lost_smem = 1;
    __threadfence_system();
    __syncthreads();
  }
  if (tid == 0) {
    *a.ticket = 0;
    __threadfence_system();
    *a.done_flag_host = *peer_lost_smem ? 0xFFFFFFFFu : a.seq;       // 0xFFFFFFFF: the host reports the lost peer
  }
}

}  // namespace exb
