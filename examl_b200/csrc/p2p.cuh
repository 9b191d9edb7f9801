// p2p.cuh -- the all-reduce of evaluateGeneric / topLevelMakenewz (evaluateGenericSpecial.c:968-975,
// makenewzGenericSpecial.c:1241-1248) done by the kernels themselves when all ranks sit on one node.
//
// At exb_comm_init every rank allocates a small receive buffer and maps every peer's buffer through CUDA IPC
// (NVLink / NVSwitch).  An exchange number xseq counts the all-reduces of the run, identically on every rank:
//   1. a rank stores its contribution into slot (xseq & 1), row `rank`, of EVERY rank's buffer (its own included),
//   2. fences, then raises flag[rank] = xseq in every buffer (st.release.sys),
//   3. waits until all flags in its OWN buffer have reached xseq (ld.acquire.sys),
//   4. adds the nranks rows in rank order -- the same bits on every GPU.
// Two slots suffice: nobody can finish exchange k+1 before everybody has contributed to it, i.e. has finished
// reading exchange k.  A peer that never shows up is reported after ~20 s instead of hanging the GPU.
#pragma once
#include "engine.cuh"

namespace exb {

// receive buffer of one rank: [2 slots][nranks][stride] doubles, then [nranks] flags
__device__ __forceinline__ size_t p2p_row(const P2pArgs &x, unsigned int xseq, int rank, int stride)
{
  return ((size_t)(xseq & 1u) * x.nranks + rank) * stride;
}

// Steps 2-4 for ONE block (all its threads call; blockDim.x >= nranks).  Step 1 must already have been done by this
// GPU (any of its blocks) and be ordered before this call (e.g. __threadfence_system + a grid-level arrival count).
// Returns false when a peer was lost.  out / out_host may be null.
__device__ __forceinline__ bool p2p_finish_exchange(const P2pArgs &x, unsigned int xseq, int nvals, int stride,
                                                    double *out, double *out_host)
{
  __shared__ int lost;
  const int tid = threadIdx.x;
  if (tid == 0) lost = 0;
  __threadfence_system();                        // this GPU's remote stores are ordered before the flags
  __syncthreads();
  if (tid < x.nranks) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(x.peer_flags[tid] + x.rank), "r"(xseq) : "memory");
    const unsigned int *f = x.peer_flags[x.rank] + tid;       // rank tid's contribution arrives in MY buffer
    unsigned int v;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if ((int)(v - xseq) >= 0) break;
      if (clock64() - t0 > 40000000000LL) { lost = 1; break; }     // ~20 s: a peer is gone
      __nanosleep(100);
    }
  }
  __syncthreads();
  const double *mine = x.peer_buf[x.rank] + p2p_row(x, xseq, 0, stride);
  for (int i = tid; i < nvals; i += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < x.nranks; r++) t += mine[(size_t)r * stride + i];   // rank order: same bits everywhere
    if (out) out[i] = t;
    if (out_host) out_host[i] = t;
  }
  __syncthreads();
  return !lost;
}

}  // namespace exb
