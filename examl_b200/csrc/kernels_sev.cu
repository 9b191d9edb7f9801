// kernels_sev.cu -- gap columns for the -S memory saving option (SEV, "subtree equality vectors").
//
// The reference stores, per inner node, only the sites that are not all-gap in the subtree below the node, plus ONE
// shared vector -- the "gap column" -- that every all-gap site equals (x3_gapColumn, newviewGenericSpecial.c:1160-1216;
// newviewGTRGAMMA_AVX_GAPPED_SAVE avxLikelihood.c:1806, ...PROT... :3125).  The engine applies the same idea at tile
// granularity: a (node, tile) pair whose sites are all gaps below the node owns no CLV block and is neither computed
// nor stored (engine.cu, stage.cuh); consumers read the node's gap column instead.  This kernel computes the gap
// columns of a descriptor's nodes in entry order -- the newview of a single all-gap site per (node, partition) -- and
// books the scaling events of the all-gap sites (the reference's scaleGap bookkeeping).
#include "engine.cuh"

namespace exb {

__global__ void __launch_bounds__(128)
gapcol_kernel(const GapcolArgs a)
{
  __shared__ double xq[80], xr[80], ab[80];
  const int pc = blockIdx.x, part = a.cls_part_global[pc];
  if (!a.mask[part]) return;
  const int S = a.states, span = 4 * S, tid = threadIdx.x;
  const double *model = a.models + (size_t)pc * a.model_doubles;
  const bool aa = S == 20;
  const int c = tid / S, m = tid % S;            // (category, state) of this thread, tid < span
  // model-block accessors (engine.cuh): DNA has one eigen-system, AA one per category
  const double *EI = model + (aa ? kAaEI + c * 400 : 8), *EV = model + (aa ? kAaEV + c * 400 : 24);
  const double *gaptip = model + (aa ? kAaTip + c * 460 + 22 * 20 : 40 + 15 * 4);     // tipVector[undetermined]
  const size_t node_stride = (size_t)a.nparts_cls * span;
  for (int e = a.first; e < a.count; e++) {
    const DevEntry en = a.entries[e];
    const double *drow = a.dtab + ((size_t)e * a.nparts_cls + pc) * a.dtab_row;     // [c][side][k]
    if (tid < span) {
      xq[tid] = en.q <= a.mxtips ? gaptip[m] : a.gapcol[(size_t)(en.q - a.mxtips - 1) * node_stride + (size_t)pc * span + tid];
      xr[tid] = en.r <= a.mxtips ? gaptip[m] : a.gapcol[(size_t)(en.r - a.mxtips - 1) * node_stride + (size_t)pc * span + tid];
    }
    __syncthreads();
    if (tid < span) {                            // a[l] = sum_j P[l][j] x[j], P[l][j] = EI[l][j] d[j], P[l][0] = 1; l = m
      const double *dq = drow + (size_t)c * 2 * S, *dr = dq + S;
      double aq = xq[c * S], ar = xr[c * S];
      for (int j = 1; j < S; j++) {
        aq = fma(EI[m * S + j], dq[j] * xq[c * S + j], aq);
        ar = fma(EI[m * S + j], dr[j] * xr[c * S + j], ar);
      }
      ab[tid] = aq * ar;
    }
    __syncthreads();
    double v = 0.0;
    if (tid < span) {
      v = ab[c * S] * EV[m];
      for (int l = 1; l < S; l++) v = fma(ab[c * S + l], EV[l * S + m], v);
    }
    const int tiny = __syncthreads_and(tid >= span || fabs(v) < kMinLikelihood);
    if (en.tipCase != 0 && tiny) {               // TIP_TIP is never rescaled
      v *= kTwoToThe256;
      const int w = a.gapw[(size_t)e * a.nparts_cls + pc];
      if (tid == 0 && w) atomicAdd(&a.scale_inc[(size_t)e * a.nparts + part], (unsigned int)w);
    }
    if (tid < span) a.gapcol[(size_t)(en.p - a.mxtips - 1) * node_stride + (size_t)pc * span + tid] = v;
    __syncthreads();                             // the next entry may read this gap column
  }
}

void launch_gapcol(const GapcolArgs &a, cudaStream_t s)
{
  if (a.nparts_cls > 0 && a.count > a.first) gapcol_kernel<<<a.nparts_cls, 128, 0, s>>>(a);
}

}  // namespace exb
