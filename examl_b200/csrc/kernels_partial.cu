// kernels_partial.cu -- per-site likelihood at an explicit rate: the kernel behind exb_evaluate_partial, the batched
// form of evaluatePartialGeneric (evaluatePartialGenericSpecial.c:259) that optimizeRateCategories
// (optimizeModel.c:1792-1922) calls once per site and candidate rate under -m PSR.
//
// The reference recomputes the WHOLE traversal for one site from the tips (computeVectorGTRCAT :912,
// computeVectorGTRCATPROT :493) with a private lVector.  Here one thread owns one (site, rate) pair and walks the
// descriptor once; the handful of conditional vectors that are alive at any point of a post-order traversal live in
// a per-thread stack of "slots" (allocated on the host by liveness, see exb_evaluate_partial) laid out
// [slot][state][thread] so that every access is coalesced and the working set of the grid stays in L2.
// The exponentials depend on the pair's own rate, so they are computed in the thread (2*(S-1) per entry); that, not
// memory, bounds the kernel.
#include "engine.cuh"

namespace exb {

template <int S, int CODES>
__global__ void __launch_bounds__(kPartialThreads)
partial_kernel(const PartialArgs a)
{
  __shared__ __align__(16) double sEIGN[S], sEI[S * S], sEV[S * S], sTip[CODES * S];
  const int tid = threadIdx.x;
  constexpr int kEign = S == 4 ? 0 : kAaEign, kEI = S == 4 ? 8 : kAaEI, kEV = S == 4 ? 24 : kAaEV, kTip = S == 4 ? 40 : kAaTip;
  double *const sc = a.scratch + (size_t)blockIdx.x * a.nslots * S * kPartialThreads + tid;
  int loaded = -1;
  for (int ch = blockIdx.x; ch < a.nchunks; ch += gridDim.x) {
    const PartialChunk c = a.chunks[ch];
    if (c.part_cls != loaded) {                         // block-uniform
      __syncthreads();
      const double *m = a.models + (size_t)c.part_cls * a.model_doubles;
      for (int i = tid; i < S; i += kPartialThreads) sEIGN[i] = m[kEign + i];
      for (int i = tid; i < S * S; i += kPartialThreads) { sEI[i] = m[kEI + i]; sEV[i] = m[kEV + i]; }
      for (int i = tid; i < CODES * S; i += kPartialThreads) sTip[i] = m[kTip + i];
      loaded = c.part_cls;
      __syncthreads();
    }
    if (tid >= (int)c.count) continue;
    const PartialPair pr = a.pairs[c.begin + tid];
    const double ki = pr.rate;
    const uint8_t *tp = a.tips + pr.site;
    int scale = 0;
    for (int k = 1; k < a.count; k++) {
      const PartialEntry en = a.entries[k];
      const double lz1 = a.lzq[(size_t)k * a.nbranches + c.branch_slot] * ki;
      const double lz2 = a.lzr[(size_t)k * a.nbranches + c.branch_slot] * ki;
      double d1[S], d2[S], x3[S];
      if (en.tipCase == 2) {
        #pragma unroll
        for (int j = 0; j < S; j++) d1[j] = sc[(size_t)(en.q * S + j) * kPartialThreads];
      } else {
        const double *t = &sTip[S * tp[(size_t)(en.q - 1) * a.N]];
        #pragma unroll
        for (int j = 0; j < S; j++) d1[j] = t[j];
      }
      if (en.tipCase == 0) {
        const double *t = &sTip[S * tp[(size_t)(en.r - 1) * a.N]];
        #pragma unroll
        for (int j = 0; j < S; j++) d2[j] = t[j];
      } else {
        #pragma unroll
        for (int j = 0; j < S; j++) d2[j] = sc[(size_t)(en.r * S + j) * kPartialThreads];
      }
      #pragma unroll
      for (int j = 1; j < S; j++) {                     // d = x * exp(EIGN * (lz * ki)), :145-146
        d1[j] *= exp(sEIGN[j] * lz1);
        d2[j] *= exp(sEIGN[j] * lz2);
      }
      #pragma unroll
      for (int j = 0; j < S; j++) x3[j] = 0.0;
      #pragma unroll 4
      for (int j = 0; j < S; j++) {
        double u1 = 0.0, u2 = 0.0;
        #pragma unroll
        for (int l = 0; l < S; l++) { u1 += d1[l] * sEI[j * S + l]; u2 += d2[l] * sEI[j * S + l]; }
        const double x12 = u1 * u2;
        #pragma unroll
        for (int l = 0; l < S; l++) x3[l] += x12 * sEV[j * S + l];
      }
      bool small = true;
      #pragma unroll
      for (int j = 0; j < S; j++) small = small && (x3[j] < kMinLikelihood) && (x3[j] > -kMinLikelihood);
      if (small) {                                      // :171-180
        #pragma unroll
        for (int j = 0; j < S; j++) x3[j] *= kTwoToThe256;
        scale++;
      }
      #pragma unroll
      for (int j = 0; j < S; j++) sc[(size_t)(en.p * S + j) * kPartialThreads] = x3[j];
    }
    // root branch: tip vector of entries[0].pNumber against the slot of entries[0].qNumber, :228-249
    const PartialEntry root = a.entries[0];
    const double *x1 = &sTip[S * tp[(size_t)(root.p - 1) * a.N]];
    const double lz = a.root_lz[c.branch_slot] * ki;
    double term = x1[0] * sc[(size_t)(root.q * S) * kPartialThreads];
    #pragma unroll
    for (int j = 1; j < S; j++) term += x1[j] * sc[(size_t)(root.q * S + j) * kPartialThreads] * exp(sEIGN[j] * lz);
    a.out[c.begin + tid] = (log(fabs(term)) + scale * a.log_minlikelihood) * (double)a.wgt[pr.site];
  }
}

void launch_partial(const PartialArgs &a, int states, int blocks, cudaStream_t s)
{
  if (states == 4) partial_kernel<4, 16><<<blocks, kPartialThreads, 0, s>>>(a);
  else             partial_kernel<20, 23><<<blocks, kPartialThreads, 0, s>>>(a);
}

int partial_max_blocks(int states, int device)
{
  int per_sm = 0, sms = 0;
  if (states == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, partial_kernel<4, 16>, kPartialThreads, 0);
  else             cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, partial_kernel<20, 23>, kPartialThreads, 0);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (per_sm < 1) per_sm = 1;
  if (sms < 1) sms = 148;
  return per_sm * sms;
}

}  // namespace exb
