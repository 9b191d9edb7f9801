// core_tile.cuh -- per-tile body of the makenewz "core" step (first and second derivative of lnL with respect to
// lz = log z over the sum table), shared by the one-shot core kernels (one launch per Newton-Raphson iteration,
// multi-rank path) and the device-resident Newton-Raphson kernel (kernels_nr.cu).
//   coreGTRGAMMA      makenewzGenericSpecial.c:2309-2393        coreGTRGAMMAPROT  makenewzGenericSpecial.c:2581-2655
#pragma once
#include "engine.cuh"

namespace exb {

struct CoreSmem { double d0[80], d1[80], d2[80], red[32]; };

// model-block accessors (layouts in engine.cuh): AA keeps one eigen-system per category
template <int S> __device__ __forceinline__ double m_eign(const double *m, int c, int k) { return S == 20 ? m[kAaEign + c * 20 + k] : m[k]; }
template <int S> __device__ __forceinline__ double m_rate(const double *m, int c) { return S == 20 ? m[kAaRates + c] : m[S + c]; }
template <int S> __device__ __forceinline__ double m_weight(const double *m, int c) { return S == 20 ? m[kAaWeights + c] : 1.0; }

__device__ __forceinline__ void core_ld256(const double *p, double *v)
{
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}

// All threads of a 256-thread CTA; result valid in thread 0.
template <int S>
__device__ __forceinline__ void core_tile(const Tile &tile, const double *sum, const int *wgt, const double *model,
                                          double lz, CoreSmem &sh, double &out1, double &out2)
{
  constexpr int SPAN = 4 * S;
  const int tid = threadIdx.x, c = tid & 3;
  __syncthreads();                                 // the tables may still be read by the previous tile
  if (tid < SPAN) {
    const int cc = tid / S, k = tid % S;
    const double ki = m_rate<S>(model, cc), ek = m_eign<S>(model, cc, k);
    sh.d0[tid] = k ? exp(ek * ki * lz) : 1.0;
    sh.d1[tid] = k ? ek * ki : 0.0;
    sh.d2[tid] = k ? ek * ek * (ki * ki) : 0.0;
  }
  __syncthreads();
  double acc1 = 0.0, acc2 = 0.0;
  const double wc = m_weight<S>(model, c);
  const int npass = (tile.nsites + 63) >> 6;
  for (int u = 0; u < npass; u++) {
    const int i = (u << 6) + (tid >> 2);
    const bool valid = i < (int)tile.nsites;
    const size_t site = (size_t)tile.site0 + (valid ? i : 0);
    const double *p = sum + site * SPAN + c * S;
    const double w = (double)wgt[site];            // early: its latency hides under the contraction
    double L = 0.0, L1 = 0.0, L2 = 0.0;
    #pragma unroll
    for (int j = 0; j < S; j += 4) {
      double s[4];
      core_ld256(p + j, s);
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        const double t = sh.d0[c * S + j + k] * s[k];
        L += t; L1 += t * sh.d1[c * S + j + k]; L2 += t * sh.d2[c * S + j + k];
      }
    }
    if (S == 20) { const double w = wc; L *= w; L1 *= w; L2 *= w; }     // LG4X category weights (1.0 otherwise)
    L  += __shfl_xor_sync(0xffffffffu, L, 1);   L  += __shfl_xor_sync(0xffffffffu, L, 2);
    L1 += __shfl_xor_sync(0xffffffffu, L1, 1);  L1 += __shfl_xor_sync(0xffffffffu, L1, 2);
    L2 += __shfl_xor_sync(0xffffffffu, L2, 1);  L2 += __shfl_xor_sync(0xffffffffu, L2, 2);
    if (c == 0 && valid) {
      const double inv = 1.0 / fabs(L);
      const double g = L1 * inv, h = L2 * inv;
      acc1 += w * g;
      acc2 += w * (h - g * g);
    }
  }
  out1 = block_sum(acc1, sh.red);
  out2 = block_sum(acc2, sh.red);
}

// CAT / PSR: one rate category per site (coreCAT_FLEX makenewzGenericSpecial.c:394-519).  Thread = site.
//   d[cat][k] = exp((EIGN[k] lz) r_cat);  L = sum_k d s,  L' = sum_k d s EIGN[k],  L'' = sum_k d s EIGN[k]^2
//   dlnL += (r w) L'/|L|,   d2lnL += (r^2 w) (L''/|L| - (L'/|L|)^2)
struct CoreCatSmem { double d[64 * 20]; double red[32]; };

template <int S>
__device__ __forceinline__ void core_tile_cat(const Tile &tile, const double *sum, const int *wgt, const double *model,
                                              const int *rate_cat, const double *rates, int max_cats,
                                              double lz, CoreCatSmem &sh, double &out1, double &out2)
{
  const int tid = threadIdx.x;
  __syncthreads();
  for (int i = tid; i < max_cats * S; i += blockDim.x) {
    const int n = i / S, k = i % S;
    sh.d[i] = k ? exp((m_eign<S>(model, 0, k) * lz) * rates[n]) : 1.0;
  }
  __syncthreads();
  double acc1 = 0.0, acc2 = 0.0;
  for (unsigned i = tid; i < tile.nsites; i += blockDim.x) {
    const size_t site = (size_t)tile.site0 + i;
    const int cat = rate_cat[site];
    const double *p = sum + site * S;
    const double *d = sh.d + cat * S;
    double L = 0.0, L1 = 0.0, L2 = 0.0;
    #pragma unroll
    for (int j = 0; j < S; j += 4) {
      double s[4];
      core_ld256(p + j, s);
      #pragma unroll
      for (int k = 0; k < 4; k++) {
        if (j + k == 0) { L = s[0]; continue; }
        const double ek = m_eign<S>(model, 0, j + k), t = d[j + k] * s[k];
        L += t; L1 += t * ek; L2 += t * (ek * ek);
      }
    }
    const double inv = 1.0 / fabs(L), r = rates[cat], w = (double)wgt[site];
    const double g = L1 * inv, h = L2 * inv;
    acc1 += (r * w) * g;
    acc2 += (r * r * w) * (h - g * g);
  }
  out1 = block_sum(acc1, sh.red);
  out2 = block_sum(acc2, sh.red);
}

}  // namespace exb
