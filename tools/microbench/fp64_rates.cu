// Microbenchmark: DFMA vs DMMA.8x8x4 issue rate on sm_100a (decides the AA contraction strategy).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rates fp64_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dmma_kernel(double *out, int iters, double seed)
{
  double c[CHAINS][2];
  double a = seed + threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
  #pragma unroll
  for (int i = 0; i < CHAINS; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < CHAINS; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void dfma_kernel(double *out, int iters, double seed)
{
  double c[CHAINS];
  double a = seed + threadIdx.x * 1e-9, b = 1e-9;
  #pragma unroll
  for (int i = 0; i < CHAINS; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
    #pragma unroll
    for (int i = 0; i < CHAINS; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
  #pragma unroll
  for (int i = 0; i < CHAINS; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double *out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  const int iters = 20000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32, blocks = sms * (warps * 32 / threads);
    float ms = time_it([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0); });
    double n = (double)blocks * (threads / 32) * iters * 8;
    printf("DMMA 8 chains, %2d warps/SM: %.3f ms  %.3f DMMA/clk/SM (@1.9GHz)  %.1f TFLOP/s\n", warps, ms,
           n / sms / (ms * 1e-3 * 1.9e9), n * 512 / (ms * 1e-3) / 1e12);
    ms = time_it([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0); });
    n = (double)blocks * (threads / 32) * iters * 8;
    printf("DFMA 8 chains, %2d warps/SM: %.3f ms  %.3f warp-DFMA/clk/SM  %.1f TFLOP/s\n", warps, ms,
           n / sms / (ms * 1e-3 * 1.9e9), n * 64 / (ms * 1e-3) / 1e12);
  }
  for (int ch = 1; ch <= 4; ch *= 2) {
    float ms;
    if (ch == 1) ms = time_it([&] { dmma_kernel<1><<<sms, 128>>>(out, iters, 1.0); });
    else if (ch == 2) ms = time_it([&] { dmma_kernel<2><<<sms, 128>>>(out, iters, 1.0); });
    else ms = time_it([&] { dmma_kernel<4><<<sms, 128>>>(out, iters, 1.0); });
    printf("DMMA latency probe: %d chain(s), 1 warp/SMSP: %.1f clk per dependent DMMA step (@1.9GHz)\n", ch,
           ms * 1e-3 * 1.9e9 / iters);
  }
  return 0;
}
