// write_floor.cu -- how fast can B200 absorb the CLV write stream of the DNA traverse kernel?
//   pattern A: the kernel's own pattern -- CTA t (of 3907) writes, for node e = 0..197, 256 sites x 128 B = 32 KB
//              at node e's array (stride 128 MB) + t * 32 KB, one 256-bit store per thread and site quartet
//   pattern B: one linear pass over the same 25.3 GB
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_floor write_floor.cu ;  run: ./write_floor
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void st256(double *p, double a, double b, double c, double d)
{
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__global__ void __launch_bounds__(256, 2) pattern_a(double *clv, size_t N, int nodes, int ns)
{
  const size_t site0 = (size_t)blockIdx.x * 64 * ns;
  for (int e = 0; e < nodes; e++) {
    double *dst = clv + (size_t)e * N * 16;
    for (int s = 0; s < ns; s++) {
      const size_t site = site0 + s * 64 + (threadIdx.x >> 2);
      if (site < N) st256(dst + site * 16 + (threadIdx.x & 3) * 4, e, s, 1.0, 2.0);
    }
  }
}
__global__ void __launch_bounds__(256, 2) pattern_b(double *clv, size_t total4)
{
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x)
    st256(clv + i * 4, 1.0, 2.0, 3.0, 4.0);
}
int main()
{
  const size_t N = 1000000; const int nodes = 198;
  const size_t bytes = (size_t)nodes * N * 128;
  double *clv; if (cudaMalloc(&clv, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int ns : {4, 2, 1}) {
    const int grid = (int)((N + 64 * ns - 1) / (64 * ns));
    float best = 1e9f;
    for (int it = 0; it < 6; it++) {
      cudaEventRecord(e0); pattern_a<<<grid, 256>>>(clv, N, nodes, ns); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("pattern A ns=%d grid=%d: %.3f ms  %.1f GB/s\n", ns, grid, best, bytes / best / 1e6);
  }
  for (int grid : {296, 592, 1184, 148 * 16}) {
    float best = 1e9f;
    for (int it = 0; it < 6; it++) {
      cudaEventRecord(e0); pattern_b<<<grid, 256>>>(clv, bytes / 32); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("pattern B grid=%d: %.3f ms  %.1f GB/s\n", grid, best, bytes / best / 1e6);
  }
  float best = 1e9f;
  for (int it = 0; it < 4; it++) {
    cudaEventRecord(e0); cudaMemsetAsync(clv, 0, bytes); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
  }
  printf("cudaMemset: %.3f ms  %.1f GB/s\n", best, bytes / best / 1e6);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
