// umma_tf32.cu -- self-test of the tcgen05 building block the mixed-precision AA path needs:
//   D[128 x 32] (FP32, TMEM) = A[128 x 24] (TF32, written to TMEM by the threads, one row per thread)
//                            x B[24 x 32]^T-major (TF32, shared memory, K-major canonical layout, no swizzle)
// as three tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=32, K=8) issued by one thread, completion through
// tcgen05.commit on an mbarrier, result read back with tcgen05.ld.  Compared with a host FP64 product of the same
// TF32-rounded operands.  Also times a long dependent chain of these MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32 umma_tf32.cu ;  run: ./umma_tf32
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int M = 128, N = 32, K = 24;
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
constexpr int kCols = 128;           // TMEM columns allocated: A at [0, 24), D at [32, 64)

// B element (n, k) in the canonical K-major no-swizzle layout: core matrix = 8 rows (n) x 16 bytes (4 k)
__host__ __device__ inline int b_offset_words(int n, int k) { return ((k >> 2) * (N / 8) + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                       // version = 1 (Blackwell)
  return d;                                     // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__global__ void __launch_bounds__(128, 1) umma_test(const uint32_t *A, const uint32_t *B, float *D, int reps, long long *cycles)
{
  __shared__ __align__(128) uint32_t sB[K * N];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < K * N; i += 128) { const int n = i / K, k = i % K; sB[b_offset_words(n, k)] = B[n * K + k]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);  // this warp's 32 lanes; thread = lane = row
  // A row of this thread -> TMEM columns [0, 24)
  uint32_t a[24];
  #pragma unroll
  for (int k = 0; k < 24; k++) a[k] = A[tid * K + k];
  #pragma unroll
  for (int j = 0; j < 3; j++)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(lane_addr + 8 * j), "r"(a[8 * j]), "r"(a[8 * j + 1]), "r"(a[8 * j + 2]), "r"(a[8 * j + 3]),
                    "r"(a[8 * j + 4]), "r"(a[8 * j + 5]), "r"(a[8 * j + 6]), "r"(a[8 * j + 7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long t0 = clock64();
  if (tid == 0) {
    for (int r = 0; r < reps; r++) {
      #pragma unroll
      for (int j = 0; j < 3; j++) {
        const uint64_t bdesc = make_desc(smem_u32(sB) + (uint32_t)(2 * j) * (N / 8) * 128u, (N / 8) * 128u, 128u);
        const uint32_t acc = (j > 0 || r > 0) ? 1u : 0u;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     :: "r"(tbase + 32u), "r"(tbase + 8u * j), "l"(bdesc), "r"(kIdesc), "r"(acc), "r"(0u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the MMAs
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE;\n\tbra W;\n\tDONE:\n\t}"
               :: "r"(smem_u32(&bar)) : "memory");
  const long long t1 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t d[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                 "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]),
                 "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]),
                 "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
               : "r"(lane_addr + 32u) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  #pragma unroll
  for (int n = 0; n < 32; n++) D[tid * N + n] = __uint_as_float(d[n]);
  if (tid == 0) *cycles = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(kCols) : "memory");
}

static uint32_t tf32_bits(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu; return u; }
static float bits_f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main()
{
  std::vector<uint32_t> A(M * K), B(N * K);
  srand(7);
  for (auto &v : A) v = tf32_bits((float)rand() / RAND_MAX - 0.5f);
  for (auto &v : B) v = tf32_bits((float)rand() / RAND_MAX - 0.5f);
  for (int m = 0; m < M; m++) for (int k = 20; k < K; k++) A[m * K + k] = 0;      // K padding 20 -> 24
  uint32_t *dA, *dB; float *dD; long long *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  for (int reps : {1, 1000}) {
    umma_test<<<1, 128>>>(dA, dB, dD, reps, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> D(M * N); long long cyc;
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) {
      double ref = 0;
      for (int k = 0; k < K; k++) ref += (double)bits_f(A[m * K + k]) * (double)bits_f(B[n * K + k]);
      ref *= reps;
      maxerr = fmax(maxerr, fabs(D[m * N + n] - ref)); maxref = fmax(maxref, fabs(ref));
    }
    printf("reps %d: max |D - ref| = %.3e (max |ref| %.3e) %s; %lld cycles = %.1f per MMA (M128 N32 K8)\n", reps, maxerr, maxref,
           maxerr <= 2e-6 * maxref * (reps > 1 ? 50 : 1) ? "OK" : "MISMATCH", cyc, (double)cyc / (3.0 * reps));
  }
  return 0;
}
