// micro-benchmark (GPU box): how fast does one SM ingest SMALL cp.async.bulk copies from L2?
//   persistent CTA per SM; W issuing warps; every lane issues `per_lane` copies of `bytes` each per round into a shared-memory
//   ring slot of its own, all completing on one mbarrier per round; source offsets pseudo-random inside a `span_mb` buffer
//   (L2-resident when small). Prints bytes / cycle / SM and GB/s. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_copy_rate bulk_copy_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(1024, 1) k(const char *src, size_t span, int bytes, int warps, int per_lane, int rounds, unsigned long long *cycles) {
  extern __shared__ __align__(128) char ring[];
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(&bar)), "r"(1u)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const int ncopy = warps * 32 * per_lane;
  unsigned long long t0 = clock64();
  uint32_t seed = blockIdx.x * 9781u + threadIdx.x * 6271u + 17u;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(&bar)), "r"((uint32_t)(ncopy * bytes)) : "memory");
    __syncthreads();
    if (warp < warps) {
      for (int q = 0; q < per_lane; ++q) {
        seed = seed * 1664525u + 1013904223u;
        const size_t off = ((size_t)(seed >> 8) * (size_t)bytes) % (span - bytes);
        const int slot = (warp * 32 + lane) * per_lane + q;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(saddr(ring + (size_t)slot * bytes)), "l"(src + (off & ~(size_t)15)), "r"((uint32_t)bytes), "r"(saddr(&bar)) : "memory");
      }
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(saddr(&bar)), "r"((uint32_t)(r & 1)) : "memory");
    __syncthreads();
  }
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main(int argc, char **argv) {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t span_mb = argc > 1 ? atoi(argv[1]) : 16;
  const size_t span = span_mb << 20;
  char *src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
  unsigned long long *cyc; cudaMallocManaged(&cyc, sizeof(unsigned long long) * sms);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("span %zu MB, %d SMs\n", span_mb, sms);
  const int cfgs[][3] = {{2048, 1, 1}, {2048, 1, 3}, {2048, 2, 1}, {2048, 3, 1}, {2048, 1, 2}, {1024, 1, 4}, {1024, 4, 1}, {4096, 1, 1}, {512, 4, 2}, {16384, 1, 1}};   // bytes, warps, per_lane (warps*32*per_lane*bytes <= 200 KB)
  for (auto &c : cfgs) {
    const int bytes = c[0], warps = c[1], per_lane = c[2], rounds = 200;
    const size_t smem = (size_t)warps * 32 * per_lane * bytes;
    if (smem > 200 * 1024) continue;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<sms, 1024, smem>>>(src, span, bytes, warps, per_lane, 5, cyc); cudaDeviceSynchronize();
    cudaEventRecord(a); k<<<sms, 1024, smem>>>(src, span, bytes, warps, per_lane, rounds, cyc); cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long mx = 0; for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double total = (double)sms * rounds * smem;
    printf("copy %5d B  issuing warps %d  copies/lane %d  (%6.1f KB per round per SM): %7.1f B/cycle/SM  %8.1f GB/s  round %6.0f cycles  err %s\n", bytes, warps, per_lane,
           smem / 1024.0, (double)rounds * smem / (double)mx, total / (ms * 1e-3) / 1e9, (double)mx / rounds, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
