// Microbenchmark: tcgen05.ld throughput per SM for several shapes and warp counts (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(SHAPE)                                                                                                  \
  asm volatile("tcgen05.ld.sync.aligned." SHAPE ".b32 "                                                               \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                              \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"              \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),      \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),           \
                 "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),           \
                 "=r"(r[30]), "=r"(r[31])                                                                            \
               : "r"(taddr)                                                                                          \
               : "memory")

// MODE 0: 32x32b.x32 (32 lanes x 32 columns)   1: 16x256b.x8 (16 lanes x 64 columns)   2: 16x128b.x16 (16 lanes x 64 columns)
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(int iters, int depth, unsigned long long *cycles, unsigned *sink) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    uint32_t r[32];
    for (int d = 0; d < depth; d++) {
      const uint32_t col = (uint32_t)(((i * depth + d) * 64 + (warp >> 2) * 32) & 511 & ~63);
      const uint32_t taddr = base + col + (MODE == 0 ? 0u : ((uint32_t)((i & 1) * 16) << 16));
      if (MODE == 0) LD32("32x32b.x32");
      if (MODE == 1) LD32("16x256b.x8");
      if (MODE == 2) LD32("16x128b.x16");
      acc += r[0] ^ r[31];
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
  unsigned long long *cyc;
  unsigned *sink;
  cudaMalloc(&cyc, 148 * 8);
  cudaMalloc(&sink, 4);
  const int iters = 20000;
  for (int mode = 0; mode < 3; mode++)
    for (int warps : {4, 8, 16, 32})
      for (int depth : {1, 2, 4}) {
        for (int rep = 0; rep < 2; rep++) {
          if (mode == 0) k<0><<<148, warps * 32>>>(iters, depth, cyc, sink);
          if (mode == 1) k<1><<<148, warps * 32>>>(iters, depth, cyc, sink);
          if (mode == 2) k<2><<<148, warps * 32>>>(iters, depth, cyc, sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("mode %d warps %d depth %d: %s\n", mode, warps, depth, cudaGetErrorString(e));
            return 1;
          }
        }
        unsigned long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        unsigned long long mx = 0;
        for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
        const double bytes = (double)warps * iters * depth * 4096.0;   // every load moves 32 registers x 32 lanes x 4 B
        printf("{\"shape\": \"%s\", \"warps\": %d, \"loads_in_flight\": %d, \"bytes_per_clk_per_sm\": %.1f}\n",
               mode == 0 ? "32x32b.x32" : mode == 1 ? "16x256b.x8" : "16x128b.x16", warps, depth, bytes / (double)mx);
      }
  return 0;
}
