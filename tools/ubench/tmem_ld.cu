// Micro-benchmark: tcgen05.ld throughput / latency from TMEM on sm_100a (measurement aid for the chain-kernel epilogue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[X]);
template <> __device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
template <> __device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// nwarps warps (warp w reads lane quarter w % 4); each issues `inflight` loads of X columns, waits, repeats `iters` times
template <int X>
__global__ void bench(int iters, int inflight, long long *out, uint32_t *sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[4][X];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < inflight) ld<X>(base + ((uint32_t)((it * 4 + k) * X) & 511u & ~(uint32_t)(X - 1)), r[k]);
    ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < inflight) {
#pragma unroll
        for (int j = 0; j < X; ++j) acc ^= r[k][j];
      }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(slot), "r"(512u) : "memory");
}

int main() {
  long long *out;
  uint32_t *sink;
  cudaMalloc(&out, 1024);
  cudaMalloc(&sink, 64);
  const int iters = 2000;
  printf("shape warps inflight cycles/iter bytes/clk(SM)\n");
  for (int X : {16, 32})
    for (int nw : {1, 4, 8, 16})
      for (int inf : {1, 2, 4}) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          if (X == 16) bench<16><<<1, nw * 32>>>(iters, inf, out, sink);
          else bench<32><<<1, nw * 32>>>(iters, inf, out, sink);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        }
        const double cyc = (double)h / iters;
        const double bytes = (double)nw * inf * X * 32 * 4;
        printf("x%-3d %5d %8d %10.1f %10.1f\n", X, nw, inf, cyc, bytes / cyc);
      }
  return 0;
}
