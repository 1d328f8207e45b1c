// Measured throughput of the LEGACY tensor path on one B200: mma.sync.m16n8k16 f16 x f16 -> f32 (SASS HMMA.16816.F32) and
// m16n8k8 tf32, as a function of resident warps per SM -- the ceiling of the register-resident small-bond family (smpo_rr.cuh).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/hmma_peak tools/ubench/hmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC> __global__ void hmma_loop(float *out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one dependent chain per warp: latency of an accumulating HMMA
__global__ void hmma_chain(float *out, int iters, long long *cycles) {
  float c[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
  if (c[0] == 123.456f) out[threadIdx.x] = c[0];
}

template <typename F> static float best_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  float *out;
  long long *cyc;
  cudaMalloc(&out, 148 * 64 * 1024 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  const int iters = 20000;
  printf("{\"hmma_16816_f16\": {");
  const int warps_per_sm[] = {4, 8, 16, 32, 64};
  for (int k = 0; k < 5; ++k) {
    const int w = warps_per_sm[k];
    const int threads = w >= 8 ? 256 : w * 32, ctas = 148 * (w >= 8 ? w / 8 : 1);
    float ms = best_ms([&] { hmma_loop<8><<<ctas, threads>>>(out, iters); }, 5);
    double fl = 2.0 * 16 * 8 * 16 * 8 * (double)iters * ((double)ctas * threads / 32);
    printf("\"%d_warps_per_sm\": %.1f%s", w, fl / (ms * 1e-3) / 1e12, k < 4 ? ", " : "");
  }
  printf("}, \"unit\": \"TFLOP/s dense, 8 independent accumulators per warp\", ");
  hmma_chain<<<1, 32>>>(out, 10000, cyc);
  long long h = 0;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("\"dependent_hmma_latency_cycles\": %.1f, \"cuda_error\": \"%s\"}\n", (double)h / 10000.0, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
