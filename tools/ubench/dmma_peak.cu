// Measured FP64 tensor-pipe peak (mma.sync.m8n8k4.f64, SASS DMMA) of one B200: the denominator of the float64 roofline
// fractions (SURVEY 8d had only the nominal 40 TFLOP/s).  Prints one JSON line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/dmma_peak tools/ubench/dmma_peak.cu
// Also reports the plain DFMA rate (FP64 FMA pipe) for context.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC> __global__ void __launch_bounds__(256) dmma_loop(double *out, int iters) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_loop(double *out, int iters) {
  double c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> static float time_ms(F launch, int reps) {
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
  double *out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  const int iters = 20000, grid = 148 * 4;
  // burst: best of 10 launches
  float ms8 = time_ms([&] { dmma_loop<8><<<grid, 256>>>(out, iters); }, 10);
  float ms16 = time_ms([&] { dmma_loop<16><<<grid, 256>>>(out, iters); }, 10);
  const double flops8 = 2.0 * 8 * 8 * 4 * 8 * (double)iters * (grid * 8);    // per MMA 512 flops, NACC per iter, warps
  const double flops16 = 2.0 * 8 * 8 * 4 * 16 * (double)iters * (grid * 8);
  double tf8 = flops8 / (ms8 * 1e-3) / 1e12, tf16 = flops16 / (ms16 * 1e-3) / 1e12;
  // sustained: back to back for ~3 s
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int n = 0;
  cudaEventRecord(e0);
  const int batch = (int)(3000.0f / ms16) + 1;
  for (; n < batch; ++n) dmma_loop<16><<<grid, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms_s;
  cudaEventElapsedTime(&ms_s, e0, e1);
  double tf_s = flops16 * n / (ms_s * 1e-3) / 1e12;
  float msf = time_ms([&] { dfma_loop<<<grid * 2, 256>>>(out, iters); }, 10);
  double tff = 2.0 * 8 * (double)iters * (grid * 2 * 256) / (msf * 1e-3) / 1e12;
  cudaError_t err = cudaGetLastError();
  printf("{\"fp64_dmma_tflops\": %.2f, \"fp64_dmma_tflops_sustained\": %.2f, \"fp64_dmma_tflops_8acc\": %.2f, "
         "\"fp64_dfma_tflops\": %.2f, \"how\": \"mma.sync.m8n8k4.f64, 16 independent accumulators per warp, 8 warps x 4 CTAs per SM, "
         "best of 10 (burst) and back to back for 3 s (sustained)\", \"cuda_error\": \"%s\"}\n",
         tf16, tf_s, tf8, tff, cudaGetErrorString(err));
  return err == cudaSuccess ? 0 : 1;
}
