// Micro-probe: do DFMA (FP64 pipe) and DMMA.8x8x4 (tensor pipe) execute concurrently on sm_100a?
//   mode 0: 16 warps/SM of DFMA chains      -> T0
//   mode 1: 16 warps/SM of DMMA chains      -> T1
//   mode 2: 16 + 16 warps/SM, both          -> T2  (independent pipes: ~max(T0,T1); shared: ~T0+T1)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(1024) probe(int mode, int iters, double* out) {
  const int warp = threadIdx.x >> 5;
  const bool do_mma = mode == 1 || (mode == 2 && warp >= 16);
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i;
  if (do_mma) {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 8; i++) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  } else {
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
#pragma unroll
      for (int i = 0; i < 16; i++) c[i] = fma(c[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; mode++) {
    const int threads = mode == 2 ? 1024 : 512;
    probe<<<148, threads>>>(mode, 100, d);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<<<148, threads>>>(mode, iters, d);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM: DFMA warps do 32 DFMA x 32 lanes per iter; DMMA warps do 8 x 256 FMA per iter
    const double fma_dfma = (mode == 1) ? 0 : 16.0 * 32 * 32 * iters * 148;
    const double fma_dmma = (mode == 0) ? 0 : 16.0 * 8 * 256 * iters * 148;
    printf("{\"mode\": %d, \"ms\": %.3f, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f}\n", mode, ms,
           2 * fma_dfma / ms * 1e-9, 2 * fma_dmma / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
