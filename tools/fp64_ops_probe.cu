// Micro-probe: per-instruction throughput of the FP64 pipe on sm_100a for the operand forms the
// covariance kernels use.  16 independent chains per thread, 4 warps per SMSP.
//  0 DFMA r,r,r   1 DADD r,r   2 DMUL r,r   3 DADD r,imm64   4 DFMA r,r,imm   5 DFMA r,UR,r (kernel-arg operand)
//  6 mix: DADD+DADD+DADD+3xDFMA+DMUL+DFMA (the exp2s sequence, dependent within a chain)
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(int iters, double* out, double ua, double ub) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i + 1e-3 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
      if (MODE == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(c[i]) : "d"(a));
      if (MODE == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(c[i]) : "d"(a));
      if (MODE == 3) asm volatile("add.rn.f64 %0, %0, 0d4338000000000000;" : "+d"(c[i]));
      if (MODE == 4) asm volatile("fma.rn.f64 %0, %0, %1, 0d3F662E42FEFA39EF;" : "+d"(c[i]) : "d"(a));
      if (MODE == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(ua), "d"(ub));
      if (MODE == 6) {
        double t, f, h;
        asm volatile("add.rn.f64 %0, %1, 0d4338000000000000;" : "=d"(t) : "d"(c[i]));
        asm volatile("add.rn.f64 %0, %0, 0dC338000000000000;" : "+d"(t));
        asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(f) : "d"(c[i]), "d"(t));
        asm volatile("fma.rn.f64 %0, %1, %2, 0d3E2C6B08D704A0C0;" : "=d"(h) : "d"(f), "d"(ua));
        asm volatile("fma.rn.f64 %0, %0, %1, 0d3ECEBFBDFF82C58F;" : "+d"(h) : "d"(f));
        asm volatile("fma.rn.f64 %0, %0, %1, 0d3F662E42FEFA39EF;" : "+d"(h) : "d"(f));
        asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(h) : "d"(f));
        asm volatile("fma.rn.f64 %0, %1, %2, %1;" : "=d"(c[i]) : "d"(a), "d"(h));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
void run(double* d) {
  const int iters = MODE == 6 ? 4000 : 20000;
  const int per = MODE == 6 ? 8 : 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<148, 512>>>(100, d, 1.0000001, 0.9999999);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<148, 512>>>(iters, d, 1.0000001, 0.9999999);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"mode\": %d, \"ms\": %.3f, \"nominal_cycles_per_fp64_instr_per_smsp\": %.3f}\n", MODE, ms,
         ms * 1e-3 * 1.965e9 / (4.0 * 16 * iters * per));
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  run<0>(d); run<1>(d); run<2>(d); run<3>(d); run<4>(d); run<5>(d); run<6>(d);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
