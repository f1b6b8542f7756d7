// Micro-probe: does a DFMA (16 lanes/clk/SMSP -> 2 cycles per warp instruction) leave its second
// cycle free for an integer / LDS instruction, or does it block the issue port for both?
// Every instruction is pinned with asm volatile.  Per loop iteration and thread: 16 independent DFMA
// chains, plus per DFMA:  mode 1: 1 LOP3   mode 2: 2 LOP3   mode 3: 1 IMAD   mode 4: 1 LDS.64 per 4 DFMA
//                         mode 5: 1 IMAD + 1 LOP3 + LDS.64 per 2 DFMA (roughly the exp2s mix)
// 4 warps per SMSP.
#include <cuda_runtime.h>
#include <stdio.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(int iters, double* out, int seed) {
  __shared__ double tbl[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) tbl[i] = 1.0 + i * 1e-12;
  __syncthreads();
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(tbl) + 8 * (threadIdx.x & 15);
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[16], tv[4] = {0, 0, 0, 0};
  unsigned q[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i] = i; q[i] = seed + i * 7 + threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
      if (MODE == 1 || MODE == 2 || MODE == 5) asm volatile("and.b32 %0, %0, 0x7ffffff;" : "+r"(q[i]));
      if (MODE == 2) asm volatile("xor.b32 %0, %0, 0x10101;" : "+r"(q[i]));
      if (MODE == 3 || MODE == 5) asm volatile("mad.lo.u32 %0, %0, 129, %1;" : "+r"(q[i]) : "r"(seed));
      if (MODE == 4 && (i & 3) == 0)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(tv[i >> 2]) : "r"(sbase + ((q[i] + it) & 63) * 128));
      if (MODE == 5 && (i & 1) == 0)
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(tv[(i >> 1) & 3]) : "r"(sbase + (q[i] & 63) * 128));
    }
  }
  double s = tv[0] + tv[1] + tv[2] + tv[3];
  unsigned t = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) { s += c[i]; t += q[i]; }
  if (s == 12345.678 || t == 0xdeadbeefu) out[0] = s + t;
}

template <int MODE>
void run(double* d) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<148, 512>>>(100, d, 3);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<MODE><<<148, 512>>>(iters, d, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = 16.0 * 16 * 32 * iters * 148;   // warps * chains * lanes
  printf("{\"mode\": %d, \"ms\": %.3f, \"dfma_tflops\": %.2f, \"cycles_per_dfma_per_smsp\": %.3f}\n", MODE, ms,
         2 * dfma / ms * 1e-9, ms * 1e-3 * 1.965e9 / (4.0 * 16 * iters));
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  run<0>(d); run<1>(d); run<2>(d); run<3>(d); run<4>(d); run<5>(d);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
