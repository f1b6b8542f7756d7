// Micro-probe: sustained rate of fastmath.cuh's exp2s in isolation (no global traffic): NCH independent
// evaluations per thread and iteration, 4 or 8 warps per SMSP.  Variants strip parts of the function to
// see what keeps it from the FP64-pipe bound of 8 instructions x 2 cycles = 16 cycles per warp-level exp.
//   V 0 full exp2s   1 no clamp   2 no clamp, no table (T = 1)   3 no clamp, no scale   4 FP64 part only
#include <cuda_runtime.h>
#include <stdio.h>
#include "../multifidelity_datafusion_gps_b200/csrc/fastmath.cuh"

template <int V>
__device__ __forceinline__ double exp_var(double u, unsigned tbl) {
  const double MAGIC = 6755399441055744.0;
  if (V == 0) {
    const unsigned hi = min((unsigned)__double2hiint(u), 0xC10FF000u);
    u = __hiloint2double((int)hi, __double2loint(u));
  }
  const double t = u + MAGIC;
  const int n = __double2loint(t);
  const double f = u - (t - MAGIC);
  double h = fma(f, 2.239395190875157e-12, 3.3083026805413713e-09);
  h = fma(h, f, 3.6655655969101062e-06);
  h = fma(h, f, 2.7076061740622863e-03);
  double T = 1.0;
  if (V == 0 || V == 1 || V == 3) {
    asm("{\n\t.reg .u32 j, a;\n\tand.b32 j, %1, 255;\n\tmad.lo.u32 a, j, 128, %2;\n\tld.shared.f64 %0, [a];\n\t}"
        : "=d"(T) : "r"(n), "r"(tbl));
  }
  const double res = fma(T, h * f, T);
  if (V == 3 || V == 4) return res + (V == 4 ? 0.0 : 0.0);
  int hi;
  asm("{\n\t.reg .u32 q;\n\tand.b32 q, %1, 0xffffff00;\n\tmad.lo.u32 %0, q, 4096, %2;\n\t}"
      : "=r"(hi) : "r"(n), "r"(__double2hiint(res)));
  return __hiloint2double(hi, __double2loint(res));
}

// V 5: rounding on the XU pipe (FRND.F64 + F2I.S32.F64 saturating) instead of the magic-number DADDs
template <>
__device__ __forceinline__ double exp_var<5>(double u, unsigned tbl) {
  double r;
  int n;
  asm("cvt.rni.f64.f64 %0, %1;" : "=d"(r) : "d"(u));
  asm("cvt.rni.s32.f64 %0, %1;" : "=r"(n) : "d"(u));
  n = max(n, -261632);
  const double f = u - r;
  double h = fma(f, 2.239395190875157e-12, 3.3083026805413713e-09);
  h = fma(h, f, 3.6655655969101062e-06);
  h = fma(h, f, 2.7076061740622863e-03);
  double T;
  asm("{\n\t.reg .u32 j, a;\n\tand.b32 j, %1, 255;\n\tmad.lo.u32 a, j, 128, %2;\n\tld.shared.f64 %0, [a];\n\t}"
      : "=d"(T) : "r"(n), "r"(tbl));
  const double res = fma(T, h * f, T);
  int hi;
  asm("{\n\t.reg .u32 q;\n\tand.b32 q, %1, 0xffffff00;\n\tmad.lo.u32 %0, q, 4096, %2;\n\t}"
      : "=r"(hi) : "r"(n), "r"(__double2hiint(res)));
  return __hiloint2double(hi, __double2loint(res));
}

template <int V, int NCH>
__global__ void __launch_bounds__(1024) probe(int iters, double* out, double scale) {
  extern __shared__ double stbl[];
  for (int i = threadIdx.x; i < fm::EXP_TBL_DOUBLES; i += blockDim.x) stbl[i] = exp2((double)(i >> 4) / 256.0);
  __syncthreads();
  const unsigned tbl = fm::lane_table(stbl);
  double c[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) c[i] = -1.0 - 0.37 * i - 1e-3 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) c[i] = exp_var<V>(c[i] * scale, tbl) - 1.5;   // 2 extra FP64 ops (DMUL, DADD)
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) s += c[i];
  if (s == 12345.678) out[0] = s;
}

template <int V, int NCH>
void run(double* d, int threads) {
  const int iters = 4000;
  cudaFuncSetAttribute(probe<V, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, fm::EXP_TBL_BYTES);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<V, NCH><<<148, threads, fm::EXP_TBL_BYTES>>>(100, d, 300.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  probe<V, NCH><<<148, threads, fm::EXP_TBL_BYTES>>>(iters, d, 300.0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps_per_smsp = threads / 32.0 / 4.0;
  printf("{\"variant\": %d, \"chains\": %d, \"warps_per_smsp\": %.0f, \"ms\": %.3f, \"nominal_cycles_per_warp_exp\": %.2f}\n",
         V, NCH, warps_per_smsp, ms, ms * 1e-3 * 1.965e9 / (warps_per_smsp * NCH * iters));
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  run<5, 8>(d, 512); run<5, 16>(d, 512); run<5, 8>(d, 1024);
  run<0, 8>(d, 512); run<1, 8>(d, 512); run<2, 8>(d, 512); run<3, 8>(d, 512); run<4, 8>(d, 512);
  run<0, 16>(d, 512); run<0, 8>(d, 1024); run<0, 4>(d, 1024); run<4, 16>(d, 512);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
