// FP64 exponential for the covariance kernels, shaped around what bounds them on B200.
//
// Covariance assembly is FP64-pipe / issue bound, not HBM bound (ncu, r01: fp64 pipe 51 %, issue
// 77 %, dram 22 %), so the exponential is the thing to shrink -- both its FP64 instructions (2 issue
// cycles each per warp) and the integer/select instructions that compete for the same issue slots.
//
// The caller passes the argument PRE-SCALED, u = x * 256/ln2 (the kernel hyper-parameters are folded
// into the scaled coefficients on the host, including log(variance), see make_kparams):
//   n = rint(u) = 256 m + j,  f = u - n  in [-1/2, 1/2]
//   exp(x) = 2^m * 2^(j/256) * exp(f ln2/256),  exp(f ln2/256) - 1 = f (c1 + c2 f + c3 f^2 + c4 f^3)
// (truncation (ln2/512)^5/120 = 3.8e-17 relative).  8 FP64 instructions (CUDA's exp(): ~17 plus
// special-case handling), 1 shared-memory lookup and 5 integer instructions:
//   - no range-reduction multiplies (u is already in table units; n*1 is exact),
//   - underflow is handled by clamping the HIGH WORD of u as an unsigned integer (one VIMNMX; for
//     negative doubles the high word grows with |u|), so that m >= -1022: arguments below
//     -708.4 return a positive number < 4.5e-308 instead of 0 (absolute error < 4.5e-308),
//   - 2^m is applied by an integer add on the exponent field.
// Error <= ~1 ulp of the result plus the rounding of u itself (|x| * 2^-53 relative).
// Valid for u < 2^31 - 1 (x < 4.0e6; make_kparams rejects log-variances beyond +-700).
#pragma once
#include <cuda_runtime.h>

namespace fm {

constexpr int TBL_BITS = 8;
constexpr int TBL_N = 1 << TBL_BITS;                 // entries 2^(j/256), correctly rounded (host exp2l)
constexpr int TBL_REP = 16;
// The lookup index differs from lane to lane, so a plain table would serialise on bank conflicts
// (r01 ncu: the LSU was as busy as the FP64 pipe).  The table is therefore replicated 16 times,
// copy r living in bank pair r: entry j of copy r is at word j*16 + r, and a lane always reads copy
// (lane & 15) -> the 16 lanes of a half-warp hit 16 different bank pairs whatever their j.
constexpr int EXP_TBL_DOUBLES = TBL_N * TBL_REP;     // 32 KB
constexpr int EXP_TBL_BYTES = EXP_TBL_DOUBLES * 8;
constexpr double U_PER_X = 369.32993046757463;       // 256 / ln 2

// fill the replicated table from the 256-entry global one (every thread of the block calls this,
// then __syncthreads())
__device__ __forceinline__ void load_exp_table(double* stbl, const double* __restrict__ gtbl) {
  for (int i = threadIdx.x; i < EXP_TBL_DOUBLES; i += blockDim.x) stbl[i] = gtbl[i >> 4];
}

// shared-window byte address of this lane's table copy
__device__ __forceinline__ unsigned lane_table(const double* stbl) {
  return (unsigned)__cvta_generic_to_shared(stbl) + 8u * (threadIdx.x & 15);
}

// exp(u * ln2/256); tbl = lane_table(...)
__device__ __forceinline__ double exp2s(double u, unsigned tbl) {
  const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
  {  // u = max(u, -1022*256) for u <= 0, on the high word only (the low word perturbs u by < 2^-15)
    const unsigned hi = min((unsigned)__double2hiint(u), 0xC10FF000u);
    u = __hiloint2double((int)hi, __double2loint(u));
  }
  const double t = u + MAGIC;
  const int n = __double2loint(t);
  const double f = u - (t - MAGIC);
  double h = fma(f, 2.239395190875157e-12, 3.3083026805413713e-09);   // c4, c3
  h = fma(h, f, 3.6655655969101062e-06);                               // c2
  h = fma(h, f, 2.7076061740622863e-03);                               // c1
  // index and scale as mask + multiply-add pairs (LOP3 + IMAD each); left to itself the compiler
  // emits shift, mask, add triples
  double T;
  asm("{\n\t.reg .u32 j, a;\n\t"
      "and.b32 j, %1, 255;\n\t"
      "mad.lo.u32 a, j, 128, %2;\n\t"
      "ld.shared.f64 %0, [a];\n\t}"
      : "=d"(T) : "r"(n), "r"(tbl));
  const double res = fma(T, h * f, T);
  int hi;   // hi(res) + m * 2^20, m = n >> 8
  asm("{\n\t.reg .u32 q;\n\t"
      "and.b32 q, %1, 0xffffff00;\n\t"
      "mad.lo.u32 %0, q, 4096, %2;\n\t}"
      : "=r"(hi) : "r"(n), "r"(__double2hiint(res)));
  return __hiloint2double(hi, __double2loint(res));
}

// Same arithmetic with a plain 256-entry table (entry j at byte 8*j of tbl_flat): for single-CTA kernels on
// tiny problems, where staging the 32 KB bank-replicated copy costs more than the bank conflicts it avoids.
// Bit-identical results to exp2s.
__device__ __forceinline__ double exp2s_flat(double u, unsigned tbl_flat) {
  const double MAGIC = 6755399441055744.0;
  {
    const unsigned hi = min((unsigned)__double2hiint(u), 0xC10FF000u);
    u = __hiloint2double((int)hi, __double2loint(u));
  }
  const double t = u + MAGIC;
  const int n = __double2loint(t);
  const double f = u - (t - MAGIC);
  double h = fma(f, 2.239395190875157e-12, 3.3083026805413713e-09);
  h = fma(h, f, 3.6655655969101062e-06);
  h = fma(h, f, 2.7076061740622863e-03);
  double T;
  asm("{\n\t.reg .u32 j, a;\n\t"
      "and.b32 j, %1, 255;\n\t"
      "mad.lo.u32 a, j, 8, %2;\n\t"
      "ld.shared.f64 %0, [a];\n\t}"
      : "=d"(T) : "r"(n), "r"(tbl_flat));
  const double res = fma(T, h * f, T);
  int hi;
  asm("{\n\t.reg .u32 q;\n\t"
      "and.b32 q, %1, 0xffffff00;\n\t"
      "mad.lo.u32 %0, q, 4096, %2;\n\t}"
      : "=r"(hi) : "r"(n), "r"(__double2hiint(res)));
  return __hiloint2double(hi, __double2loint(res));
}

}  // namespace fm
