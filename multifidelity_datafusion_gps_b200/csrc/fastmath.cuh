// FP64 exp for non-positive arguments, tuned for the covariance kernels.
//
// Covariance assembly on B200 is bound by the FP64 pipe, not by HBM (ncu: 127 instructions per element
// with CUDA's exp(), fp64 pipe 40 % active, dram 14 %), so the exponential is the thing to shrink:
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128
//   exp(x) = 2^m * 2^(j/64) * exp(r),  exp(r) by a degree-5 polynomial (|r|^6/720 < 4e-17)
// 10 FP64 instructions (CUDA's exp: ~17 FP64 + special-case handling) plus one 8-byte shared-memory
// table lookup.  Error <= ~1 ulp (table entries are correctly rounded, generated with exp2l()).
// Arguments below -708 return 0 (the true value is < 3e-308).
#pragma once
#include <cuda_runtime.h>

namespace fm {

// (plain device memory: every thread reads a different entry, which a __constant__ bank would serialise)
static __device__ const double c_exp2_tbl[64] = {
    1, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284,
    1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199,
    1.0905077326652577, 1.1023825833078409, 1.1143867425958924, 1.1265216186082418,
    1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812,
    1.189207115002721, 1.2021567314527031, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.2553807570246911, 1.2690509571917332, 1.2828700160787783,
    1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.3396675240533029,
    1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112,
    1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647,
    1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.5590044002378369, 1.5759808451078865, 1.593142151342267,
    1.6104903319492543, 1.6280274218573478, 1.6457554781539649, 1.6636765803267364,
    1.681792830507429, 1.7001063537185235, 1.7186192981224779, 1.7373338352737062,
    1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989,
    1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.9784560263879509};

// The lookup index differs from lane to lane, so a plain 64-entry shared table would serialise on
// bank conflicts (ncu: the LSU was as busy as the FP64 pipe).  The table is therefore replicated 16
// times, copy r living in bank pair r: entry j of copy r is at word j*16 + r, and a lane always reads
// copy (lane & 15) -> the 16 lanes of a half-warp hit 16 different bank pairs whatever their j.
constexpr int EXP_TBL_DOUBLES = 64 * 16;   // 8 KB

// fill the replicated table (call from every thread of the block, then __syncthreads())
__device__ __forceinline__ void load_exp_table(double* stbl) {
  for (int i = threadIdx.x; i < EXP_TBL_DOUBLES; i += blockDim.x) stbl[i] = c_exp2_tbl[i >> 4];
}

// stbl_lane = stbl + (lane & 15)
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ stbl_lane) {
  const double MAGIC = 6755399441055744.0;               // 1.5 * 2^52: round-to-nearest-integer trick
  const double t = fma(x, 92.332482616893657, MAGIC);    // x * 64/ln2
  const int n = __double2loint(t);
  const double nd = t - MAGIC;
  double r = fma(nd, -0.010830424667801708, x);          // ln2/64, high part (29 significant bits)
  r = fma(nd, -2.8447437476806321e-11, r);               // low part
  double h = fma(r, 8.3333333333333332e-3, 4.1666666666666664e-2);
  h = fma(h, r, 1.6666666666666666e-1);
  h = fma(h, r, 0.5);
  h = fma(h, r, 1.0);
  const double T = stbl_lane[(n & 63) << 4];
  double res = fma(T, h * r, T);
  const int hi = __double2hiint(res) + ((n >> 6) << 20);  // scale by 2^m
  res = __hiloint2double(hi, __double2loint(res));
  return x < -708.0 ? 0.0 : res;
}

}  // namespace fm
