// K1 covariance assembly and K5 gradient reduction.
//
// K1 replaces kern.K(X) for the reference's two kernels (src/abstractMFGP.py:59-60 and :62-80:
// RBF(z)*RBF(x) + RBF(x)) plus GPy's diag.add(Ky, noise + 1e-8): one fused pass, no N x N
// temporaries (GPy materialises three distance matrices and ~6 elementwise temporaries).
// The product k1*k2 is folded into a single exponential, so a composite element costs two exps.
// HBM-bound by contract (8 bytes written per element, inputs (N x D) stay in shared memory), but on
// B200 the FP64 pipe is the tighter bound for the exponentials; fastmath.cuh's exp_neg (10 FP64
// instructions + one table lookup) is what keeps the kernel near the memory roofline.
//
// K5 replaces GPy's update_gradients_full chain (stationary.py / prod.py / add.py) for
// dL_dK = 0.5 (alpha alpha^T - K^-1): one streaming pass over K^-1's lower triangle that
// recomputes the kernel factors from X on the fly and produces six sums
//   S0 = sum G K12, S1 = sum G K12 rz2, S2 = sum G K12 rx2, S3 = sum G K3, S4 = sum G K3 rx2, S5 = tr G
// in a fixed order (persistent blocks, fixed tile->block map, two-stage reduction).
#include "common.cuh"
#include "fastmath.cuh"

namespace {

constexpr int AT = 64;   // tile edge

__device__ __forceinline__ void tile_from_linear(int tt, int& ti, int& tj) {
  ti = (int)((sqrt(8.0 * tt + 1.0) - 1.0) * 0.5);
  while ((long)(ti + 1) * (ti + 2) / 2 <= tt) ti++;
  while ((long)ti * (ti + 1) / 2 > tt) ti--;
  tj = tt - ti * (ti + 1) / 2;
}

constexpr int BT = 128;  // CTA tile edge of the assembly / gradient kernels: 2x2 sub-blocks of AT

// sX[dd][r] = X[row0 + r][dd] for r < TR (0 beyond N); dd-major so that a row of threads reads
// consecutive shared-memory words
template <int TR>
__device__ __forceinline__ void load_rows(double* sX, const double* __restrict__ X, int N, int D, int row0) {
  for (int r = threadIdx.x / 8; r < TR; r += blockDim.x / 8) {     // 8 threads per input row
    const int gr = row0 + r;
    for (int dd = threadIdx.x & 7; dd < D; dd += 8) sX[dd * TR + r] = gr < N ? X[(long)gr * D + dd] : 0.0;
  }
}

// squared distances of the thread's 4x4 sub-block: rows ty+16i, columns 4tx+j
template <int TR>
__device__ __forceinline__ void sub_block_dist(const double* sXi, const double* sXj, int D, int d, int tx,
                                               int ty, double (&rx)[4][4], double (&rz)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) rx[i][j] = rz[i][j] = 0.0;
  for (int dd = 0; dd < D; dd++) {
    double xi[4], xj[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xi[i] = sXi[dd * TR + ty + 16 * i];          // broadcast within a half-warp
    const double2 p0 = *reinterpret_cast<const double2*>(sXj + dd * TR + 4 * tx);   // 16-byte loads: no conflicts
    const double2 p1 = *reinterpret_cast<const double2*>(sXj + dd * TR + 4 * tx + 2);
    xj[0] = p0.x; xj[1] = p0.y; xj[2] = p1.x; xj[3] = p1.y;
    if (dd < d) {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double t = xi[i] - xj[j];
          rx[i][j] = fma(t, t, rx[i][j]);
        }
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double t = xi[i] - xj[j];
          rz[i][j] = fma(t, t, rz[i][j]);
        }
    }
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256, 2)
    assemble_kernel(KParams kp, const double* __restrict__ X, int N, double diag_add,
                    double* __restrict__ K, long ldk, int lower_only, int tiles, int nrows) {
  extern __shared__ __align__(16) double dsm[];   // sXi[D][BT] | sXj[D][BT]
  __shared__ double stbl_all[fm::EXP_TBL_DOUBLES];
  int ti, tj;
  if (lower_only) {
    tile_from_linear(blockIdx.x, ti, tj);
  } else {
    ti = blockIdx.x / tiles;
    tj = blockIdx.x % tiles;
  }
  const int D = kp.D, d = kp.d;
  double* sXi = dsm;
  double* sXj = dsm + D * BT;
  fm::load_exp_table(stbl_all);
  const double* stbl = stbl_all + (threadIdx.x & 15);
  load_rows<BT>(sXi, X, N, D, ti * BT);
  load_rows<BT>(sXj, X, N, D, tj * BT);
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool has3 = kp.s3 != 0.0;
#pragma unroll 1
  for (int sb = 0; sb < 4; sb++) {
    const int si = sb >> 1, sj = sb & 1;
    const int row0 = ti * BT + si * AT, col0 = tj * BT + sj * AT;
    if (row0 >= nrows || col0 >= nrows) continue;
    if (lower_only && col0 > row0) continue;        // sub-block strictly above the diagonal
    double rx[4][4], rz[4][4];
    sub_block_dist<BT>(sXi + si * AT, sXj + sj * AT, D, d, tx, ty, rx, rz);
    if (row0 + AT <= N && col0 + AT <= N) {
      // Interior sub-block (almost all of them): straight-line code, no per-element predicates, so
      // the 32 exponentials of a thread are scheduled together and their constants stay in registers.
      const double c12 = kp.c12, az = kp.az, ax = kp.ax, s3 = kp.s3, a3 = kp.a3;
      double v[4][4];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) v[i][j] = c12 * fm::exp_neg(fma(az, rz[i][j], ax * rx[i][j]), stbl);
      if (has3) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) v[i][j] = fma(s3, fm::exp_neg(a3 * rx[i][j], stbl), v[i][j]);
      }
      if (row0 == col0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (ty + 16 * i == 4 * tx + j) v[i][j] += diag_add;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        double* dst = K + (long)(row0 + ty + 16 * i) * ldk + col0 + 4 * tx;
        if (VEC) {
          reinterpret_cast<double2*>(dst)[0] = make_double2(v[i][0], v[i][1]);
          reinterpret_cast<double2*>(dst)[1] = make_double2(v[i][2], v[i][3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) dst[j] = v[i][j];
        }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = row0 + ty + 16 * i;
      if (r >= nrows) continue;
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = col0 + 4 * tx + j;
        double val;
        if (r < N && c < N) {
          val = kp.c12 * fm::exp_neg(fma(kp.az, rz[i][j], kp.ax * rx[i][j]), stbl);
          if (has3) val = fma(kp.s3, fm::exp_neg(kp.a3 * rx[i][j], stbl), val);
          if (r == c) val += diag_add;
        } else {
          val = (r == c) ? 1.0 : 0.0;   // identity pad block
        }
        v[j] = val;
      }
      double* dst = K + (long)r * ldk + col0 + 4 * tx;
      if (VEC) {
        // nrows is a multiple of 64 here (padded buffer): no column guard needed
        reinterpret_cast<double2*>(dst)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2*>(dst)[1] = make_double2(v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (col0 + 4 * tx + j < nrows) dst[j] = v[j];
      }
    }
  }
}

constexpr int GR_BLOCKS = MFGP_NUM_SMS * 4;

__global__ void __launch_bounds__(256, 2)
    grad_reduce_kernel(KParams kp, const double* __restrict__ X, int N,
                       const double* __restrict__ Kinv, long ld, const double* __restrict__ alpha,
                       int ntiles_lin, double* __restrict__ partials) {
  extern __shared__ __align__(16) double dsm[];   // sXi[D][BT] | sXj[D][BT]
  __shared__ __align__(16) double sAi_t[BT], sAj_t[BT];
  __shared__ double red[8][6];
  __shared__ double stbl_all[fm::EXP_TBL_DOUBLES];
  const int D = kp.D, d = kp.d;
  double* sXi_t = dsm;
  double* sXj_t = dsm + D * BT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool has3 = kp.s3 != 0.0;
  fm::load_exp_table(stbl_all);
  const double* stbl = stbl_all + (threadIdx.x & 15);
  double S[6] = {0, 0, 0, 0, 0, 0};
  for (int tt = blockIdx.x; tt < ntiles_lin; tt += gridDim.x) {
    int ti, tj;
    tile_from_linear(tt, ti, tj);
    __syncthreads();
    load_rows<BT>(sXi_t, X, N, D, ti * BT);
    load_rows<BT>(sXj_t, X, N, D, tj * BT);
    if (threadIdx.x < BT) {
      int r = ti * BT + threadIdx.x;
      sAi_t[threadIdx.x] = r < N ? alpha[r] : 0.0;
    } else {
      int c = tj * BT + threadIdx.x - BT;
      sAj_t[threadIdx.x - BT] = c < N ? alpha[c] : 0.0;
    }
    __syncthreads();
#pragma unroll 1
   for (int sb = 0; sb < 4; sb++) {
    const int si = sb >> 1, sj = sb & 1;
    const int row0 = ti * BT + si * AT, col0 = tj * BT + sj * AT;
    if (row0 >= N || col0 > row0) continue;         // beyond the data / strictly above the diagonal
    const double* sXi = sXi_t + si * AT;
    const double* sXj = sXj_t + sj * AT;
    const double* sAi = sAi_t + si * AT;
    const double* sAj = sAj_t + sj * AT;
    // issue the K^-1 loads first so that they overlap the distance computation
    double kin[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = row0 + ty + 16 * i;
      if (r < N) {
        const double2* src = reinterpret_cast<const double2*>(Kinv + (long)r * ld + col0 + 4 * tx);
        const double2 q0 = src[0], q1 = src[1];
        kin[i][0] = q0.x; kin[i][1] = q0.y; kin[i][2] = q1.x; kin[i][3] = q1.y;
      } else {
        kin[i][0] = kin[i][1] = kin[i][2] = kin[i][3] = 0.0;
      }
    }
    double rx[4][4], rz[4][4];
    sub_block_dist<BT>(sXi, sXj, D, d, tx, ty, rx, rz);
    const double2 a0 = *reinterpret_cast<const double2*>(sAj + 4 * tx);
    const double2 a1 = *reinterpret_cast<const double2*>(sAj + 4 * tx + 2);
    const double aj[4] = {a0.x, a0.y, a1.x, a1.y};
    if (col0 < row0 && row0 + AT <= N) {
      // interior off-diagonal sub-block: every element counts twice (weight 1), no predicates
      const double c12 = kp.c12, az = kp.az, ax = kp.ax, s3 = kp.s3, a3 = kp.a3;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const double ai = sAi[ty + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const double G = ai * aj[j] - kin[i][j];
          const double gk = G * (c12 * fm::exp_neg(fma(az, rz[i][j], ax * rx[i][j]), stbl));
          S[0] += gk;
          S[1] = fma(gk, rz[i][j], S[1]);
          S[2] = fma(gk, rx[i][j], S[2]);
          if (has3) {
            const double g3 = G * s3 * fm::exp_neg(a3 * rx[i][j], stbl);
            S[3] += g3;
            S[4] = fma(g3, rx[i][j], S[4]);
          }
        }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = row0 + ty + 16 * i;
      if (r >= N) continue;
      const double ai = sAi[ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = col0 + 4 * tx + j;
        if (c > r || c >= N) continue;
        const double w = (c == r) ? 0.5 : 1.0;   // G = 0.5(aa^T - Kinv); off-diagonal counted twice
        const double G = w * (ai * aj[j] - kin[i][j]);
        const double k12 = kp.c12 * fm::exp_neg(fma(kp.az, rz[i][j], kp.ax * rx[i][j]), stbl);
        const double gk = G * k12;
        S[0] += gk;
        S[1] = fma(gk, rz[i][j], S[1]);
        S[2] = fma(gk, rx[i][j], S[2]);
        if (has3) {
          const double g3 = G * kp.s3 * fm::exp_neg(kp.a3 * rx[i][j], stbl);
          S[3] += g3;
          S[4] = fma(g3, rx[i][j], S[4]);
        }
        if (c == r) S[5] += G;
      }
    }
   }
  }
  // block reduction in a fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; q++) {
    double v = S[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += red[w][threadIdx.x];
    partials[blockIdx.x * 8 + threadIdx.x] = v;
  }
}

__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nblocks,
                                       double* __restrict__ out8) {
  // 6 warps, one quantity each; lanes stride over blocks, then a shuffle tree: fixed order
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (q >= 6) return;
  double v = 0.0;
  for (int b = lane; b < nblocks; b += 32) v += partials[b * 8 + q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) out8[q] = v;
}

}  // namespace

int assemble_configure(mfgp_ctx* h) {
  const int smem = 2 * MFGP_MAX_D * BT * (int)sizeof(double);   // 64 KB at the maximum input width
  CUDA_TRY(h, cudaFuncSetAttribute(assemble_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(h, cudaFuncSetAttribute(assemble_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CUDA_TRY(h, cudaFuncSetAttribute(grad_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return 0;
}

int assemble_launch(mfgp_ctx* h, const KParams& kp, const double* X, int N, double diag_add,
                    double* K, long long ldk, int uplo, int npad_identity) {
  const int nrows = npad_identity > 0 ? npad_identity : N;
  const int tiles = (nrows + BT - 1) / BT;
  const int lower = uplo == MFGP_UPLO_LOWER;
  const int grid = lower ? tiles * (tiles + 1) / 2 : tiles * tiles;
  const bool vec = (nrows % AT == 0) && (ldk % 2 == 0) && ((uintptr_t)K % 16 == 0);
  const size_t smem = (size_t)2 * kp.D * BT * sizeof(double);
  prof_begin(h, PC_ASSEMBLE);
  if (vec)
    assemble_kernel<true><<<grid, 256, smem, h->stream>>>(kp, X, N, diag_add, K, ldk, lower, tiles, nrows);
  else
    assemble_kernel<false><<<grid, 256, smem, h->stream>>>(kp, X, N, diag_add, K, ldk, lower, tiles, nrows);
  prof_end(h, PC_ASSEMBLE);
  LAUNCH_CHECK(h);
  return 0;
}

int grad_reduce_launch(mfgp_ctx* h, const KParams& kp, const double* X, int N, const double* Kinv,
                       long long ld, const double* alpha, double* d_out8) {
  const int tiles = (N + BT - 1) / BT;
  const int nlin = tiles * (tiles + 1) / 2;
  const int grid = nlin < GR_BLOCKS ? nlin : GR_BLOCKS;
  const size_t smem = (size_t)2 * kp.D * BT * sizeof(double);
  prof_begin(h, PC_GRAD);
  grad_reduce_kernel<<<grid, 256, smem, h->stream>>>(kp, X, N, Kinv, ld, alpha, nlin, h->d_partials);
  prof_end(h, PC_GRAD);
  LAUNCH_CHECK(h);
  reduce_partials_kernel<<<1, 192, 0, h->stream>>>(h->d_partials, grid, d_out8);
  LAUNCH_CHECK(h);
  return 0;
}
