// Blocked FP64 Cholesky (K2), triangular inverse, K^-1 = W^T W (K4) and the O(N^2) solves (K3).
// Replaces GPy.util.linalg.pdinv / dpotrs / dpotri (LAPACK on the CPU) as reached from
// GPRegression construction and every optimiser step (reference src/MFDataFusion.py:93-100,
// src/abstractMFGP.py:131-137).  Everything is lower / row-major on an npad x npad buffer whose
// pad block is the identity (npad multiple of 128).
//
// Recursive formulation so that almost all flops are large DMMA GEMMs (gemm.cuh):
//   potrf(n):  potrf(A11); A21 <- A21 L11^-T (recursive TRSM, leaf = multiply by the 128x128 leaf
//              inverse); A22 -= A21 A21^T (SYRK, lower tiles); potrf(A22)
//   trtri(n):  trtri(11), trtri(22); W21 = -W22 L21 W11
//   lauum:     Kinv = W^T W, one launch over the lower tiles with k >= row0
#include "gemm.cuh"

namespace {

constexpr int LEAF = 128;
constexpr int LEAF_LD = LEAF + 1;

// One CTA factorises a 128x128 diagonal block in shared memory and inverts the factor.
// L is kept in the lower triangle of S, the columns of W = L^-1 are built in the upper triangle.
__global__ void __launch_bounds__(256, 1)
    leaf_potrf_inv_kernel(double* A, long lda, double* W, long ldw, int* info, int j0) {
  extern __shared__ double S[];   // LEAF x LEAF_LD
  __shared__ double dinv[LEAF], dsq[LEAF];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < LEAF * LEAF; idx += 256) {
    int i = idx >> 7, j = idx & 127;
    S[i * LEAF_LD + j] = A[(long)i * lda + j];
  }
  __syncthreads();
  for (int k = 0; k < LEAF; k++) {
    __syncthreads();   // trailing update of step k-1 is complete
    double p = S[k * LEAF_LD + k];
    if (!(p > 0.0)) {   // also catches NaN
      if (tid == 0 && info[0] == 0) info[0] = j0 + k + 1;
      p = 1.0;
    }
    const double lkk = sqrt(p);
    const double rinv = 1.0 / lkk;
    if (tid == 0) {
      dsq[k] = lkk;     // S[k][k] itself is left untouched: other threads may still be reading it
      dinv[k] = rinv;
    }
    for (int i = k + 1 + tid; i < LEAF; i += 256) S[i * LEAF_LD + k] *= rinv;
    __syncthreads();
    // trailing update of the lower triangle: S[i][j] -= S[i][k] * S[j][k], k < j <= i
    const int n = LEAF - k - 1;
    for (int ii = tid >> 4; ii < n; ii += 16) {
      const int i = k + 1 + ii;
      const double lik = S[i * LEAF_LD + k];
      for (int jj = tid & 15; jj <= ii; jj += 16) {
        const int j = k + 1 + jj;
        S[i * LEAF_LD + j] -= lik * S[j * LEAF_LD + k];
      }
    }
  }
  __syncthreads();
  // inverse: thread c builds column c of W by forward substitution, stored at S[c][i], i > c
  if (tid < LEAF) {
    const int c = tid;
    const double wcc = dinv[c];
    for (int i = c + 1; i < LEAF; i++) {
      double s0 = S[i * LEAF_LD + c] * wcc, s1 = 0.0;
      int k = c + 1;
      for (; k + 1 < i; k += 2) {
        s0 = fma(S[i * LEAF_LD + k], S[c * LEAF_LD + k], s0);
        s1 = fma(S[i * LEAF_LD + k + 1], S[c * LEAF_LD + k + 1], s1);
      }
      if (k < i) s0 = fma(S[i * LEAF_LD + k], S[c * LEAF_LD + k], s0);
      S[c * LEAF_LD + i] = -(s0 + s1) * dinv[i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < LEAF * LEAF; idx += 256) {
    int i = idx >> 7, j = idx & 127;
    double l, w;
    if (j < i) {
      l = S[i * LEAF_LD + j];
      w = S[j * LEAF_LD + i];
    } else if (j == i) {
      l = dsq[i];
      w = dinv[i];
    } else {
      l = 0.0;
      w = 0.0;
    }
    A[(long)i * lda + j] = l;
    W[(long)i * ldw + j] = w;
  }
}

template <bool A_KC, bool B_KC>
int launch_gemm(mfgp_ctx* h, const dg::GemmParams& p) {
  int tiles = p.lower_only ? p.tiles_m * (p.tiles_m + 1) / 2 : p.tiles_m * p.tiles_n;
  if (tiles <= 0) return 0;
  dg::gemm_kernel<A_KC, B_KC><<<tiles, dg::THREADS, dg::SMEM_BYTES, h->stream>>>(p);
  LAUNCH_CHECK(h);
  return 0;
}

dg::GemmParams gp(const double* A, long lda, const double* B, long ldb, double* C, long ldc, int m,
                  int n, int k, double alpha, double beta) {
  dg::GemmParams p;
  memset(&p, 0, sizeof(p));
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc;
  p.tiles_m = m / dg::BM; p.tiles_n = n / dg::BN; p.K = k;
  p.alpha = alpha; p.beta = beta;
  return p;
}

inline int split(int n) { return (n / LEAF / 2) * LEAF; }   // first half, multiple of the leaf

// X (m x n at A[r0][j0]) <- X * L^-T, L = A[j0:j0+n, j0:j0+n]; leaf inverses are in W's diagonal blocks
int trsm_rec(mfgp_ctx* h, double* A, double* W, long ld, int r0, int m, int j0, int n) {
  if (n == LEAF) {
    double* X = A + (long)r0 * ld + j0;
    const double* Wl = W + (long)j0 * ld + j0;
    // X[r][c] = sum_k X[r][k] * Wl[c][k]; in place is safe: a CTA owns all 128 columns of its rows
    return launch_gemm<true, true>(h, gp(X, ld, Wl, ld, X, ld, m, LEAF, LEAF, 1.0, 0.0));
  }
  int n1 = split(n), n2 = n - n1, rc;
  if ((rc = trsm_rec(h, A, W, ld, r0, m, j0, n1))) return rc;
  // X2 -= X1 * L21^T,  L21 = A[j0+n1 : j0+n, j0 : j0+n1]
  rc = launch_gemm<true, true>(h, gp(A + (long)r0 * ld + j0, ld, A + (long)(j0 + n1) * ld + j0, ld,
                                     A + (long)r0 * ld + j0 + n1, ld, m, n2, n1, -1.0, 1.0));
  if (rc) return rc;
  return trsm_rec(h, A, W, ld, r0, m, j0 + n1, n2);
}

int potrf_rec(mfgp_ctx* h, double* A, double* W, long ld, int j0, int n) {
  if (n == LEAF) {
    const int smem = LEAF * LEAF_LD * 8;
    leaf_potrf_inv_kernel<<<1, 256, smem, h->stream>>>(A + (long)j0 * ld + j0, ld,
                                                       W + (long)j0 * ld + j0, ld, h->d_info, j0);
    LAUNCH_CHECK(h);
    return 0;
  }
  int n1 = split(n), n2 = n - n1, rc;
  if ((rc = potrf_rec(h, A, W, ld, j0, n1))) return rc;
  if ((rc = trsm_rec(h, A, W, ld, j0 + n1, n2, j0, n1))) return rc;
  // A22 -= L21 L21^T on the lower tiles
  {
    const double* L21 = A + (long)(j0 + n1) * ld + j0;
    dg::GemmParams p = gp(L21, ld, L21, ld, A + (long)(j0 + n1) * ld + j0 + n1, ld, n2, n2, n1, -1.0, 1.0);
    p.lower_only = 1;
    if ((rc = launch_gemm<true, true>(h, p))) return rc;
  }
  return potrf_rec(h, A, W, ld, j0 + n1, n2);
}

int trtri_rec(mfgp_ctx* h, const double* L, double* W, long ld, int j0, int n) {
  if (n == LEAF) return 0;   // leaf inverse already written by potrf
  int n1 = split(n), n2 = n - n1, rc;
  if ((rc = trtri_rec(h, L, W, ld, j0, n1))) return rc;
  if ((rc = trtri_rec(h, L, W, ld, j0 + n1, n2))) return rc;
  const double* W11 = W + (long)j0 * ld + j0;
  const double* W22 = W + (long)(j0 + n1) * ld + j0 + n1;
  const double* L21 = L + (long)(j0 + n1) * ld + j0;
  double* Tt = W + (long)j0 * ld + j0 + n1;          // n1 x n2 scratch in W's (unused) upper block
  double* W21 = W + (long)(j0 + n1) * ld + j0;
  // Tt[c][r] = sum_{k>=c} W11[k][c] * L21[r][k]         (Tt = (L21 W11)^T)
  {
    dg::GemmParams p = gp(W11, ld, L21, ld, Tt, ld, n1, n2, n1, 1.0, 0.0);
    p.kb_row = 1;
    if ((rc = launch_gemm<false, true>(h, p))) return rc;
  }
  // W21[r][c] = -sum_{k<=r} W22[r][k] * Tt[c][k]
  {
    dg::GemmParams p = gp(W22, ld, Tt, ld, W21, ld, n2, n1, n2, -1.0, 0.0);
    p.ke_row = 1;
    if ((rc = launch_gemm<true, true>(h, p))) return rc;
  }
  return 0;
}

}  // namespace

int linalg_configure(mfgp_ctx* h) {
  CUDA_TRY(h, cudaFuncSetAttribute(dg::gemm_kernel<true, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CUDA_TRY(h, cudaFuncSetAttribute(dg::gemm_kernel<false, true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CUDA_TRY(h, cudaFuncSetAttribute(dg::gemm_kernel<false, false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CUDA_TRY(h, cudaFuncSetAttribute(dg::trmm_sumsq_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, dg::SMEM_BYTES));
  CUDA_TRY(h, cudaFuncSetAttribute(leaf_potrf_inv_kernel,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, LEAF * LEAF_LD * 8));
  return 0;
}

int potrf_padded(mfgp_ctx* h, double* A, double* W, int npad) {
  ARG_CHECK(h, npad > 0 && npad % LEAF == 0);
  CUDA_TRY(h, cudaMemsetAsync(h->d_info, 0, 4 * sizeof(int), h->stream));
  return potrf_rec(h, A, W, npad, 0, npad);
}

int trtri_padded(mfgp_ctx* h, const double* L, double* W, int npad) {
  ARG_CHECK(h, npad > 0 && npad % LEAF == 0);
  return trtri_rec(h, L, W, npad, 0, npad);
}

int lauum_padded(mfgp_ctx* h, const double* W, double* Kinv, int npad) {
  ARG_CHECK(h, npad > 0 && npad % LEAF == 0);
  // Kinv[i][j] = sum_{k >= max(i,j)} W[k][i] * W[k][j]; lower tiles have row0 >= col0
  dg::GemmParams p = gp(W, npad, W, npad, Kinv, npad, npad, npad, npad, 1.0, 0.0);
  p.lower_only = 1;
  p.kb_row = 1;
  return launch_gemm<false, false>(h, p);
}

int trmm_sumsq(mfgp_ctx* h, const double* W, int npad, const double* Ks, long long cols_pad,
               double* out_ss) {
  ARG_CHECK(h, npad % LEAF == 0 && cols_pad % dg::BN == 0);
  if (cols_pad == 0) return 0;
  dg::trmm_sumsq_kernel<<<(unsigned)(cols_pad / dg::BN), dg::THREADS, dg::SMEM_BYTES, h->stream>>>(
      W, npad, Ks, out_ss);
  LAUNCH_CHECK(h);
  return 0;
}
