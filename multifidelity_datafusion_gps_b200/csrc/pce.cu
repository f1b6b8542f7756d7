// K9 polynomial-chaos projection on a quadrature grid.
//
// Replaces chaospy.fit_quadrature as the reference calls it (src/gpc/chaospy_wrapper.py:18-21) for
// the distributions the reference uses (independent uniforms, tests/test_mfgp_adapt_4d.py:40):
//   c_k = sum_q w_q f(x_q) phi_k(x_q),   phi_k(x) = prod_i sqrt(2 k_i + 1) P_{k_i}(2 (x_i - lb_i)/(ub_i - lb_i) - 1)
// (orthonormal Legendre basis: mean = c_0, variance = sum_{k>0} c_k^2).  The reference spends minutes
// in chaospy for this (tests/test_mfgp_adapt_4d.py:72-77); here the values f(x_q) are the GPU
// predictions already resident in HBM and the projection is one pass over them.
//
// A CTA stages a tile of NT nodes: the 1-D Legendre values of every node (recurrence, one thread
// per node) and w*f go to shared memory; then thread t owns the terms t, t+256, ... and walks the
// tile's nodes in order.  Tiles are assigned round-robin, partial sums are per CTA, and a second
// kernel adds the CTAs' partials in index order, so the result does not depend on the launch.
#include "common.cuh"

namespace {

constexpr int PCE_NT = 128;        // nodes per tile
constexpr int PCE_THREADS = 256;
constexpr int PCE_MAX_TERMS_PER_THREAD = 16;   // P <= 4096

__global__ void __launch_bounds__(PCE_THREADS)
    pce_project_kernel(const double* __restrict__ nodes, const double* __restrict__ wts,
                       const double* __restrict__ vals, long long Q, int d, int pmax,
                       const double* __restrict__ lb_scale /* [2d]: lb, 2/(ub-lb) */,
                       const int* __restrict__ mi /* (P, d) */, int P, long long ntiles,
                       double* __restrict__ partials /* (grid, P) */) {
  extern __shared__ double sm[];   // leg[NT][d][pmax+1] | wf[NT]
  const int stride = d * (pmax + 1);
  double* leg = sm;
  double* wf = sm + PCE_NT * stride;
  double acc[PCE_MAX_TERMS_PER_THREAD];
#pragma unroll
  for (int j = 0; j < PCE_MAX_TERMS_PER_THREAD; j++) acc[j] = 0.0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < PCE_NT) {
      const long long q = tile * PCE_NT + threadIdx.x;
      double* lq = leg + threadIdx.x * stride;
      if (q < Q) {
        wf[threadIdx.x] = wts[q] * vals[q];
        for (int i = 0; i < d; i++) {
          const double t = (nodes[q * d + i] - lb_scale[i]) * lb_scale[d + i] - 1.0;
          double p0 = 1.0, p1 = t;
          lq[i * (pmax + 1)] = 1.0;
          if (pmax >= 1) lq[i * (pmax + 1) + 1] = sqrt(3.0) * t;
          for (int n = 2; n <= pmax; n++) {   // n P_n = (2n-1) t P_{n-1} - (n-1) P_{n-2}
            const double p2 = ((2 * n - 1) * t * p1 - (n - 1) * p0) / n;
            lq[i * (pmax + 1) + n] = sqrt(2.0 * n + 1.0) * p2;
            p0 = p1;
            p1 = p2;
          }
        }
      } else {
        wf[threadIdx.x] = 0.0;
        for (int j = 0; j < stride; j++) lq[j] = 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PCE_MAX_TERMS_PER_THREAD; j++) {
      const int k = threadIdx.x + j * PCE_THREADS;
      if (k >= P) break;
      int deg[MFGP_MAX_D];
      for (int i = 0; i < d; i++) deg[i] = mi[k * d + i] + i * (pmax + 1);
      double a = acc[j];
      for (int nq = 0; nq < PCE_NT; nq++) {
        const double* lq = leg + nq * stride;
        double phi = wf[nq];
        for (int i = 0; i < d; i++) phi *= lq[deg[i]];
        a += phi;
      }
      acc[j] = a;
    }
  }
#pragma unroll
  for (int j = 0; j < PCE_MAX_TERMS_PER_THREAD; j++) {
    const int k = threadIdx.x + j * PCE_THREADS;
    if (k < P) partials[(long long)blockIdx.x * P + k] = acc[j];
  }
}

__global__ void pce_reduce_kernel(const double* __restrict__ partials, int nblocks, int P,
                                  double* __restrict__ coeff) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; b++) v += partials[(long long)b * P + k];
  coeff[k] = v;
}

}  // namespace

extern "C" size_t mfgp_pce_ws_bytes(int d, int P) {
  const size_t grid = 2 * MFGP_NUM_SMS;
  return (grid * (size_t)P + 2 * (size_t)d) * sizeof(double) + ((size_t)P * d * sizeof(int) + 15) / 16 * 16;
}

extern "C" int mfgp_pce_project(mfgp_handle_t h, const double* d_nodes, const double* h_lb,
                                const double* h_ub, int d, const double* d_weights,
                                const double* d_values, long long Q, const int* h_multi_index, int P,
                                int max_degree, double* d_coeff, double* h_coeff, double* d_ws,
                                size_t ws_bytes) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  ARG_CHECK(h, d_nodes && h_lb && h_ub && d_weights && d_values && h_multi_index && d_coeff && d_ws);
  ARG_CHECK(h, d >= 1 && d <= MFGP_MAX_D && Q >= 1 && max_degree >= 0 && max_degree <= 64);
  ARG_CHECK(h, P >= 1 && P <= PCE_THREADS * PCE_MAX_TERMS_PER_THREAD);
  ARG_CHECK(h, ws_bytes >= mfgp_pce_ws_bytes(d, P));
  for (int k = 0; k < P; k++)
    for (int i = 0; i < d; i++) ARG_CHECK(h, h_multi_index[k * d + i] >= 0 && h_multi_index[k * d + i] <= max_degree);
  const size_t smem = ((size_t)PCE_NT * d * (max_degree + 1) + PCE_NT) * sizeof(double);
  ARG_CHECK(h, smem <= 200 * 1024);
  const long long ntiles = (Q + PCE_NT - 1) / PCE_NT;
  const int grid = (int)(ntiles < 2 * MFGP_NUM_SMS ? ntiles : 2 * MFGP_NUM_SMS);
  double* partials = d_ws;
  double* d_lbs = d_ws + (size_t)2 * MFGP_NUM_SMS * P;
  int* d_mi = reinterpret_cast<int*>(d_lbs + 2 * d);
  double lbs[2 * MFGP_MAX_D];
  for (int i = 0; i < d; i++) {
    ARG_CHECK(h, h_ub[i] > h_lb[i]);
    lbs[i] = h_lb[i];
    lbs[d + i] = 2.0 / (h_ub[i] - h_lb[i]);
  }
  // the host tables are consumed by the copies before this call returns (pageable -> staged copy)
  CUDA_TRY(h, cudaMemcpyAsync(d_lbs, lbs, 2 * d * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(d_mi, h_multi_index, (size_t)P * d * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaFuncSetAttribute(pce_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(h, PC_MISC);
  pce_project_kernel<<<grid, PCE_THREADS, smem, h->stream>>>(d_nodes, d_weights, d_values, Q, d, max_degree, d_lbs,
                                                            d_mi, P, ntiles, partials);
  prof_end(h, PC_MISC);
  LAUNCH_CHECK(h);
  pce_reduce_kernel<<<(P + 127) / 128, 128, 0, h->stream>>>(partials, grid, P, d_coeff);
  LAUNCH_CHECK(h);
  if (h_coeff) {
    CUDA_TRY(h, cudaMemcpyAsync(h_coeff, d_coeff, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  }
  return 0;
}
