// Batched transition-matrix builder (K1): every branch x rate category in one launch.
//
// Restates get_prob_t / get_edge_transition_mat (mcmc_gamma.pyx:372-401, 439-482) and the
// closed forms ptJC (:507-514), ptF81 (:527-547), binaryptF81 (:516-525).  The reference
// evaluates those in plain C doubles without FMA contraction (x86-64 baseline), so the
// closed forms use explicit __dmul_rn/__dadd_rn: given the same x = exp(-beta d) the
// matrices are bit-identical to the reference's.  x may come from the host libm
// (bit-exact mode) or be computed here.  GTR replaces scipy.linalg.expm(Q d)
// (mcmc_gamma.pyx:481) by I + U diag(expm1(lambda d)) U^-1 from one symmetric
// eigendecomposition per (pi, rates) done on the host.
#pragma once
#include "cb_types.cuh"

namespace cb {

__global__ void __launch_bounds__(256) pmat_build_kernel(int model, int S, const double* __restrict__ pi,
                                                         double beta, const double* __restrict__ gtr,
                                                         int count, const int32_t* __restrict__ slots,
                                                         const double* __restrict__ d, const double* __restrict__ x_in,
                                                         double* __restrict__ pmats) {
  extern __shared__ double ek[];  // GTR: expm1(lambda_k d)
  const int m = blockIdx.x;
  if (m >= count) return;
  double* out = pmats + (int64_t)slots[m] * S * S;
  const double dd = d[m];
  if (model == CB_MODEL_GTR_EIG) {
    const double* lam = gtr;
    const double* U = gtr + S;
    const double* Ui = gtr + S + (int64_t)S * S;
    // P = I + U diag(expm1(lambda d)) U^-1: on short branches the small off-diagonal entries are
    // sums of O(d) terms instead of differences of O(1) terms (no cancellation against I).
    for (int q = threadIdx.x; q < S; q += blockDim.x) ek[q] = expm1(lam[q] * dd);
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
      const int i = idx / S, j = idx - i * S;
      double acc = 0.0;
      for (int q = 0; q < S; ++q) acc = fma(U[i * S + q] * ek[q], Ui[q * S + j], acc);
      out[idx] = (i == j) ? 1.0 + acc : acc;
    }
    return;
  }
  const double x = x_in ? x_in[m] : exp(__dmul_rn(-beta, dd));
  if (model == CB_MODEL_JC) {
    const double y = __ddiv_rn(__dsub_rn(1.0, x), (double)S);
    const double diag = __dadd_rn(x, y);
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
      const int i = idx / S, j = idx - i * S;
      out[idx] = (i == j) ? diag : y;
    }
  } else if (model == CB_MODEL_F81) {
    const double y = __dsub_rn(1.0, x);
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
      const int i = idx / S, j = idx - i * S;
      const double v = __dmul_rn(pi[j], y);
      out[idx] = (i == j) ? __dadd_rn(v, x) : v;
    }
  } else {  // CB_MODEL_F81_BINARY
    if (threadIdx.x == 0) {
      const double y = __dsub_rn(1.0, x);
      out[0] = __dadd_rn(pi[0], __dmul_rn(pi[1], x));
      out[1] = __dmul_rn(pi[1], y);
      out[2] = __dmul_rn(pi[0], y);
      out[3] = __dadd_rn(pi[1], __dmul_rn(pi[0], x));
    }
  }
}

}  // namespace cb
