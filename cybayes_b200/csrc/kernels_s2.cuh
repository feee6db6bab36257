// Two-state (binary alignment) pruning kernel for sm_100a -- HBM-bandwidth bound.
//
// Restates ML_gamma.pyx:24-38 for S = 2: per node, per rate category, per site
//     v_k[i]  = P_k[i][0] * L_k[0] + P_k[i][1] * L_k[1]        (k = the two child edges)
//     L[i]    = v_a[i] * v_b[i]
// plus what the reference lacks: an exact power-of-two per-site rescale (one int32 exponent
// per site and node, shared by all categories) so deep trees do not underflow.
//
// Mapping: one thread owns two adjacent sites for ALL categories, so every global access is
// a fully coalesced 16-byte (double2) access and the cross-category max needs no
// communication.  A block walks a *range* of ops for its site tile: in level mode a range
// is one node (grid.y = nodes of the level); in path mode the range is a whole dirty path
// (ML_gamma.pyx:99-114) and the on-path partial never leaves registers.  The P matrices of
// the range are staged once per block in shared memory and read back as broadcasts.
// A root op does not store anything: it folds pi, the category mean, log, the site-pattern
// weight and the accumulated exponent into a per-block sum; the last block to finish adds the
// block sums in a fixed order (bit-reproducible, no floating-point atomics).
#pragma once
#include "cb_types.cuh"

namespace cb {

__device__ __forceinline__ double2 ld_cg2(const double* p) {
  return __ldcg(reinterpret_cast<const double2*>(p));
}
__device__ __forceinline__ void st_cg2(double* p, double2 v) {
  __stcg(reinterpret_cast<double2*>(p), v);
}

// exponent e with 2^e <= m < 2^(e+1) for normal m > 0; 0 otherwise
__device__ __forceinline__ int exponent_of(double m) {
  int hi = __double2hiint(m);
  int be = (hi >> 20) & 0x7ff;
  return (be == 0 || be == 0x7ff) ? 0 : be - 1023;
}
__device__ __forceinline__ double pow2_neg(int e) {  // 2^-e, |e| <= 1022
  return __hiloint2double((1023 - e) << 20, 0);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block -> grid reduction of one double per thread.
// red must hold 32 doubles; every thread of the block must call this.
__device__ __forceinline__ void block_reduce_to_result(double v, const LaunchConst& k, int out_index,
                                                       double* red, int* last_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = (lane < nwarp) ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) {
      __stcg(&k.block_sums[(int64_t)out_index * k.max_blocks + blockIdx.x], s);
      __threadfence();
      unsigned t = atomicAdd(&k.tickets[out_index], 1u);
      *last_flag = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (*last_flag && warp == 0) {
    __threadfence();
    const double* bs = k.block_sums + (int64_t)out_index * k.max_blocks;
    double s = 0.0;
    for (unsigned i = lane; i < gridDim.x; i += 32) s += __ldcg(bs + i);
    s = warp_sum(s);
    if (lane == 0) {
      k.results[out_index] = s;
      k.tickets[out_index] = 0u;
    }
  }
}

constexpr int S2_STAGE_OPS = 64;  // P matrices of this many ops are staged in shared memory at a time

template <int C>
__global__ void __launch_bounds__(256) prune_s2_kernel(const LaunchConst k) {
  extern __shared__ double2 p_stage[];  // [ops in range][2 children][C][2 rows] as (Pi0, Pi1)
  __shared__ double red[32];
  __shared__ int last_flag;

  const RangeDesc rg = k.ranges[blockIdx.y];
  const int nops = rg.end - rg.begin;
  const int64_t P = k.n_sites;
  const int64_t site = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  const bool active = site < P;
  double lnl = 0.0;
  double2 cur[C][2];  // carried partial: [category][state] x two sites
  int cur_e0 = 0, cur_e1 = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) cur[c][0] = cur[c][1] = make_double2(0.0, 0.0);

#pragma unroll 1
  for (int o0 = 0; o0 < nops; o0 += S2_STAGE_OPS) {
    const int o1 = min(nops, o0 + S2_STAGE_OPS);
    if (o0 > 0) __syncthreads();  // everybody is done with the previous chunk's matrices
    {
      double* ps = reinterpret_cast<double*>(p_stage);
      for (int idx = threadIdx.x; idx < (o1 - o0) * 8 * C; idx += blockDim.x) {
        const int e = idx & 3, c = (idx >> 2) % C, ch = (idx / (4 * C)) & 1, o = idx / (8 * C);
        ps[idx] = __ldg(k.pmats + (int64_t)k.ops[rg.begin + o0 + o].pslot[ch][c] * 4 + e);
      }
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll 1
    for (int o = o0; o < o1; ++o) {
      const OpDesc* __restrict__ op = k.ops + rg.begin + o;
      double2 out[C][2];
      int e0 = 0, e1 = 0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = op->kind[ch];
        double2 L[C][2];
        if (kind == SRC_CARRIED) {
#pragma unroll
          for (int c = 0; c < C; ++c) { L[c][0] = cur[c][0]; L[c][1] = cur[c][1]; }
          e0 += cur_e0; e1 += cur_e1;
        } else if (kind == SRC_BUFFER) {
          const double* src = static_cast<const double*>(op->src[ch]) + site;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            L[c][0] = ld_cg2(src + (int64_t)(2 * c) * P);
            L[c][1] = ld_cg2(src + (int64_t)(2 * c + 1) * P);
          }
          const int2 se = __ldcg(reinterpret_cast<const int2*>(op->src_scale[ch] + site));
          e0 += se.x; e1 += se.y;
        } else {  // tip: state codes -> 0/1 indicator columns (utils.pyx:99-111)
          unsigned c0, c1;
          if (k.code_bytes == 1) {
            const uchar2 cc = __ldg(reinterpret_cast<const uchar2*>(static_cast<const uint8_t*>(op->src[ch]) + site));
            c0 = cc.x; c1 = cc.y;
          } else {
            const ushort2 cc = __ldg(reinterpret_cast<const ushort2*>(static_cast<const uint16_t*>(op->src[ch]) + site));
            c0 = cc.x; c1 = cc.y;
          }
          double2 t0, t1;  // t0 = indicator of state 0 for (site, site+1); t1 = state 1
          t0.x = (c0 < 2) ? (c0 == 0 ? 1.0 : 0.0) : __ldg(k.amb + (c0 - 2) * 2);
          t1.x = (c0 < 2) ? (c0 == 1 ? 1.0 : 0.0) : __ldg(k.amb + (c0 - 2) * 2 + 1);
          t0.y = (c1 < 2) ? (c1 == 0 ? 1.0 : 0.0) : __ldg(k.amb + (c1 - 2) * 2);
          t1.y = (c1 < 2) ? (c1 == 1 ? 1.0 : 0.0) : __ldg(k.amb + (c1 - 2) * 2 + 1);
#pragma unroll
          for (int c = 0; c < C; ++c) { L[c][0] = t0; L[c][1] = t1; }
        }
        const double2* pm = p_stage + ((o - o0) * 2 + ch) * C * 2;
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double2 pr = pm[c * 2 + i];  // (P[i][0], P[i][1]) broadcast
            double2 v;
            v.x = fma(pr.y, L[c][1].x, pr.x * L[c][0].x);
            v.y = fma(pr.y, L[c][1].y, pr.x * L[c][0].y);
            if (ch == 0) {
              out[c][i] = v;
            } else {
              out[c][i].x *= v.x;
              out[c][i].y *= v.y;
            }
          }
        }
      }
      if (!op->is_root) {
        double m0 = out[0][0].x, m1 = out[0][0].y;
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            m0 = fmax(m0, out[c][i].x);
            m1 = fmax(m1, out[c][i].y);
          }
        }
        const int x0 = exponent_of(m0), x1 = exponent_of(m1);
        const double f0 = pow2_neg(x0), f1 = pow2_neg(x1);
#pragma unroll
        for (int c = 0; c < C; ++c) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            cur[c][i].x = out[c][i].x * f0;
            cur[c][i].y = out[c][i].y * f1;
          }
        }
        cur_e0 = e0 + x0;
        cur_e1 = e1 + x1;
        if (op->dst != nullptr) {
          double* dst = op->dst + site;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            st_cg2(dst + (int64_t)(2 * c) * P, cur[c][0]);
            st_cg2(dst + (int64_t)(2 * c + 1) * P, cur[c][1]);
          }
          __stcg(reinterpret_cast<int2*>(op->dst_scale + site), make_int2(cur_e0, cur_e1));
        }
      } else {
        // ll_p = sum_c (pi . L_c) / n_cats ; lnL += w_p * log(ll_p)      ML_gamma.pyx:38,40
        const double pi0 = __ldg(k.pi), pi1 = __ldg(k.pi + 1);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          s0 += fma(pi1, out[c][1].x, pi0 * out[c][0].x) / k.cats;
          s1 += fma(pi1, out[c][1].y, pi0 * out[c][0].y) / k.cats;
        }
        if (op->dst != nullptr) {  // optional store of the (unscaled-at-this-node) root partial
          double* dst = op->dst + site;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            st_cg2(dst + (int64_t)(2 * c) * P, out[c][0]);
            st_cg2(dst + (int64_t)(2 * c + 1) * P, out[c][1]);
          }
          __stcg(reinterpret_cast<int2*>(op->dst_scale + site), make_int2(e0, e1));
        }
        const double2 w = ld_cg2(k.weights + site);
        const double ln2 = 0.693147180559945309417232121458;
        if (w.x != 0.0) lnl += w.x * (log(s0) + (double)e0 * ln2);
        if (w.y != 0.0) lnl += w.y * (log(s1) + (double)e1 * ln2);
      }
    }
  }
  if (rg.out_index >= 0) block_reduce_to_result(lnl, k, rg.out_index, red, &last_flag);
}

}  // namespace cb
