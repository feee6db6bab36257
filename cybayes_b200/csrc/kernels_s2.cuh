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

constexpr int S2_STAGE_OPS = 64;  // ops whose descriptors + P matrices are staged in shared memory at a time

// Vector of V doubles / ints per thread (V sites), with 8*V-byte global accesses.
template <int V> struct VecD;
template <> struct VecD<1> {
  double v[1];
  __device__ __forceinline__ void load(const double* p) { v[0] = __ldcg(p); }
  __device__ __forceinline__ void load_ca(const double* p) { v[0] = __ldca(p); }
  __device__ __forceinline__ void store(double* p) const { __stcg(p, v[0]); }
};
template <> struct VecD<2> {
  double v[2];
  __device__ __forceinline__ void load(const double* p) { const double2 t = ld_cg2(p); v[0] = t.x; v[1] = t.y; }
  __device__ __forceinline__ void load_ca(const double* p) {
    const double2 t = __ldca(reinterpret_cast<const double2*>(p)); v[0] = t.x; v[1] = t.y;
  }
  __device__ __forceinline__ void store(double* p) const { st_cg2(p, make_double2(v[0], v[1])); }
};
template <int V> __device__ __forceinline__ void load_ints(const int32_t* p, int (&e)[V]);
template <> __device__ __forceinline__ void load_ints<1>(const int32_t* p, int (&e)[1]) { e[0] = __ldcg(p); }
template <> __device__ __forceinline__ void load_ints<2>(const int32_t* p, int (&e)[2]) {
  const int2 t = __ldcg(reinterpret_cast<const int2*>(p)); e[0] = t.x; e[1] = t.y;
}
template <int V> __device__ __forceinline__ void load_ints_ca(const int32_t* p, int (&e)[V]);
template <> __device__ __forceinline__ void load_ints_ca<1>(const int32_t* p, int (&e)[1]) { e[0] = __ldca(p); }
template <> __device__ __forceinline__ void load_ints_ca<2>(const int32_t* p, int (&e)[2]) {
  const int2 t = __ldca(reinterpret_cast<const int2*>(p)); e[0] = t.x; e[1] = t.y;
}
template <int V> __device__ __forceinline__ void store_ints(int32_t* p, const int (&e)[V]);
template <> __device__ __forceinline__ void store_ints<1>(int32_t* p, const int (&e)[1]) { __stcg(p, e[0]); }
template <> __device__ __forceinline__ void store_ints<2>(int32_t* p, const int (&e)[2]) {
  __stcg(reinterpret_cast<int2*>(p), make_int2(e[0], e[1]));
}
template <int V> __device__ __forceinline__ void load_codes(const void* row, int code_bytes, int64_t site, unsigned (&c)[V]);
template <> __device__ __forceinline__ void load_codes<1>(const void* row, int code_bytes, int64_t site, unsigned (&c)[1]) {
  c[0] = (code_bytes == 1) ? (unsigned)__ldg(static_cast<const uint8_t*>(row) + site)
                           : (unsigned)__ldg(static_cast<const uint16_t*>(row) + site);
}
template <> __device__ __forceinline__ void load_codes<2>(const void* row, int code_bytes, int64_t site, unsigned (&c)[2]) {
  if (code_bytes == 1) {
    const uchar2 t = __ldg(reinterpret_cast<const uchar2*>(static_cast<const uint8_t*>(row) + site));
    c[0] = t.x; c[1] = t.y;
  } else {
    const ushort2 t = __ldg(reinterpret_cast<const ushort2*>(static_cast<const uint16_t*>(row) + site));
    c[0] = t.x; c[1] = t.y;
  }
}

// per-op record staged in shared memory (64 B)
struct S2Stage {
  double* dst;
  int32_t* dst_scale;
  const void* src[2];
  const int32_t* src_scale[2];
  int32_t kind[2];
  int32_t is_root, pad_;
};
static_assert(sizeof(S2Stage) == 64, "S2Stage is 64 bytes");

template <int V>
__device__ __forceinline__ void fetch_codes(const S2Stage& op, int code_bytes, int64_t site, unsigned (&code)[2][V]) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
    if (op.kind[ch] == SRC_TIP) load_codes<V>(op.src[ch], code_bytes, site, code[ch]);
}

__host__ __device__ inline size_t s2_smem_bytes(int ops, int C) {
  const int n = ops < S2_STAGE_OPS ? ops : S2_STAGE_OPS;
  return (size_t)n * (sizeof(S2Stage) + (size_t)16 * C * sizeof(double));
}

template <int C, int V, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) prune_s2_kernel(const LaunchConst k) {
  extern __shared__ __align__(16) unsigned char s2_smem[];
  __shared__ double red[32];
  __shared__ int last_flag;

  const RangeDesc rg = k.ranges[blockIdx.y];
  const int nops = rg.end - rg.begin;
  const int n_stage = min(nops, S2_STAGE_OPS);
  S2Stage* st = reinterpret_cast<S2Stage*>(s2_smem);
  // per (op, child): 8*C doubles.  Internal child: [C][row] pairs (Pi0, Pi1) in the first half.
  // Tip child: [C][row][4] = (Pi0, Pi1, Pi0 + Pi1, -): the edge's contribution looked up by state code.
  double* p_stage = reinterpret_cast<double*>(s2_smem + (size_t)n_stage * sizeof(S2Stage));

  const int64_t P = k.n_sites;
  const int64_t tile0 = (int64_t)blockIdx.x * (THREADS * V);
  const int64_t site = tile0 + (int64_t)threadIdx.x * V;
  const bool active = site < P;
  double lnl = 0.0;
  VecD<V> cur[C][2];  // carried partial: [category][state] x V sites
  int cur_e[V];
#pragma unroll
  for (int v = 0; v < V; ++v) cur_e[v] = 0;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int v = 0; v < V; ++v) cur[c][0].v[v] = cur[c][1].v[v] = 0.0;

#pragma unroll 1
  for (int o0 = 0; o0 < nops; o0 += S2_STAGE_OPS) {
    const int o1 = min(nops, o0 + S2_STAGE_OPS);
    if (o0 > 0) __syncthreads();  // everybody is done with the previous chunk
    // stage descriptors (one thread per op) ...
    for (int o = threadIdx.x; o < o1 - o0; o += THREADS) {
      const OpDesc* __restrict__ op = k.ops + rg.begin + o0 + o;
      S2Stage r;
      r.dst = op->dst; r.dst_scale = op->dst_scale;
      r.src[0] = op->src[0]; r.src[1] = op->src[1];
      r.src_scale[0] = op->src_scale[0]; r.src_scale[1] = op->src_scale[1];
      r.kind[0] = op->kind[0]; r.kind[1] = op->kind[1];
      r.is_root = op->is_root; r.pad_ = 0;
      st[o] = r;
    }
    // ... their P matrices (tip children get a 3-entry lookup per row: state 0, state 1, missing) ...
    for (int idx = threadIdx.x; idx < (o1 - o0) * 4 * C; idx += THREADS) {
      const int i = idx & 1, c = (idx >> 1) % C, ch = (idx / (2 * C)) & 1, o = idx / (4 * C);
      const OpDesc* __restrict__ op = k.ops + rg.begin + o0 + o;
      const double2 pr = __ldg(reinterpret_cast<const double2*>(k.pmats + (int64_t)op->pslot[ch][c] * 4) + i);
      double* base = p_stage + (size_t)(o * 2 + ch) * 8 * C;
      if (op->kind[ch] == SRC_TIP) {
        double* t = base + (c * 2 + i) * 4;
        t[0] = pr.x;
        t[1] = pr.y;
        t[2] = fma(pr.y, 1.0, pr.x * 1.0);  // all-ones column: P[i][0] + P[i][1], as the FMA chain gives it
        t[3] = 0.0;
      } else {
        reinterpret_cast<double2*>(base)[c * 2 + i] = pr;
      }
    }
    // ... and pull this tile's tip codes of the NEXT chunk towards L2 while this chunk computes
    {
      const int nxt = min(nops, o1 + S2_STAGE_OPS) - o1;
      const int tile_bytes = THREADS * V * k.code_bytes;
      const int lines = (tile_bytes + 127) / 128;
      for (int idx = threadIdx.x; idx < nxt * 2 * lines; idx += THREADS) {
        const int ln = idx % lines, oc = idx / lines;
        const OpDesc* __restrict__ op = k.ops + rg.begin + o1 + (oc >> 1);
        if (op->kind[oc & 1] == SRC_TIP) {
          const char* a = static_cast<const char*>(op->src[oc & 1]) + tile0 * k.code_bytes + ln * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
      }
    }
    __syncthreads();
    if (!active) continue;
    // tip codes are fetched one op ahead (register prefetch): by the time an op starts, its codes
    // have had a whole op of compute to arrive from L2
    unsigned code_next[2][V];
    fetch_codes<V>(st[0], k.code_bytes, site, code_next);
#pragma unroll 1
    for (int o = o0; o < o1; ++o) {
      const S2Stage& op = st[o - o0];
      unsigned code[2][V];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int v = 0; v < V; ++v) code[ch][v] = code_next[ch][v];
      if (o + 1 < o1) fetch_codes<V>(st[o + 1 - o0], k.code_bytes, site, code_next);

      VecD<V> out[C][2];
      int e_in[V];
#pragma unroll
      for (int v = 0; v < V; ++v) e_in[v] = 0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = op.kind[ch];
        const double* pbase = p_stage + (size_t)((o - o0) * 2 + ch) * 8 * C;
        const double2* pm = reinterpret_cast<const double2*>(pbase);  // (P[i][0], P[i][1]) broadcasts
        // v[i] = P[i][0] L[0] + P[i][1] L[1]; first child initialises `out`, second multiplies into it
#define CB_S2_APPLY(L0, L1)                                                   \
  _Pragma("unroll") for (int c = 0; c < C; ++c) {                            \
    _Pragma("unroll") for (int i = 0; i < 2; ++i) {                          \
      const double2 pr = pm[c * 2 + i];                                       \
      _Pragma("unroll") for (int v = 0; v < V; ++v) {                        \
        const double x = fma(pr.y, (L1), pr.x * (L0));                        \
        if (ch == 0) out[c][i].v[v] = x; else out[c][i].v[v] *= x;            \
      }                                                                       \
    }                                                                         \
  }
        if (kind == SRC_CARRIED) {
          CB_S2_APPLY(cur[c][0].v[v], cur[c][1].v[v])
#pragma unroll
          for (int v = 0; v < V; ++v) e_in[v] += cur_e[v];
        } else if (kind == SRC_BUFFER) {
          const double* src = static_cast<const double*>(op.src[ch]) + site;
          VecD<V> L[C][2];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            L[c][0].load(src + (int64_t)(2 * c) * P);
            L[c][1].load(src + (int64_t)(2 * c + 1) * P);
          }
          int se[V];
          load_ints<V>(op.src_scale[ch] + site, se);
          CB_S2_APPLY(L[c][0].v[v], L[c][1].v[v])
#pragma unroll
          for (int v = 0; v < V; ++v) e_in[v] += se[v];
        } else {  // tip: state code 0 / 1 / 2 ('?', '-', '0/1') selects the staged contribution of this edge
          // (= the FMA chain over the 0/1 indicator column of utils.pyx:99-111, bit for bit)
#pragma unroll
          for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const double x = pbase[(c * 2 + i) * 4 + min(code[ch][v], 2u)];
                if (ch == 0) out[c][i].v[v] = x; else out[c][i].v[v] *= x;
              }
            }
          }
        }
#undef CB_S2_APPLY
      }
      if (!op.is_root) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          // all entries are >= 0, so the max of the high words carries the exponent of the max
          int mh = __double2hiint(out[0][0].v[v]);
#pragma unroll
          for (int c = 0; c < C; ++c) mh = max(mh, max(__double2hiint(out[c][0].v[v]), __double2hiint(out[c][1].v[v])));
          const int be = (mh >> 20) & 0x7ff;
          const int x = (be == 0 || be == 0x7ff) ? 0 : be - 1023;
          const double f = pow2_neg(x);
#pragma unroll
          for (int c = 0; c < C; ++c) {
            cur[c][0].v[v] = out[c][0].v[v] * f;
            cur[c][1].v[v] = out[c][1].v[v] * f;
          }
          cur_e[v] = e_in[v] + x;
        }
        if (op.dst != nullptr) {
          double* dst = op.dst + site;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            cur[c][0].store(dst + (int64_t)(2 * c) * P);
            cur[c][1].store(dst + (int64_t)(2 * c + 1) * P);
          }
          store_ints<V>(op.dst_scale + site, cur_e);
        }
      } else {
        // ll_p = sum_c (pi . L_c) / n_cats ; lnL += w_p * log(ll_p)      ML_gamma.pyx:38,40
        const double pi0 = __ldg(k.pi), pi1 = __ldg(k.pi + 1);
        if (op.dst != nullptr) {  // optional store of the (unscaled-at-this-node) root partial
          double* dst = op.dst + site;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            out[c][0].store(dst + (int64_t)(2 * c) * P);
            out[c][1].store(dst + (int64_t)(2 * c + 1) * P);
          }
          store_ints<V>(op.dst_scale + site, e_in);
        }
        VecD<V> w;
        w.load(k.weights + site);
        const double ln2 = 0.693147180559945309417232121458;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < C; ++c) s += fma(pi1, out[c][1].v[v], pi0 * out[c][0].v[v]) / k.cats;
          if (w.v[v] != 0.0) lnl += w.v[v] * (log(s) + (double)e_in[v] * ln2);
        }
      }
    }
  }
  if (rg.out_index >= 0) block_reduce_to_result(lnl, k, rg.out_index, red, &last_flag);
}

}  // namespace cb
