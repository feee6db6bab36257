// Multistate (9 <= S <= 64) pruning kernel on the FP64 tensor path (DMMA), sm_100a.
//
// Same recursion as kernels_general.cuh / ML_gamma.pyx:24-38; here the per-branch contraction
//     V[i][t] = sum_j P[i][j] * L[j][t]            (S x S) @ (S x 64-site tile)
// is dense enough to be a small GEMM, so it runs on `mma.sync.aligned.m8n8k4.f64` (tcgen05 has no
// fp64 kind; on sm_100a FP64 tensor work stays on mma.sync).  P is the A operand (row-major in
// shared memory, reused by the whole 64-site tile), the child partial tile is the B operand.
//
// One block = 64-site tile x op range x ONE rate category; 8 warps arranged 2 (row halves of the
// output) x 4 (16-site column groups); each warp keeps up to 4 x 2 accumulator tiles (8x8) per
// child in registers.  Shared memory: one P matrix (padded stride == 4 mod 16 doubles: the
// A-fragment loads are conflict free) and one [S8][72] tile (stride 72 == 8 mod 16: B-fragment
// loads are conflict free), tip codes, row sums, column maxima.
// Tip children skip the GEMM: one-hot = column gather of P, all-ones = row sums, other sets dense.
// A dirty path (or the depth-first whole-tree walk) keeps the on-path partial in the shared tile.
#pragma once
#include "cb_types.cuh"
#include "kernels_s2.cuh"

namespace cb {

constexpr int DM_T = 64;          // sites per tile
constexpr int DM_THREADS = 256;   // 8 warps
constexpr int DM_LS = DM_T + 8;   // tile row stride (doubles)
constexpr int DM_MAX_MT = 4;      // m-tiles (8 rows) per warp: S <= 64

__host__ __device__ inline int dm_s8(int S) { return (S + 7) / 8 * 8; }
__host__ __device__ inline int dm_ps(int S) {  // P row stride: >= S rounded to 4, == 4 (mod 16)
  int s4 = (S + 3) / 4 * 4;
  int ps = s4;
  while (ps % 16 != 4) ps += 4;
  return ps;
}
__host__ __device__ inline size_t dm_smem_bytes(int S) {
  const int S8 = dm_s8(S);
  return sizeof(double) * ((size_t)2 * S8 * dm_ps(S) + (size_t)S8 * DM_LS + S8 + 4 * DM_T) + sizeof(int) * (size_t)(2 * DM_T);
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// global (row-major S x S) -> shared (row stride PS) without touching registers
__device__ __forceinline__ void stage_p_async(double* Pm, const double* pm, int S, int PS, int warp, int lane) {
  if ((S & 1) == 0) {
    const int h = S >> 1;
    for (int i = warp; i < S; i += DM_THREADS / 32)
      for (int jj = lane; jj < h; jj += 32) cp_async16(Pm + i * PS + 2 * jj, pm + i * S + 2 * jj);
  } else {
    for (int i = warp; i < S; i += DM_THREADS / 32)
      for (int j = lane; j < S; j += 32) cp_async8(Pm + i * PS + j, pm + i * S + j);
  }
}

// Shared memory holds the P matrices of both children (the second one streams in with cp.async while
// the tensor cores work on the first) and ONE [S8][72] tile: the two children are contracted one after
// the other into register accumulators (the carried child first, since it already is in the tile),
// and the product is written back into the same tile, which then is the carried partial of the next
// op.  ~107 KB at S = 64: two blocks share an SM, so one block's staging overlaps the other's tensor work.
// SCT = compile-time state count (0 = take it from the launch constants): with S known the staging,
// contraction and finalise loops unroll and their index arithmetic folds into immediates.
template <int SCT>
__global__ void __launch_bounds__(DM_THREADS, 2) prune_dmma_kernel(const LaunchConst k) {
  extern __shared__ __align__(16) double dsm[];
  const int S = SCT ? SCT : k.n_states;
  const int S8 = dm_s8(S), S4 = (S + 3) / 4 * 4, PS = dm_ps(S);
  double* Pm0 = dsm;                             // [2][S8][PS]
  double* tile = Pm0 + 2 * S8 * PS;              // [S8][DM_LS]
  double* rowsum = tile + S8 * DM_LS;            // [S8]
  double* colmax = rowsum + S8;                  // [4][DM_T]
  int* codes = reinterpret_cast<int*>(colmax + 4 * DM_T);  // [DM_T]
  int* cur_e = codes + DM_T;                     // [DM_T] exponent of the carried tile

  const RangeDesc rg = k.ranges[blockIdx.y];
  const int c = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t4 = lane & 3;        // mma groupID / threadID_in_group
  const int wm = warp >> 2, wn = warp & 3;       // warp grid 2 x 4
  const int n0 = wn * 16;
  const int n_mt = S8 / 8;
  const int64_t P = k.n_sites;
  const int64_t site0 = (int64_t)blockIdx.x * DM_T;
  const int fs = tid & (DM_T - 1), fg = tid >> 6;  // finalise mapping: site, row group (4 groups)

  // zero once: padded rows / columns must stay 0 (never NaN) for the whole kernel
  for (int i = tid; i < 2 * S8 * PS + S8 * DM_LS + S8 + 4 * DM_T; i += DM_THREADS) dsm[i] = 0.0;
  __syncthreads();

#pragma unroll 1
  for (int o = rg.begin; o < rg.end; ++o) {
    const OpDesc* __restrict__ op = k.ops + o;
    const int first = (op->kind[1] == SRC_CARRIED) ? 1 : 0;  // the carried child already sits in the tile
    double prod[DM_MAX_MT][2][2];
    int e_sum = 0;
    // both P matrices start streaming now: group 0 = first child (+ its tile), group 1 = second child's P
    stage_p_async(Pm0, k.pmats + (int64_t)op->pslot[first][c] * S * S, S, PS, warp, lane);
    {
      const int kind = op->kind[first];
      if (kind == SRC_BUFFER) {
        const double* src = static_cast<const double*>(op->src[first]) + (int64_t)c * S * P + site0;
        for (int j = warp; j < S; j += 8) cp_async16(tile + j * DM_LS + lane * 2, src + (int64_t)j * P + lane * 2);
      }
    }
    cp_async_commit();
    stage_p_async(Pm0 + S8 * PS, k.pmats + (int64_t)op->pslot[1 - first][c] * S * S, S, PS, warp, lane);
    cp_async_commit();
#pragma unroll 1
    for (int step = 0; step < 2; ++step) {
      const int ch = step ? 1 - first : first;
      const int kind = op->kind[ch];
      const double* Pm = Pm0 + step * S8 * PS;
      if (step == 1 && kind == SRC_BUFFER) {  // the tile is free now (barrier at the end of step 0)
        const double* src = static_cast<const double*>(op->src[ch]) + (int64_t)c * S * P + site0;
        for (int j = warp; j < S; j += 8) cp_async16(tile + j * DM_LS + lane * 2, src + (int64_t)j * P + lane * 2);
        cp_async_commit();
      }
      if (kind == SRC_BUFFER) {
        e_sum += __ldcg(op->src_scale[ch] + (int64_t)c * P + site0 + fs);
      } else if (kind == SRC_TIP) {
        if (tid < DM_T)
          codes[tid] = (k.code_bytes == 1) ? (int)__ldg(static_cast<const uint8_t*>(op->src[ch]) + site0 + tid)
                                           : (int)__ldg(static_cast<const uint16_t*>(op->src[ch]) + site0 + tid);
      } else {
        e_sum += cur_e[fs];
      }
      if (step == 0) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncthreads();
      double acc[DM_MAX_MT][2][2];
#pragma unroll
      for (int m = 0; m < DM_MAX_MT; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
      if (kind != SRC_TIP) {
        const double* bp = tile + t4 * DM_LS + n0 + g;
        const double* ap = Pm + (wm * 8 + g) * PS + t4;
        const int a_step = 16 * PS;  // two m-tiles down (the other warp row owns the one in between)
#pragma unroll 2
        for (int k0 = 0; k0 < S4; k0 += 4) {
          const double b0 = bp[0], b1 = bp[8];
#pragma unroll
          for (int m = 0; m < DM_MAX_MT; ++m) {
            if (wm + 2 * m < n_mt) {
              const double a = ap[m * a_step];
              dmma_8x8x4(acc[m][0][0], acc[m][0][1], a, b0);
              dmma_8x8x4(acc[m][1][0], acc[m][1][1], a, b1);
            }
          }
          bp += 4 * DM_LS;
          ap += 4;
        }
      } else {
        // row sums are only needed for '?' / '-' cells (code == S): skip them when the tile has none;
        // otherwise 4 threads per row add every fourth entry each and combine in a fixed order
        const int any_missing = __syncthreads_or(tid < DM_T && codes[tid] == S);
        if (any_missing) {
          const int i = tid >> 2, q = tid & 3;
          double s = 0.0;
          if (i < S) {
            const double* row = Pm + i * PS;
            for (int j = q; j < S; j += 4) s += row[j];
          }
          const double s1 = __shfl_down_sync(0xffffffffu, s, 1);
          const double s2 = __shfl_down_sync(0xffffffffu, s, 2);
          const double s3 = __shfl_down_sync(0xffffffffu, s, 3);
          if (q == 0 && i < S) rowsum[i] = ((s + s1) + s2) + s3;
          __syncthreads();
        }
#pragma unroll
        for (int m = 0; m < DM_MAX_MT; ++m) {
          const int mt = wm + 2 * m;
          if (mt < n_mt) {
            const int i = mt * 8 + g;
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int cd = codes[n0 + n * 8 + 2 * t4 + e];
                double v;
                if (cd < S) {
                  v = Pm[i * PS + cd];
                } else if (cd == S) {
                  v = rowsum[i];
                } else {
                  const double* am = k.amb + (int64_t)(cd - S) * S;
                  v = 0.0;
                  for (int j = 0; j < S; ++j) v = fma(Pm[i * PS + j], __ldg(am + j), v);
                }
                acc[m][n][e] = v;
              }
          }
        }
      }
#pragma unroll
      for (int m = 0; m < DM_MAX_MT; ++m)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (step == 0) prod[m][n][e] = acc[m][n][e]; else prod[m][n][e] *= acc[m][n][e];
          }
      __syncthreads();  // every warp is done with the tile / codes of this child
    }
#pragma unroll
    for (int m = 0; m < DM_MAX_MT; ++m) {
      const int mt = wm + 2 * m;
      if (mt < n_mt) {
#pragma unroll
        for (int n = 0; n < 2; ++n)
          *reinterpret_cast<double2*>(tile + (mt * 8 + g) * DM_LS + n0 + n * 8 + 2 * t4) =
              make_double2(prod[m][n][0], prod[m][n][1]);
      }
    }
    __syncthreads();
    // ---- finalise: rescale, store, keep as the carried partial ---------------------------------
    if (!op->is_root) {
      double m = 0.0;
      for (int j = fg; j < S; j += 4) m = fmax(m, tile[j * DM_LS + fs]);
      colmax[fg * DM_T + fs] = m;
      __syncthreads();
      m = fmax(fmax(colmax[fs], colmax[DM_T + fs]), fmax(colmax[2 * DM_T + fs], colmax[3 * DM_T + fs]));
      const int x = exponent_of(m);
      const double f = pow2_neg(x);
      double* dst = op->dst ? op->dst + (int64_t)c * S * P + site0 + fs : nullptr;
      for (int j = fg; j < S; j += 4) {
        const double v = tile[j * DM_LS + fs] * f;
        tile[j * DM_LS + fs] = v;
        if (dst) __stcg(dst + (int64_t)j * P, v);
      }
      if (fg == 0) {  // (all readers of cur_e[] passed the barriers above)
        cur_e[fs] = e_sum + x;
        if (dst) __stcg(op->dst_scale + (int64_t)c * P + site0 + fs, e_sum + x);
      }
    } else {
      if (op->dst != nullptr) {
        double* dst = op->dst + (int64_t)c * S * P + site0 + fs;
        for (int j = fg; j < S; j += 4) __stcg(dst + (int64_t)j * P, tile[j * DM_LS + fs]);
        if (fg == 0) __stcg(op->dst_scale + (int64_t)c * P + site0 + fs, e_sum);
      }
      if (fg == 0) {
        double dot = 0.0;
        for (int j = 0; j < S; ++j) dot = fma(__ldg(k.pi + j), tile[j * DM_LS + fs], dot);
        const int64_t at = ((int64_t)rg.out_index * k.n_cats + c) * P + site0 + fs;
        __stcg(k.root_dot + at, dot);
        __stcg(k.root_exp + at, e_sum);
      }
    }
    __syncthreads();  // tile, colmax and cur_e are settled before the next op streams into shared memory
  }
}

}  // namespace cb
