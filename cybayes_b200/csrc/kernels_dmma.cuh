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
// child in registers.  Shared memory: both P matrices (padded stride == 4 mod 16 doubles: the
// A-fragment loads are conflict free), three [S8][72] tiles (two children, one output/carried;
// stride 72 == 8 mod 16: B-fragment loads are conflict free), tip codes, row sums, column maxima.
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
  return sizeof(double) * ((size_t)2 * S8 * dm_ps(S) + (size_t)3 * S8 * DM_LS + 2 * S8 + 4 * DM_T) +
         sizeof(int) * (size_t)(2 * DM_T + DM_T);
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DM_THREADS) prune_dmma_kernel(const LaunchConst k) {
  extern __shared__ __align__(16) double dsm[];
  const int S = k.n_states, S8 = dm_s8(S), S4 = (S + 3) / 4 * 4, PS = dm_ps(S);
  double* Pm = dsm;                              // [2][S8][PS]
  double* tiles = Pm + 2 * S8 * PS;              // [3][S8][DM_LS]
  double* rowsum = tiles + 3 * S8 * DM_LS;       // [2][S8]
  double* colmax = rowsum + 2 * S8;              // [4][DM_T]
  int* codes = reinterpret_cast<int*>(colmax + 4 * DM_T);  // [2][DM_T]
  int* cur_e = codes + 2 * DM_T;                 // [DM_T] exponent of the carried tile

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

  // zero everything once: padded rows / columns must stay 0 (not NaN) for the whole kernel
  for (int i = tid; i < 2 * S8 * PS + 3 * S8 * DM_LS + 2 * S8 + 4 * DM_T; i += DM_THREADS) dsm[i] = 0.0;
  int cur_tile = -1;

#pragma unroll 1
  for (int o = rg.begin; o < rg.end; ++o) {
    const OpDesc* __restrict__ op = k.ops + o;
    const int kind0 = op->kind[0], kind1 = op->kind[1];
    int tile_of[2];
    {
      const bool carried = (kind0 == SRC_CARRIED || kind1 == SRC_CARRIED) && cur_tile >= 0;
      int used = carried ? (1 << cur_tile) : 0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = ch ? kind1 : kind0;
        if (kind == SRC_CARRIED) {
          tile_of[ch] = cur_tile;
        } else if (kind == SRC_BUFFER) {
          const int tt = (used & 1) ? ((used & 2) ? 2 : 1) : 0;
          used |= 1 << tt;
          tile_of[ch] = tt;
        } else {
          tile_of[ch] = -1;
        }
      }
      // the output tile is the remaining one
      const int nxt = (used & 1) ? ((used & 2) ? 2 : 1) : 0;
      __syncthreads();  // previous op is done with tiles, P and codes
      // ---- stage inputs --------------------------------------------------------------------
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = ch ? kind1 : kind0;
        if (kind == SRC_BUFFER) {
          const double* src = static_cast<const double*>(op->src[ch]) + (int64_t)c * S * P + site0;
          double* dt = tiles + tile_of[ch] * S8 * DM_LS;
          for (int j = warp; j < S; j += 8) {
            const double2 v = ld_cg2(src + (int64_t)j * P + lane * 2);
            *reinterpret_cast<double2*>(dt + j * DM_LS + lane * 2) = v;
          }
        } else if (kind == SRC_TIP) {
          if (tid < DM_T)
            codes[ch * DM_T + tid] = (k.code_bytes == 1)
                ? (int)__ldg(static_cast<const uint8_t*>(op->src[ch]) + site0 + tid)
                : (int)__ldg(static_cast<const uint16_t*>(op->src[ch]) + site0 + tid);
        }
        const double* pm = k.pmats + (int64_t)op->pslot[ch][c] * S * S;
        double* pd = Pm + ch * S8 * PS;
        for (int idx = tid; idx < S * S; idx += DM_THREADS) {
          const int i = idx / S, j = idx - i * S;
          pd[i * PS + j] = __ldg(pm + idx);
        }
      }
      __syncthreads();
      if (kind0 == SRC_TIP || kind1 == SRC_TIP) {
        for (int idx = tid; idx < 2 * S; idx += DM_THREADS) {
          const int ch = idx / S, i = idx - ch * S;
          const double* row = Pm + ch * S8 * PS + i * PS;
          double s = 0.0;
          for (int j = 0; j < S; ++j) s += row[j];
          rowsum[ch * S8 + i] = s;
        }
        __syncthreads();
      }
      // ---- contraction + product -----------------------------------------------------------
      double prod[DM_MAX_MT][2][2];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = ch ? kind1 : kind0;
        const double* pd = Pm + ch * S8 * PS;
        double acc[DM_MAX_MT][2][2];
#pragma unroll
        for (int m = 0; m < DM_MAX_MT; ++m)
#pragma unroll
          for (int n = 0; n < 2; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
        if (kind != SRC_TIP) {
          const double* lt = tiles + tile_of[ch] * S8 * DM_LS;
          for (int k0 = 0; k0 < S4; k0 += 4) {
            const double b0 = lt[(k0 + t4) * DM_LS + n0 + g];
            const double b1 = lt[(k0 + t4) * DM_LS + n0 + 8 + g];
#pragma unroll
            for (int m = 0; m < DM_MAX_MT; ++m) {
              const int mt = wm + 2 * m;
              if (mt < n_mt) {
                const double a = pd[(mt * 8 + g) * PS + k0 + t4];
                dmma_8x8x4(acc[m][0][0], acc[m][0][1], a, b0);
                dmma_8x8x4(acc[m][1][0], acc[m][1][1], a, b1);
              }
            }
          }
        } else {
#pragma unroll
          for (int m = 0; m < DM_MAX_MT; ++m) {
            const int mt = wm + 2 * m;
            if (mt < n_mt) {
              const int i = mt * 8 + g;
#pragma unroll
              for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int cd = codes[ch * DM_T + n0 + n * 8 + 2 * t4 + e];
                  double v;
                  if (cd < S) {
                    v = pd[i * PS + cd];
                  } else if (cd == S) {
                    v = rowsum[ch * S8 + i];
                  } else {
                    const double* am = k.amb + (int64_t)(cd - S) * S;
                    v = 0.0;
                    for (int j = 0; j < S; ++j) v = fma(pd[i * PS + j], __ldg(am + j), v);
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
              if (ch == 0) prod[m][n][e] = acc[m][n][e]; else prod[m][n][e] *= acc[m][n][e];
            }
      }
      double* ot = tiles + nxt * S8 * DM_LS;
#pragma unroll
      for (int m = 0; m < DM_MAX_MT; ++m) {
        const int mt = wm + 2 * m;
        if (mt < n_mt) {
#pragma unroll
          for (int n = 0; n < 2; ++n)
            *reinterpret_cast<double2*>(ot + (mt * 8 + g) * DM_LS + n0 + n * 8 + 2 * t4) =
                make_double2(prod[m][n][0], prod[m][n][1]);
        }
      }
      __syncthreads();
      // ---- finalise: exponents, rescale, store --------------------------------------------
      int e_sum = 0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = ch ? kind1 : kind0;
        if (kind == SRC_CARRIED) e_sum += cur_e[fs];
        else if (kind == SRC_BUFFER) e_sum += __ldcg(op->src_scale[ch] + (int64_t)c * P + site0 + fs);
      }
      if (!op->is_root) {
        double m = 0.0;
        for (int j = fg; j < S; j += 4) m = fmax(m, ot[j * DM_LS + fs]);
        colmax[fg * DM_T + fs] = m;
        __syncthreads();
        m = fmax(fmax(colmax[fs], colmax[DM_T + fs]), fmax(colmax[2 * DM_T + fs], colmax[3 * DM_T + fs]));
        const int x = exponent_of(m);
        const double f = pow2_neg(x);
        double* dst = op->dst ? op->dst + (int64_t)c * S * P + site0 + fs : nullptr;
        for (int j = fg; j < S; j += 4) {
          const double v = ot[j * DM_LS + fs] * f;
          ot[j * DM_LS + fs] = v;
          if (dst) __stcg(dst + (int64_t)j * P, v);
        }
        __syncthreads();  // every reader of cur_e[] is done before it is overwritten
        if (fg == 0) {
          cur_e[fs] = e_sum + x;
          if (dst) __stcg(op->dst_scale + (int64_t)c * P + site0 + fs, e_sum + x);
        }
        cur_tile = nxt;
      } else {
        if (op->dst != nullptr) {
          double* dst = op->dst + (int64_t)c * S * P + site0 + fs;
          for (int j = fg; j < S; j += 4) __stcg(dst + (int64_t)j * P, ot[j * DM_LS + fs]);
          if (fg == 0) __stcg(op->dst_scale + (int64_t)c * P + site0 + fs, e_sum);
        }
        if (fg == 0) {
          double dot = 0.0;
          for (int j = 0; j < S; ++j) dot = fma(__ldg(k.pi + j), ot[j * DM_LS + fs], dot);
          const int64_t at = ((int64_t)rg.out_index * k.n_cats + c) * P + site0 + fs;
          __stcg(k.root_dot + at, dot);
          __stcg(k.root_exp + at, e_sum);
        }
      }
    }
  }
}

}  // namespace cb
