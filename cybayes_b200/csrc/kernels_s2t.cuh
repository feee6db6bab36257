// Two-state pruning kernel for LARGE alignments on sm_100a: tile-interleaved partials, a shared-memory
// stack for the walk, bulk-async (TMA) stores -- the HBM-write-bound version of kernels_s2.cuh.
//
// Same arithmetic as prune_s2_kernel (ML_gamma.pyx:24-38 + the exact power-of-two rescale), bit for bit;
// what changes is where the bytes live and how they move:
//
//   * Layout.  A partial buffer is cut into TILES of 32 sites (one warp).  A tile is one contiguous block of
//     32 * (2C * 8 + 4) bytes (2176 for C = 4): 2C rows of 32 doubles ([category][state][site]) followed by the
//     32 rescale exponents.  The same image is used in global memory and in shared memory, so a warp moves a
//     node's tile with ONE bulk-async copy (cp.async.bulk, SASS UBLKCP) issued by an elected lane: no per-row
//     64-bit address arithmetic, no 9 STG per thread and op, and DRAM sees 2 KB bursts instead of 256-byte
//     pieces 8 MB apart.
//   * Walk with a stack in shared memory.  A warp walks the whole op range for its tile depth-first.  The
//     child finished last is carried in registers (as before); the other child of a node with two internal
//     children is PUSHED into one of K tile buffers in shared memory and popped by its parent -- it is never
//     read back from L2/HBM.  The host orders the walk (dynamic programme over the tree, cybayes_b200.cu) so
//     that almost every such child fits the K slots (C4: 1 of 210 misses with K = 4, 31 with K = 3); the
//     misses ("spills") go through global memory with plain stores/loads.  In steady state the kernel issues
//     NO global loads except the tip codes (1 byte per site and tip, prefetched one op ahead).
//   * Op images.  Everything a block used to recompute per op while staging (P rows with the missing-data
//     column, the 9-row lookup tables of folded cherries) is built ONCE per evaluation by s2t_image_kernel
//     into a 1408-byte image per op; the main kernel streams the images of its range through a two-chunk ring
//     in shared memory with bulk-async loads completing on mbarriers (the warp that finishes a chunk last
//     issues the refill: no producer warp, no block-wide barrier in the op loop).
#pragma once
#include "cb_types.cuh"
#include "kernels_s2.cuh"

namespace cb {

constexpr int S2T_W = 32;          // sites per tile (one warp)
constexpr int S2T_WARPS = 8;       // warps per block
constexpr int S2T_THREADS = S2T_W * S2T_WARPS;
constexpr int S2T_RING_OPS = 8;    // op images per ring chunk (two chunks in flight)
constexpr int S2T_STAGING = 2;     // tile buffers 0..1 per warp rotate as store staging; stack slots follow

__host__ __device__ constexpr int s2t_tile_bytes(int C) { return S2T_W * (2 * C * 8 + 4); }

// One op as the main kernel consumes it from shared memory.
struct S2TImage {
  double tab[2][s2_child_stage(CB_S2_MAX_CATS)];  // per child: see s2_child_stage (P rows / tip rows / cherry table)
  char* dst;                 // tile-layout partial buffer, or nullptr (not stored)
  const char* src[2];        // SRC_BUFFER: tile-layout partial buffer; SRC_TIP: code row
  const void* ctip[2][2];    // SRC_CHERRY: code rows of the cherry's two tips
  int32_t kind[2];
  int32_t is_root;
  int32_t spill;             // 1: the stored partial is read back by a later op of this launch -> plain stores
  int32_t out_buf;           // shared-memory tile buffer that receives the result (-1: registers only)
  int32_t in_buf[2];         // SRC_STACK: tile buffer holding the child
  int32_t pad_[3];
};
static_assert(sizeof(S2TImage) == 1408, "S2TImage must stay 1408 bytes (a multiple of 16)");

__host__ __device__ inline size_t s2t_smem_bytes(int n_bufs, int C) {
  return (size_t)2 * S2T_RING_OPS * sizeof(S2TImage) + (size_t)S2T_WARPS * n_bufs * s2t_tile_bytes(C) + 64;
}

// ---------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ unsigned s2t_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2t_mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s2t_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void s2t_mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(s2t_smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void s2t_mbar_wait(void* bar, unsigned parity) {
  const unsigned a = s2t_smem_addr(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "S2T_WAIT_%=:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra S2T_DONE_%=;\n"
      " bra S2T_WAIT_%=;\n"
      "S2T_DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void s2t_bulk_g2s(void* smem, const void* gmem, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s2t_smem_addr(smem)),
               "l"(gmem), "r"(bytes), "r"(s2t_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void s2t_bulk_s2g(void* gmem, const void* smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem), "r"(s2t_smem_addr(smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void s2t_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void s2t_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void s2t_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ------------------------------------------------------------------------------------------ pre-pass
// One block (32 threads) per op: descriptor + per-child tables.  Exactly the arithmetic prune_s2_kernel does
// while staging a chunk, so both kernels see the same table entries bit for bit.
template <int C>
__global__ void __launch_bounds__(32) s2t_image_kernel(const LaunchConst k, int n_ops, S2TImage* __restrict__ images) {
  constexpr int CS = s2_child_stage(CB_S2_MAX_CATS);
  const int o = blockIdx.x;
  if (o >= n_ops) return;
  const OpDesc* __restrict__ op = k.ops + o;
  S2TImage* im = images + o;
  const int t = threadIdx.x;
  if (t == 0) {
    im->dst = reinterpret_cast<char*>(op->dst);
    im->src[0] = static_cast<const char*>(op->src[0]);
    im->src[1] = static_cast<const char*>(op->src[1]);
    im->ctip[0][0] = op->ctip[0][0]; im->ctip[0][1] = op->ctip[0][1];
    im->ctip[1][0] = op->ctip[1][0]; im->ctip[1][1] = op->ctip[1][1];
    im->kind[0] = op->kind[0]; im->kind[1] = op->kind[1];
    im->is_root = op->is_root;
    im->spill = op->spill;
    im->out_buf = op->out_buf;
    im->in_buf[0] = op->in_buf[0]; im->in_buf[1] = op->in_buf[1];
    im->pad_[0] = im->pad_[1] = im->pad_[2] = 0;
  }
  // P matrices of internal and tip children (one thread per child, category and row)
  for (int idx = t; idx < 4 * C; idx += 32) {
    const int i = idx & 1, c = (idx >> 1) % C, ch = idx / (2 * C);
    const int kind = op->kind[ch];
    if (kind == SRC_CHERRY) continue;
    const double2 pr = __ldg(reinterpret_cast<const double2*>(s2_pmat(k, op->pslot[ch][c])) + i);
    double* base = &im->tab[ch][0];
    if (kind == SRC_TIP) {
      double* q = base + (c * 2 + i) * 4;
      q[0] = pr.x;
      q[1] = pr.y;
      q[2] = s2_tip_term(pr, 2);
      q[3] = 0.0;
    } else {
      reinterpret_cast<double2*>(base)[c * 2 + i] = pr;
    }
  }
  // lookup tables of folded cherries (one thread per child and pair of tip codes)
  for (int idx = t; idx < 18; idx += 32) {
    const int q = idx % 9, ch = idx / 9;
    if (op->kind[ch] != SRC_CHERRY) continue;
    const int ca = q / 3, cb_ = q - 3 * ca;
    double L[C][2];
    int mh = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double2* pa = reinterpret_cast<const double2*>(s2_pmat(k, op->cslot[ch][0][c]));
      const double2* pb = reinterpret_cast<const double2*>(s2_pmat(k, op->cslot[ch][1][c]));
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        L[c][j] = s2_tip_term(__ldg(pa + j), ca) * s2_tip_term(__ldg(pb + j), cb_);
        mh = max(mh, __double2hiint(L[c][j]));
      }
    }
    const int be = (mh >> 20) & 0x7ff;
    const int x = (be == 0 || be == 0x7ff) ? 0 : be - 1023;
    const double f = pow2_neg(x);
    double* tab = &im->tab[ch][0] + q * (2 * C + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const double l0 = L[c][0] * f, l1 = L[c][1] * f;
      const double2* pp = reinterpret_cast<const double2*>(s2_pmat(k, op->pslot[ch][c]));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double2 pr = __ldg(pp + i);
        tab[c * 2 + i] = fma(pr.y, l1, pr.x * l0);
      }
    }
    tab[2 * C] = (double)x;
    // a cherry that stays in the returned cache keeps its own copy of its two edges' P matrices
    if (q == 0 && op->crec_out[ch] >= 0) {
      double* out = k.pmats_lib + (int64_t)op->crec_out[ch] * 2 * C * 4;
#pragma unroll
      for (int tt = 0; tt < 2; ++tt)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const double* src = s2_pmat(k, op->cslot[ch][tt][c]);
#pragma unroll
          for (int e = 0; e < 4; ++e) out[(tt * C + c) * 4 + e] = __ldg(src + e);
        }
    }
  }
  (void)CS;
}

// ---------------------------------------------------------------------------------------- main kernel
__device__ __forceinline__ void s2t_fetch_codes(const S2TImage& im, int code_bytes, int64_t site, unsigned (&code)[2][2]) {
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    if (im.kind[ch] == SRC_TIP) {
      unsigned c1[1];
      load_codes<1>(im.src[ch], code_bytes, site, c1);
      code[ch][0] = c1[0];
    } else if (im.kind[ch] == SRC_CHERRY) {
      unsigned c1[1];
      load_codes<1>(im.ctip[ch][0], code_bytes, site, c1);
      code[ch][0] = c1[0];
      load_codes<1>(im.ctip[ch][1], code_bytes, site, c1);
      code[ch][1] = c1[0];
    }
  }
}

template <int C>
__global__ void __launch_bounds__(S2T_THREADS, 2) prune_s2t_kernel(const LaunchConst k, const S2TImage* __restrict__ images,
                                                                    int n_bufs) {
  static_assert(C <= CB_S2_MAX_CATS, "2-state kernel supports at most CB_S2_MAX_CATS categories");
  constexpr int TB = s2t_tile_bytes(C);
  constexpr int CHUNK_BYTES = S2T_RING_OPS * (int)sizeof(S2TImage);
  extern __shared__ __align__(128) unsigned char s2t_smem[];
  __shared__ double red[32];
  __shared__ int last_flag;
  __shared__ int done_cnt[2];
  __shared__ __align__(8) unsigned long long full_bar[2];

  const RangeDesc rg = k.ranges[blockIdx.y];
  const int nops = rg.end - rg.begin;
  const int n_chunks = (nops + S2T_RING_OPS - 1) / S2T_RING_OPS;
  const S2TImage* __restrict__ gimg = images + rg.begin;
  S2TImage* ring = reinterpret_cast<S2TImage*>(s2t_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* mybufs = s2t_smem + (size_t)2 * CHUNK_BYTES + (size_t)warp * n_bufs * TB;

  const int64_t n_tiles = k.n_sites / S2T_W;
  const int64_t tile = (int64_t)blockIdx.x * S2T_WARPS + warp;
  const int64_t tiles_left = n_tiles - (int64_t)blockIdx.x * S2T_WARPS;
  const int n_active = tiles_left < S2T_WARPS ? (int)tiles_left : S2T_WARPS;   // warps of this block that own a tile
  const bool active = warp < n_active;
  const int64_t site = tile * S2T_W + lane;
  const size_t tile_off = (size_t)tile * TB;

  auto issue_chunk = [&](int c) {  // one thread: bulk-load the images of chunk c into its ring half
    const int b = c & 1;
    const int n = min(S2T_RING_OPS, nops - c * S2T_RING_OPS);
    const unsigned bytes = (unsigned)n * (unsigned)sizeof(S2TImage);
    s2t_mbar_expect_tx(&full_bar[b], bytes);
    s2t_bulk_g2s(ring + (size_t)b * S2T_RING_OPS, gimg + (size_t)c * S2T_RING_OPS, bytes, &full_bar[b]);
  };

  if (threadIdx.x == 0) {
    s2t_mbar_init(&full_bar[0], 1);
    s2t_mbar_init(&full_bar[1], 1);
    done_cnt[0] = done_cnt[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    issue_chunk(0);
    if (n_chunks > 1) issue_chunk(1);
  }
  __syncthreads();

  double lnl = 0.0;
  if (active) {
    double cur[C][2];  // carried partial
    int cur_e = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) cur[c][0] = cur[c][1] = 0.0;
    unsigned code_next[2][2] = {{0u, 0u}, {0u, 0u}};
    bool groups_pending = false;

    s2t_mbar_wait(&full_bar[0], 0);
    s2t_fetch_codes(ring[0], k.code_bytes, site, code_next);

#pragma unroll 1
    for (int o = 0; o < nops; ++o) {
      const int chunk = o / S2T_RING_OPS, oc = o - chunk * S2T_RING_OPS;
      const S2TImage& im = ring[(size_t)(chunk & 1) * S2T_RING_OPS + oc];
      unsigned code[2][2];
#pragma unroll
      for (int ch = 0; ch < 2; ++ch)
#pragma unroll
        for (int t = 0; t < 2; ++t) code[ch][t] = code_next[ch][t];
      if (o + 1 < nops) {  // tip codes one op ahead; entering a new chunk also pulls the following chunk's tips to L2
        const int c1 = (o + 1) / S2T_RING_OPS, o1 = (o + 1) - c1 * S2T_RING_OPS;
        if (o1 == 0) {
          s2t_mbar_wait(&full_bar[c1 & 1], (unsigned)(c1 >> 1) & 1u);
          const int nxt0 = (c1 + 1) * S2T_RING_OPS;  // first op of the chunk after the one we enter
          const int e = nxt0 + (lane >> 2);
          if (e < nops && warp == c1 % n_active) {  // one warp per block and chunk does it
            const OpDesc* __restrict__ dn = k.ops + rg.begin + e;
            const int ch = (lane >> 1) & 1, t = lane & 1;
            const int kind = dn->kind[ch];
            const void* row = kind == SRC_TIP ? (t == 0 ? dn->src[ch] : nullptr) : (kind == SRC_CHERRY ? dn->ctip[ch][t] : nullptr);
            if (row != nullptr) {
              const char* a = static_cast<const char*>(row) + (tile - warp) * S2T_W * k.code_bytes;   // the block's 256 sites
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a + 128));
            }
          }
        }
        s2t_fetch_codes(ring[(size_t)(c1 & 1) * S2T_RING_OPS + o1], k.code_bytes, site, code_next);
      }

      double out[C][2];
      int e_in = 0;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int kind = im.kind[ch];
        const double* pbase = &im.tab[ch][0];
        const double2* pm = reinterpret_cast<const double2*>(pbase);
#define CB_S2T_APPLY(L0, L1)                                            \
  _Pragma("unroll") for (int c = 0; c < C; ++c) {                       \
    _Pragma("unroll") for (int i = 0; i < 2; ++i) {                     \
      const double2 pr = pm[c * 2 + i];                                  \
      const double x = fma(pr.y, (L1), pr.x * (L0));                     \
      if (ch == 0) out[c][i] = x; else out[c][i] *= x;                   \
    }                                                                    \
  }
        if (kind == SRC_CARRIED) {
          CB_S2T_APPLY(cur[c][0], cur[c][1])
          e_in += cur_e;
        } else if (kind == SRC_STACK) {
          const double* sb = reinterpret_cast<const double*>(mybufs + (size_t)im.in_buf[ch] * TB) + lane;
          double L[C][2];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            L[c][0] = sb[(2 * c) * S2T_W];
            L[c][1] = sb[(2 * c + 1) * S2T_W];
          }
          const int se = reinterpret_cast<const int*>(sb - lane + 2 * C * S2T_W)[lane];
          CB_S2T_APPLY(L[c][0], L[c][1])
          e_in += se;
        } else if (kind == SRC_BUFFER) {
          const double* gb = reinterpret_cast<const double*>(im.src[ch] + tile_off) + lane;
          double L[C][2];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            L[c][0] = __ldcg(gb + (2 * c) * S2T_W);
            L[c][1] = __ldcg(gb + (2 * c + 1) * S2T_W);
          }
          const int se = __ldcg(reinterpret_cast<const int*>(gb - lane + 2 * C * S2T_W) + lane);
          CB_S2T_APPLY(L[c][0], L[c][1])
          e_in += se;
        } else if (kind == SRC_TIP) {
          const unsigned cd = min(code[ch][0], 2u);
#pragma unroll
          for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double x = pbase[(c * 2 + i) * 4 + cd];
              if (ch == 0) out[c][i] = x; else out[c][i] *= x;
            }
          }
        } else {  // folded cherry: the pair of tip codes selects a precomputed row
          const double* tab = pbase + (min(code[ch][0], 2u) * 3 + min(code[ch][1], 2u)) * (2 * C + 1);
#pragma unroll
          for (int c = 0; c < C; ++c) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double x = tab[c * 2 + i];
              if (ch == 0) out[c][i] = x; else out[c][i] *= x;
            }
          }
          e_in += (int)tab[2 * C];
        }
#undef CB_S2T_APPLY
      }

      const bool is_root = im.is_root != 0;
      int e_out = e_in;
      if (!is_root) {
        int mh = __double2hiint(out[0][0]);
#pragma unroll
        for (int c = 0; c < C; ++c) mh = max(mh, max(__double2hiint(out[c][0]), __double2hiint(out[c][1])));
        const int be = (mh >> 20) & 0x7ff;
        const int x = (be == 0 || be == 0x7ff) ? 0 : be - 1023;
        const double f = pow2_neg(x);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          cur[c][0] = out[c][0] * f;
          cur[c][1] = out[c][1] * f;
        }
        cur_e = e_in + x;
        e_out = cur_e;
      }
      // what gets written: the rescaled partial of an ordinary node, the partial as computed for an (optionally stored) root
      char* const dst = im.dst;
      const int ob = im.out_buf;
      if (ob >= 0) {
        if (groups_pending) {  // the bulk copy that last read this tile buffer is at least two groups old (host rule)
          if (lane == 0) s2t_bulk_wait_read<1>();
          __syncwarp();
        }
        double* sb = reinterpret_cast<double*>(mybufs + (size_t)ob * TB);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          sb[(2 * c) * S2T_W + lane] = is_root ? out[c][0] : cur[c][0];
          sb[(2 * c + 1) * S2T_W + lane] = is_root ? out[c][1] : cur[c][1];
        }
        reinterpret_cast<int*>(sb + 2 * C * S2T_W)[lane] = e_out;
        if (dst != nullptr && !im.spill) {
          s2t_fence_async_smem();
          __syncwarp();
          if (lane == 0) s2t_bulk_s2g(dst + tile_off, sb, TB);
          groups_pending = true;
        }  // (a pushed tile is popped by the lanes that wrote it: columns are lane-private, no barrier needed)
      }
      if (dst != nullptr && (ob < 0 || im.spill)) {  // plain coalesced stores (spills are re-read through the generic proxy)
        double* gb = reinterpret_cast<double*>(dst + tile_off) + lane;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          __stcg(gb + (2 * c) * S2T_W, is_root ? out[c][0] : cur[c][0]);
          __stcg(gb + (2 * c + 1) * S2T_W, is_root ? out[c][1] : cur[c][1]);
        }
        __stcg(reinterpret_cast<int*>(gb - lane + 2 * C * S2T_W) + lane, e_out);
      }
      if (is_root) {
        // ll_p = sum_c (pi . L_c) / n_cats ; lnL += w_p * log(ll_p)      ML_gamma.pyx:38,40
        const double pi0 = __ldg(k.pi), pi1 = __ldg(k.pi + 1);
        const double w = __ldg(k.weights + site);
        const double ln2 = 0.693147180559945309417232121458;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < C; ++c) s += fma(pi1, out[c][1], pi0 * out[c][0]) / k.cats;
        if (w != 0.0) lnl += w * (log(s) + (double)e_in * ln2);
      }

      // leaving a chunk: the warp that finishes it last refills its ring half with the chunk two ahead
      if (oc == S2T_RING_OPS - 1 && chunk + 2 < n_chunks) {
        __syncwarp();
        if (lane == 0) {
          const int b = chunk & 1;
          if (atomicAdd(&done_cnt[b], 1) == n_active - 1) {
            done_cnt[b] = 0;
            s2t_fence_async_smem();  // the warps' reads of this ring half before the async writes
            issue_chunk(chunk + 2);
          }
        }
      }
    }
    if (groups_pending && lane == 0) s2t_bulk_wait_all();
  }
  if (rg.out_index >= 0) block_reduce_to_result(lnl, k, rg.out_index, red, &last_flag);
}

}  // namespace cb
