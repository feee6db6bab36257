// Multistate (32 <= S <= 64) pruning on the FP64 tensor path, register-carried variant for LARGE
// alignments (sm_100a).  Same recursion as kernels_dmma.cuh / ML_gamma.pyx:24-38, transposed:
//
//     V^T[site][i] = sum_j L^T[site][j] * P^T[j][i]        (16 sites x S) @ (S x S)   per warp
//
// so the child partial is the A operand of `mma.sync.m8n8k4.f64` and P^T the B operand.  In that
// form the accumulator fragment a thread ends up with (site g; states 8n+2t, 8n+2t+1) is exactly
// the A fragment the NEXT contraction needs, provided the k index is enumerated in the permuted
// order (n, e) -> state 8n + 2t + e -- a contraction may sum its terms in any fixed order.  The
// on-path (carried) partial therefore never leaves registers: no shared-memory tile, no
// block-wide barrier in the op loop.  Each of the 8 warps owns 16 sites x all states and walks the
// op range on its own; the only thing the warps share is the read-only P matrix of the current
// child, streamed through a ring of shared-memory stages:
//
//   * A pre-pass (rc_restage_kernel) lays every P matrix the ops use out ONCE as the exact stage
//     image (rows padded to a multiple of 8 states, 16-byte units XOR-swizzled by row bits, zero
//     padding), in the order the ops consume them.  A stage is then filled by bulk-async (TMA)
//     copies completing on an mbarrier -- no LSU work, no registers, no producer warp: the warp that
//     is LAST to finish with a stage (shared-memory counter) re-arms its barrier and issues the
//     copy of the matrix NST jobs ahead.  The op descriptor travels with the first child's matrix.
//   * The swizzle makes the B fragments of two k blocks one conflict-free LDS.128.
//   * For a TIP child the staged image is P TRANSPOSED, followed by precomputed rows for the non-state codes: row S
//     holds the row sums of P ('?' / '-' cells), rows S+1.. the dot products with the other ambiguity sets.  A tip
//     child is then one row per site, read as 16-byte pairs straight into the accumulator layout: no tensor work, no
//     per-warp row sums, no shuffles, no branch on missing data.
//   * Buffer children (nodes with two internal children, inputs of a dirty path) are loaded from
//     HBM straight into A-fragment registers: per request 4 rows x 64 contiguous bytes, every
//     sector fully used; a thread reads back exactly the addresses it wrote itself.
//   * Per site rescale (power of two, integer max of the high words) and the store of the cached
//     partial come straight from the accumulator registers; max / pi-dot reductions are two
//     shuffles inside a quad.
//   * Warps that walk the ops in lockstep leave the FP64 pipe idle whenever they all gather / rescale
//     at the same time.  Warps 4..7 (the second warp of each SM sub-partition) may only start a stage
//     once warps 0..3 have finished it, so while one warp of a sub-partition is between contractions
//     its neighbour is inside one (rc_stagger).
//
// One block = 128 sites x op range x ONE rate category, 8 warps (two per sub-partition, 255 registers),
// one block per SM.
#pragma once
#include "cb_types.cuh"
#include "kernels_s2.cuh"

namespace cb {

constexpr int RC_WARPS = 8;
constexpr int RC_THREADS = RC_WARPS * 32;
constexpr int RC_WSITES = 16;                     // sites per warp (two 8-row m-tiles)
constexpr int RC_T = RC_WARPS * RC_WSITES;        // sites per block
constexpr int RC_DESC_D = sizeof(OpDesc) / 8;     // doubles taken by the op descriptor copy of a stage
constexpr int RC_EXTRA_ROWS = 8;                  // rows of a tip image beyond the padded states: codes S .. (row sums, ambiguity sets)

// All kernels here are compiled per PADDED state count S8 (a multiple of 8: 32, 40, 48, 56, 64); the real state count
// S (S8 - 8 < S <= S8) is a run-time value that only matters in the last state tile.
template <int S8_> struct RcCfg {
  static constexpr int S8 = S8_;
  static_assert(S8 % 8 == 0, "padded state count");
  static constexpr int NT = S8 / 8;                                  // 8-state tiles (n tiles = k block pairs)
  static constexpr int PS = S8;                                      // P row stride in a stage (unpadded, swizzled)
  static constexpr bool SWZ_HALF = (S8 % 16 == 0);                   // rows start in the same 128-byte bank window
  static constexpr int ROWS = S8 + RC_EXTRA_ROWS;                    // rows of a stage image
  static constexpr int MAT_D = ROWS * PS;
  static constexpr unsigned MAT_BYTES = MAT_D * 8;
  static constexpr int STAGE_D = MAT_D + RC_DESC_D;
  static constexpr int NST_FIT = (226 * 1024) / (STAGE_D * 8);
  static constexpr int NST = NST_FIT > 8 ? 8 : NST_FIT;              // ring depth
  static constexpr size_t SMEM = (size_t)NST * STAGE_D * 8 + 3 * 8 * 8;  // + full / lag barriers, release counters
  static_assert(NST >= 3, "state count too large for the stage ring");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(MAT_BYTES % 64 == 0, "bulk copies are issued in four 16-byte aligned parts");
};

// swizzle of row `row` of a stage, as a mask on the (double) column index: 16-byte unit index ^= row bits (2,1)
// and, when the row stride is a multiple of 128 bytes, row bit 0 -> unit bit 2 (consecutive rows in different halves)
template <bool HALF> __host__ __device__ __forceinline__ int rc_swz(int row) {
  return ((((row >> 1) & 3) | (HALF ? (row & 1) << 2 : 0)) << 1);
}

__device__ __forceinline__ unsigned rc_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rc_mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(rc_smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void rc_mbar_arrive(void* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(rc_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void rc_mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(rc_smem_addr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void rc_mbar_wait(void* bar, unsigned parity) {
  const unsigned a = rc_smem_addr(bar);
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      "RC_WAIT_%=:\n"
      " mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra RC_DONE_%=;\n"
      " bra RC_WAIT_%=;\n"
      "RC_DONE_%=:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool rc_mbar_test(void* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      " .reg .pred p;\n"
      " mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      " selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(rc_smem_addr(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// bulk-async copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void rc_bulk_g2s(void* smem, const void* gmem, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(rc_smem_addr(smem)),
               "l"(gmem), "r"(bytes), "r"(rc_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void rc_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void rc_dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

// Pre-pass of an evaluation: staged[((o * 2 + step) * C + c)] <- pool slot ops[o].pslot[child(step)][c], as a stage image.
//   internal child:  rows i < S8: P[i][.] with the 16-byte units of the row swizzled by rc_swz(i) (B fragments);
//   tip child:       row r < S: P[.][r] (P transposed); row S: sum_j P[.][j]; row S + k: sum_j P[.][j] amb[k][j];
//                    unit index ^= (r & 1) << 2 when rows are a multiple of 128 bytes (two sites' rows in different banks).
// One block per matrix.
template <int S8>
__global__ void __launch_bounds__(256) rc_restage_kernel(const LaunchConst k, int n_ops, double* __restrict__ staged) {
  using Cfg = RcCfg<S8>;
  constexpr int U = S8 / 2, ROWS = Cfg::ROWS;   // U: 16-byte units per row
  const int S = k.n_states;
  const int job = blockIdx.x;               // (o * 2 + step) * C + c
  const int C = k.n_cats;
  const int c = job % C, os = job / C, step = os & 1, o = os >> 1;
  if (o >= n_ops) return;
  const OpDesc* op = k.ops + o;
  const int first = (op->kind[1] == SRC_CARRIED) ? 1 : 0;
  const int child = step ? 1 - first : first;
  const double* __restrict__ pm = k.pmats + (int64_t)op->pslot[child][c] * S * S;
  double2* dst = reinterpret_cast<double2*>(staged + (int64_t)job * Cfg::MAT_D);
  if (op->kind[child] != SRC_TIP) {
    for (int idx = threadIdx.x; idx < ROWS * U; idx += blockDim.x) {
      const int r = idx / U, u = idx - r * U;
      double2 v = make_double2(0.0, 0.0);
      if (r < S) {
        if (2 * u < S) v.x = __ldg(pm + r * S + 2 * u);
        if (2 * u + 1 < S) v.y = __ldg(pm + r * S + 2 * u + 1);
      }
      dst[r * U + (((2 * u) ^ rc_swz<Cfg::SWZ_HALF>(r)) >> 1)] = v;
    }
    return;
  }
  for (int idx = threadIdx.x; idx < ROWS * U; idx += blockDim.x) {
    const int r = idx / U, u = idx - r * U;   // row = tip code, columns = states 2u, 2u + 1
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * u + e;
      if (i >= S) continue;
      const double* row = pm + i * S;
      if (r < S) {
        v[e] = __ldg(row + r);
      } else if (r == S) {
        double s = 0.0;
        for (int j = 0; j < S; ++j) s += __ldg(row + j);
        v[e] = s;
      } else if (r - S < k.n_amb) {
        const double* am = k.amb + (int64_t)(r - S) * S;
        double s = 0.0;
        for (int j = 0; j < S; ++j) s = fma(__ldg(row + j), __ldg(am + j), s);
        v[e] = s;
      }
    }
    dst[r * U + (u ^ (Cfg::SWZ_HALF ? (r & 1) << 2 : 0))] = make_double2(v[0], v[1]);
  }
}

template <int S8>
__device__ __forceinline__ void rc_load_codes(int (&cd)[2], const void* src, int64_t wsite, int code_bytes) {
#pragma unroll
  for (int m = 0; m < 2; ++m)
    cd[m] = (code_bytes == 1) ? (int)__ldg(static_cast<const uint8_t*>(src) + wsite + 8 * m)
                              : (int)__ldg(static_cast<const uint16_t*>(src) + wsite + 8 * m);
}

// A tip child: acc (x)= the row of the transposed image selected by the site's code (a state, the all-ones set or
// another ambiguity set), two states per 16-byte load.  MUL: multiply into acc instead of overwriting it.
template <int S8, bool MUL>
__device__ __forceinline__ void rc_tip_child(double (&acc)[2][RcCfg<S8>::NT][2], const double* Pm, const int (&cd)[2], int t4,
                                             int S, const LaunchConst& k) {
  constexpr int NT = RcCfg<S8>::NT, PS = RcCfg<S8>::PS, ROWS = RcCfg<S8>::ROWS;
  constexpr bool HALF = RcCfg<S8>::SWZ_HALF;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int row = cd[m];
    if (row < ROWS) {
      const double* rp = Pm + row * PS + 2 * t4;
      const int par = HALF ? (row & 1) : 0;   // odd rows keep their 64-byte halves swapped
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double2 v = *reinterpret_cast<const double2*>(rp + 8 * (n ^ par));
        acc[m][n][0] = MUL ? acc[m][n][0] * v.x : v.x;
        acc[m][n][1] = MUL ? acc[m][n][1] * v.y : v.y;
      }
    } else {
      // more ambiguity sets than spare rows: dense dot of the transposed columns with the 0/1 membership vector
      const double* am = k.amb + (int64_t)(row - S) * S;
#pragma unroll
      for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double v = 0.0;
#pragma unroll 1
          for (int j = 0; j < S; ++j) {
            const int col = (8 * n + 2 * t4 + e) ^ ((HALF && (j & 1)) ? 8 : 0);
            v = fma(Pm[j * PS + col], __ldg(am + j), v);
          }
          acc[m][n][e] = MUL ? acc[m][n][e] * v : v;
        }
    }
  }
}

// acc = cur (16 sites x S, A fragments in registers) @ P^T (B fragments from the stage).
// k blocks (j,0), (j,1) of one state tile come from a single conflict-free LDS.128.
template <int S8>
__device__ __forceinline__ void rc_contract(double (&acc)[2][RcCfg<S8>::NT][2], const double (&cur)[2][RcCfg<S8>::NT][2],
                                            const double* Pm, int g, int t4) {
  constexpr int NT = RcCfg<S8>::NT, PS = RcCfg<S8>::PS;
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
  // B fragment of (state tile n, k pair j): row 8n + g, logical column 8j + 2*t4.  The row swizzle only depends
  // on g: (8j + 2 t4) ^ swz = 8 (j ^ hb) + ((2 t4) ^ lo), so two lane-constant bases (even / odd j) make every
  // address base + compile-time offset.
  const int swg = rc_swz<RcCfg<S8>::SWZ_HALF>(g);
  const int hb = (swg >> 3) & 1;
  const double* p_even = Pm + g * PS + ((2 * t4) ^ (swg & 7)) + 8 * hb;
  const double* p_odd = p_even - 16 * hb;
  // Issue order: per k pair j, groups of up to four state tiles -- first all their (j,0) blocks, then all their (j,1)
  // blocks -- so a DMMA never depends on one of the 7 issued before it (a lone warp would otherwise wait on the
  // ~27-cycle accumulate latency every other instruction).
  constexpr int NG = 4;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const double* pj = ((j & 1) ? p_odd : p_even) + 8 * j;
#pragma unroll
    for (int n0 = 0; n0 < NT; n0 += NG) {
      double2 b[NG];
#pragma unroll
      for (int i = 0; i < NG; ++i)
        if (n0 + i < NT) b[i] = *reinterpret_cast<const double2*>(pj + (8 * (n0 + i)) * PS);
#pragma unroll
      for (int i = 0; i < NG; ++i)
        if (n0 + i < NT) {
          rc_dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][0], b[i].x);
          rc_dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][0], b[i].x);
        }
#pragma unroll
      for (int i = 0; i < NG; ++i)
        if (n0 + i < NT) {
          rc_dmma(acc[0][n0 + i][0], acc[0][n0 + i][1], cur[0][j][1], b[i].y);
          rc_dmma(acc[1][n0 + i][0], acc[1][n0 + i][1], cur[1][j][1], b[i].y);
        }
    }
  }
}

// a stored partial -> A-fragment registers (+ its exponents): per request 4 rows x 64 contiguous bytes
template <int S8>
__device__ __forceinline__ void rc_load_buffer(double (&cur)[2][RcCfg<S8>::NT][2], int (&e_sum)[2], const void* src,
                                               const int32_t* sscale, int c, int S, int64_t P, int64_t wsite, int t4) {
  constexpr int NT = RcCfg<S8>::NT;
  const double* sp = static_cast<const double*>(src) + (int64_t)c * S * P + wsite;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int stt = 8 * n + 2 * t4 + e;
        cur[m][n][e] = (n < NT - 1 || stt < S) ? __ldcg(sp + (int64_t)stt * P + 8 * m) : 0.0;
      }
    e_sum[m] += __ldcg(sscale + (int64_t)c * P + wsite + 8 * m);
  }
}

// Fill the stage of job q (= 2 * (op - begin) + step) of this block: re-arm its barrier with the byte count, then
// the matrix in four bulk copies and, for a first child, the op descriptor.  Executed by ONE thread.
template <int S8>
__device__ __forceinline__ void rc_issue_job(const LaunchConst& k, const RangeDesc& rg, int c, int q, double* rsm,
                                             unsigned long long* full_bar) {
  using Cfg = RcCfg<S8>;
  const int st = q % Cfg::NST, step = q & 1, o = rg.begin + (q >> 1);
  double* stage = rsm + (size_t)st * Cfg::STAGE_D;
  const char* src = reinterpret_cast<const char*>(k.staged + ((int64_t)(o * 2 + step) * k.n_cats + c) * Cfg::MAT_D);
  constexpr unsigned PART = Cfg::MAT_BYTES / 4;
  rc_mbar_expect_tx(full_bar + st, Cfg::MAT_BYTES + (step == 0 ? (unsigned)sizeof(OpDesc) : 0u));
#pragma unroll
  for (int p = 0; p < 4; ++p) rc_bulk_g2s(reinterpret_cast<char*>(stage) + p * PART, src + p * PART, PART, full_bar + st);
  if (step == 0) rc_bulk_g2s(stage + Cfg::MAT_D, k.ops + o, (unsigned)sizeof(OpDesc), full_bar + st);
}

// EXACT: the state count is S8 itself (a compile-time value: the padding guards of the last state tile fold away).
template <int S8, bool EXACT>
__global__ void __launch_bounds__(RC_THREADS, 1) prune_dmma_rc_kernel(const LaunchConst k) {
  using Cfg = RcCfg<S8>;
  constexpr int NT = Cfg::NT, NST = Cfg::NST;
  const int S = EXACT ? S8 : k.n_states;
  extern __shared__ __align__(128) double rsm[];
  unsigned long long* full_bar = reinterpret_cast<unsigned long long*>(rsm + (size_t)NST * Cfg::STAGE_D);
  unsigned long long* lag_bar = full_bar + 8;                         // warps 0..3 have finished a stage
  int* done_cnt = reinterpret_cast<int*>(lag_bar + 8);                // warps that have finished a stage

  const RangeDesc rg = k.ranges[blockIdx.y];
  const int c = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t P = k.n_sites;
  const int64_t site0 = (int64_t)blockIdx.x * RC_T;
  const int64_t left = P - site0;
  const int n_active = left >= RC_T ? RC_WARPS : (int)(left / RC_WSITES);  // P is a multiple of 64: 4 or 8
  const int n_jobs = 2 * (rg.end - rg.begin);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      rc_mbar_init(full_bar + s, 1);                             // the issuer's expect_tx arrive; the copies complete the bytes
      rc_mbar_init(lag_bar + s, n_active < 4 ? n_active : 4);
      done_cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");   // barriers visible to the async proxy
    for (int q = 0; q < NST && q < n_jobs; ++q) rc_issue_job<S8>(k, rg, c, q, rsm, full_bar);
  }
  __syncthreads();  // the only block-wide barrier
  if (warp >= n_active) return;

  const int g = lane >> 2, t4 = lane & 3;               // mma groupID / threadID_in_group
  const int64_t wsite = site0 + warp * RC_WSITES + g;   // this thread's sites: wsite, wsite + 8
  const int grp = warp >> 2;
  const bool lag_wait = k.rc_stagger != 0 && grp > 0;
  const bool lag_post = k.rc_stagger != 0 && grp == 0 && n_active > 4;
  double cur[2][NT][2];                                 // carried partial, [m-tile][state tile][e]: state 8n + 2*t4 + e
  int cur_e[2] = {0, 0};
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) cur[m][n][0] = cur[m][n][1] = 0.0;

  // this warp is done with stage st (job q): the last warp to say so refills the stage with job q + NST
  auto release = [&](int st, int q) {
    __syncwarp();
    if (lane == 0) {
      if (lag_post) rc_mbar_arrive(lag_bar + st);
      __threadfence_block();
      if (atomicAdd(done_cnt + st, 1) == n_active - 1) {
        done_cnt[st] = 0;
        if (q + NST < n_jobs) {
          asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // the warps' reads before the async writes
          rc_issue_job<S8>(k, rg, c, q + NST, rsm, full_bar);
        }
      }
    }
  };

  int cdn[2] = {0, 0};   // tip codes of the NEXT op's first child, requested while this op rescales (when its stage is in)
  bool have_cdn = false;
  int st = 0, round = 0, q = 0;
#pragma unroll 1
  for (int o = rg.begin; o < rg.end; ++o) {
    double acc[2][NT][2];
    int e_sum[2] = {0, 0};
    // ---- first child (the carried one when there is one); the op descriptor travels with its matrix ----
    const double* Pm = rsm + (size_t)st * Cfg::STAGE_D;
    rc_mbar_wait(full_bar + st, (unsigned)(round & 1));
    if (lag_wait) rc_mbar_wait(lag_bar + st, (unsigned)(round & 1));
    const OpDesc* d = reinterpret_cast<const OpDesc*>(Pm + Cfg::MAT_D);
    const int first = (d->kind[1] == SRC_CARRIED) ? 1 : 0;
    const int kind0 = d->kind[first], kind1 = d->kind[1 - first];
    const void* src1 = d->src[1 - first];
    const int32_t* scale1 = d->src_scale[1 - first];
    const int is_root = d->is_root;
    double* dst = d->dst;
    int32_t* dst_scale = d->dst_scale;
    int cd1[2] = {0, 0};
    if (kind1 == SRC_TIP) {
      rc_load_codes<S8>(cd1, src1, wsite, k.code_bytes);  // in flight during the first child's contraction
    } else {
      // the stored partial of the second child: pull its 16 sites x S doubles towards L2 now
      const char* sp = reinterpret_cast<const char*>(static_cast<const double*>(src1) + (int64_t)c * S * P + wsite - g);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int stt = lane + 32 * (p & 1);
        if (stt < S) rc_prefetch_l2(sp + ((int64_t)stt * P + 8 * (p >> 1)) * 8);
      }
    }
    if (kind0 == SRC_TIP) {
      int cd0[2] = {cdn[0], cdn[1]};
      if (!have_cdn) rc_load_codes<S8>(cd0, d->src[first], wsite, k.code_bytes);
      rc_tip_child<S8, false>(acc, Pm, cd0, t4, S, k);
    } else {
      if (kind0 == SRC_BUFFER) {
        rc_load_buffer<S8>(cur, e_sum, d->src[first], d->src_scale[first], c, S, P, wsite, t4);
      } else {
        e_sum[0] += cur_e[0];
        e_sum[1] += cur_e[1];
      }
      rc_contract<S8>(acc, cur, Pm, g, t4);
    }
    release(st, q);
    ++q;
    if (++st == NST) { st = 0; ++round; }
    // ---- second child: a tip multiplies into acc; a stored partial is contracted into a second accumulator set ----
    Pm = rsm + (size_t)st * Cfg::STAGE_D;
    rc_mbar_wait(full_bar + st, (unsigned)(round & 1));
    if (lag_wait) rc_mbar_wait(lag_bar + st, (unsigned)(round & 1));
    if (kind1 == SRC_TIP) {
      rc_tip_child<S8, true>(acc, Pm, cd1, t4, S, k);
    } else {
      double acc2[2][NT][2];
      rc_load_buffer<S8>(cur, e_sum, src1, scale1, c, S, P, wsite, t4);
      rc_contract<S8>(acc2, cur, Pm, g, t4);
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) { acc[m][n][0] *= acc2[m][n][0]; acc[m][n][1] *= acc2[m][n][1]; }
    }
    release(st, q);
    ++q;
    if (++st == NST) { st = 0; ++round; }
    // a look at the next op: if its first stage has landed and starts with a tip, request the codes now
    have_cdn = false;
    if (o + 1 < rg.end && __all_sync(0xffffffffu, rc_mbar_test(full_bar + st, (unsigned)(round & 1)))) {
      const OpDesc* d2 = reinterpret_cast<const OpDesc*>(rsm + (size_t)st * Cfg::STAGE_D + Cfg::MAT_D);
      const int f2 = (d2->kind[1] == SRC_CARRIED) ? 1 : 0;
      if (d2->kind[f2] == SRC_TIP) {
        rc_load_codes<S8>(cdn, d2->src[f2], wsite, k.code_bytes);
        have_cdn = true;
      }
    }

    // ---- rescale, store; the result stays in registers as the carried partial --------------------
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int64_t site = wsite + 8 * m;
      if (!is_root) {
        // partials are >= 0, so the largest value has the largest high word: integer max, exponent from it
        int hm = 0;
#pragma unroll
        for (int n = 0; n < NT; ++n) hm = max(hm, max(__double2hiint(acc[m][n][0]), __double2hiint(acc[m][n][1])));
        hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 1));
        hm = max(hm, __shfl_xor_sync(0xffffffffu, hm, 2));
        const int be = (hm >> 20) & 0x7ff;
        const int x = (be == 0 || be == 0x7ff) ? 0 : be - 1023;   // == exponent_of(max)
        const double f = pow2_neg(x);
        cur_e[m] = e_sum[m] + x;
#pragma unroll
        for (int n = 0; n < NT; ++n) { cur[m][n][0] = acc[m][n][0] * f; cur[m][n][1] = acc[m][n][1] * f; }
      } else {
        cur_e[m] = e_sum[m];
#pragma unroll
        for (int n = 0; n < NT; ++n) { cur[m][n][0] = acc[m][n][0]; cur[m][n][1] = acc[m][n][1]; }
      }
      if (dst != nullptr) {
        double* dp = dst + (int64_t)c * S * P + site;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int stt = 8 * n + 2 * t4 + e;
            if (n < NT - 1 || stt < S) __stcg(dp + (int64_t)stt * P, cur[m][n][e]);
          }
        if (t4 == 0) __stcg(dst_scale + (int64_t)c * P + site, cur_e[m]);
      }
      if (is_root) {
        double dot = 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int stt = 8 * n + 2 * t4 + e;
            const double pv = (n < NT - 1 || stt < S) ? __ldg(k.pi + stt) : 0.0;
            dot = fma(pv, cur[m][n][e], dot);
          }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        if (t4 == 0) {
          const int64_t at = ((int64_t)rg.out_index * k.n_cats + c) * P + site;
          __stcg(k.root_dot + at, dot);
          __stcg(k.root_exp + at, e_sum[m]);
        }
      }
    }
  }
}

}  // namespace cb
