// cybayes_b200 -- host scheduler and C ABI (see include/cybayes_b200.h).
//
// One context drives one B200: leaf state codes, a pool of P matrices and a pool of
// partial-likelihood buffers live in HBM; an evaluation is turned into a handful of kernel
// launches (one per tree level, or ONE for a dirty path / a batch of candidate paths) on a
// single stream, followed by an optional scalar NCCL all-reduce and an 8-byte read-back.
// Snapshots are immutable node -> buffer tables with reference-counted buffers: the
// copy-on-write equivalent of the reference's aliased cache dicts (ML_gamma.pyx:114).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "cb_types.cuh"
#include "kernels_dmma.cuh"
#include "kernels_dmma_rc.cuh"
#include "kernels_general.cuh"
#include "kernels_pmat.cuh"
#include "kernels_s2.cuh"

using namespace cb;

// ------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define REQUIRE(cond, ...)            \
  do {                                \
    if (!(cond)) return fail(__VA_ARGS__); \
  } while (0)


// No C++ exception may cross the C ABI: every entry point that allocates runs under this guard.
template <typename F>
static int guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return fail("out of host memory");
  } catch (const std::exception& e) {
    return fail("internal error: %s", e.what());
  } catch (...) {
    return fail("internal error: unknown exception");
  }
}

// ------------------------------------------------------------------------------- NCCL (dlopen)
struct NcclId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  REQUIRE(g_nccl.lib, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (int (*)(NcclId*))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(ncclComm_t*, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.CommDestroy = (int (*)(ncclComm_t))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
  REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy,
          "libnccl is missing expected symbols");
  return 0;
}
enum { NCCL_DOUBLE = 8, NCCL_SUM = 0 };  // ncclFloat64 / ncclSum in nccl.h

// ------------------------------------------------------------------------------ context
constexpr int DMMA_RC_MIN_STATES = 9;   // measured: 1.65x (S = 23, 2304 patterns) .. 3.1x (S = 30, 16384) over the plain FP64 kernel
constexpr int64_t DMMA_RC_MIN_SITES = 1024;  // measured crossover against the 64-site tile kernel + level schedule (S = 47, 64)

struct Buffer {
  double* data = nullptr;
  int32_t* scale = nullptr;
  int refs = 0;
};
struct Snapshot {
  std::vector<int32_t> buf_of_node;     // index by node id; -1 = absent
  std::vector<int32_t> cherry_of_node;  // 2-state family: record of a folded cherry (never materialised), or -1
  int refs = 0;
};
// A folded cherry kept by a snapshot: its two tips and, in the library's own P pool at slots
// [rec * 2C, (rec + 1) * 2C), copies of the P matrices of its two tip edges (tip 0: C slots, tip 1: C slots).
struct CherryRec {
  int32_t tip[2] = {0, 0};
  int refs = 0;
};

struct cb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_stage = nullptr, ev_mark[2] = {nullptr, nullptr};
  int sm_count = 148;
  // alignment
  int n_taxa = 0, n_states = 0, n_cats = 0, code_bytes = 1, n_amb = 0;
  int64_t n_sites = 0, P = 0;  // real and padded pattern counts
  bool family_s2 = false;
  bool use_dmma = false;  // general family, 32 <= S <= 64: FP64 tensor-core kernel (below that the plain kernel wins)
  bool dmma_rc = false;   // ... its register-carried variant (128-site blocks): large alignments, compile-time S
  int rc_stagger = 1;     // anti-lockstep barriers between the warps of one SM sub-partition, see kernels_dmma_rc.cuh
  double* d_staged = nullptr;  // P matrices of the current evaluation in stage layout and consumption order
  size_t staged_bytes = 0;
  int s2_vec = 1;  // sites per thread of the 2-state kernel on large alignments
  int s2_minb = 3; // its __launch_bounds__ min blocks per SM (experiment knob)
  bool s2_stream_stores = true;  // st.global.cs for partials the walk never reads back (10.67 vs 11.0 ms on C4)
  void* d_codes = nullptr;
  double* d_weights = nullptr;
  double* d_amb = nullptr;
  double* d_pi = nullptr;
  // P matrices
  double* d_pmats = nullptr;
  int pmat_cap = 0;
  // partial buffers
  std::vector<Buffer> buffers;
  std::vector<int> free_buffers;
  size_t buffer_bytes = 0;
  std::vector<Snapshot> snaps;
  std::vector<int> free_snaps;
  std::vector<CherryRec> recs;
  std::vector<int> free_recs;
  double* d_pmats_lib = nullptr;
  int lib_cap = 0;  // records the library pool can hold
  // staging
  OpDesc* h_ops = nullptr;
  OpDesc* d_ops = nullptr;
  int ops_cap = 0;
  RangeDesc* h_ranges = nullptr;
  RangeDesc* d_ranges = nullptr;
  int ranges_cap = 0;
  double* d_block_sums = nullptr;
  unsigned* d_tickets = nullptr;
  double* d_results = nullptr;
  double* h_results = nullptr;
  double* d_root_dot = nullptr;
  int32_t* d_root_exp = nullptr;
  int out_cap = 0, max_blocks = 0;
  int last_n_out = 0;
  // scratch for pmat builds
  void* d_scratch = nullptr;
  void* h_scratch = nullptr;
  size_t scratch_cap = 0;
  // L2 flush
  void* d_flush = nullptr;
  size_t flush_bytes = 0;
  // NCCL
  ncclComm_t comm = nullptr;
  int n_ranks = 1;
  // work vectors reused across evals
  std::vector<int32_t> op_of_node, level_of_op, order;
  // stats
  int64_t launches = 0, h2d = 0, d2h = 0, dev_bytes = 0;
  bool timing_valid = false;
};

static int dev_alloc(cb_ctx* c, void** p, size_t bytes) {
  CU(cudaMalloc(p, bytes));
  c->dev_bytes += (int64_t)bytes;
  return 0;
}
static void dev_free(cb_ctx* c, void* p, size_t bytes) {
  if (p) {
    cudaFree(p);
    c->dev_bytes -= (int64_t)bytes;
  }
}

extern "C" const char* cb_last_error(void) { return g_err.c_str(); }
extern "C" int cb_version(void) { return 100; }
extern "C" int cb_device_count(int* out) {
  REQUIRE(out, "null argument");
  CU(cudaGetDeviceCount(out));
  return 0;
}

static int create_impl(int device, cb_ctx** out);
extern "C" int cb_create(int device, cb_ctx** out) {
  return guarded([&] { return create_impl(device, out); });
}
static int create_impl(int device, cb_ctx** out) {
  REQUIRE(out, "null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail("no CUDA device available (%s); cybayes_b200 has no CPU fallback", cudaGetErrorString(e));
  REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  CU(cudaSetDevice(device));
  cb_ctx* c = new cb_ctx();
  c->device = device;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c->ev0));
  CU(cudaEventCreate(&c->ev1));
  CU(cudaEventCreateWithFlags(&c->ev_stage, cudaEventDisableTiming));
  CU(cudaEventCreate(&c->ev_mark[0]));
  CU(cudaEventCreate(&c->ev_mark[1]));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (const char* v = getenv("CYBAYES_S2_V")) c->s2_vec = (atoi(v) == 2) ? 2 : 1;
  if (const char* v = getenv("CYBAYES_S2_MINB")) c->s2_minb = (atoi(v) == 4) ? 4 : 3;
  if (getenv("CYBAYES_S2_NO_CS")) c->s2_stream_stores = false;
  CU(cudaFuncSetAttribute(prune_general_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#define CB_DMMA_ATTR(SS) CU(cudaFuncSetAttribute(prune_dmma_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
  CB_DMMA_ATTR(0); CB_DMMA_ATTR(32); CB_DMMA_ATTR(40); CB_DMMA_ATTR(47); CB_DMMA_ATTR(48); CB_DMMA_ATTR(56); CB_DMMA_ATTR(64);
#undef CB_DMMA_ATTR
#define CB_RC_ATTR(SS)                                                                                                          \
  CU(cudaFuncSetAttribute(prune_dmma_rc_kernel<SS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RcCfg<SS>::SMEM)); \
  CU(cudaFuncSetAttribute(prune_dmma_rc_kernel<SS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RcCfg<SS>::SMEM))
  CB_RC_ATTR(16); CB_RC_ATTR(24); CB_RC_ATTR(32); CB_RC_ATTR(40); CB_RC_ATTR(48); CB_RC_ATTR(56); CB_RC_ATTR(64);
#undef CB_RC_ATTR
  if (const char* v = getenv("CYBAYES_RC_STAGGER")) c->rc_stagger = atoi(v) != 0;
  *out = c;
  return 0;
}

static void free_alignment(cb_ctx* c) {
  for (auto& b : c->buffers) dev_free(c, b.data, c->buffer_bytes);
  c->buffers.clear();
  c->free_buffers.clear();
  c->snaps.clear();
  c->free_snaps.clear();
  c->recs.clear();
  c->free_recs.clear();
  if (c->d_pmats_lib) dev_free(c, c->d_pmats_lib, (size_t)c->lib_cap * 2 * c->n_cats * c->n_states * c->n_states * 8);
  c->d_pmats_lib = nullptr;
  c->lib_cap = 0;
  if (c->d_codes) dev_free(c, c->d_codes, (size_t)c->n_taxa * c->P * c->code_bytes);
  if (c->d_weights) dev_free(c, c->d_weights, (size_t)c->P * 8);
  if (c->d_amb) dev_free(c, c->d_amb, (size_t)std::max(1, c->n_amb) * c->n_states * 8);
  if (c->d_pi) dev_free(c, c->d_pi, (size_t)c->n_states * 8);
  if (c->d_pmats) dev_free(c, c->d_pmats, (size_t)c->pmat_cap * c->n_states * c->n_states * 8);
  if (c->d_staged) dev_free(c, c->d_staged, c->staged_bytes);
  c->d_staged = nullptr;
  c->staged_bytes = 0;
  c->d_codes = nullptr;
  c->d_weights = c->d_amb = c->d_pi = c->d_pmats = nullptr;
  c->pmat_cap = 0;
  if (c->d_block_sums) dev_free(c, c->d_block_sums, (size_t)c->out_cap * c->max_blocks * 8);
  if (c->d_tickets) dev_free(c, c->d_tickets, (size_t)c->out_cap * 4);
  if (c->d_results) dev_free(c, c->d_results, (size_t)c->out_cap * 8);
  if (c->d_root_dot) dev_free(c, c->d_root_dot, (size_t)c->out_cap * c->n_cats * c->P * 8);
  if (c->d_root_exp) dev_free(c, c->d_root_exp, (size_t)c->out_cap * c->n_cats * c->P * 4);
  if (c->h_results) cudaFreeHost(c->h_results);
  c->d_block_sums = c->d_results = c->d_root_dot = nullptr;
  c->d_tickets = nullptr;
  c->d_root_exp = nullptr;
  c->h_results = nullptr;
  c->out_cap = 0;
}

extern "C" int cb_destroy(cb_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  free_alignment(c);
  if (c->d_ops) dev_free(c, c->d_ops, (size_t)c->ops_cap * sizeof(OpDesc));
  if (c->d_ranges) dev_free(c, c->d_ranges, (size_t)c->ranges_cap * sizeof(RangeDesc));
  if (c->h_ops) cudaFreeHost(c->h_ops);
  if (c->h_ranges) cudaFreeHost(c->h_ranges);
  if (c->d_scratch) dev_free(c, c->d_scratch, c->scratch_cap);
  if (c->h_scratch) cudaFreeHost(c->h_scratch);
  if (c->d_flush) dev_free(c, c->d_flush, c->flush_bytes);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->ev_stage);
  cudaEventDestroy(c->ev_mark[0]);
  cudaEventDestroy(c->ev_mark[1]);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

// --------------------------------------------------------------------------------- NCCL
extern "C" int cb_nccl_unique_id(void* id128_out) {
  REQUIRE(id128_out, "null argument");
  if (load_nccl()) return 1;
  NcclId id;
  int r = g_nccl.GetUniqueId(&id);
  REQUIRE(r == 0, "ncclGetUniqueId: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  memcpy(id128_out, &id, sizeof id);
  return 0;
}
extern "C" int cb_comm_init(cb_ctx* c, const void* id128, int rank, int n_ranks) {
  REQUIRE(c && id128, "null argument");
  REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank %d of %d", rank, n_ranks);
  if (load_nccl()) return 1;
  CU(cudaSetDevice(c->device));
  NcclId id;
  memcpy(&id, id128, sizeof id);
  int r = g_nccl.CommInitRank(&c->comm, n_ranks, id, rank);
  REQUIRE(r == 0, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  c->n_ranks = n_ranks;
  return 0;
}

// ---------------------------------------------------------------------------- alignment
static int ensure_scratch(cb_ctx* c, size_t bytes) {
  if (bytes <= c->scratch_cap) return 0;
  size_t cap = std::max(bytes, c->scratch_cap * 2);
  cap = (cap + 4095) & ~(size_t)4095;
  CU(cudaStreamSynchronize(c->stream));
  if (c->d_scratch) dev_free(c, c->d_scratch, c->scratch_cap);
  if (c->h_scratch) cudaFreeHost(c->h_scratch);
  c->d_scratch = c->h_scratch = nullptr;
  c->scratch_cap = 0;
  if (dev_alloc(c, &c->d_scratch, cap)) return 1;
  CU(cudaMallocHost(&c->h_scratch, cap));
  c->scratch_cap = cap;
  return 0;
}

static int set_tips_impl(cb_ctx* c, int n_taxa, int64_t n_sites, int n_states, int n_cats, const void* codes,
                         int code_bytes, const double* amb_sets, int n_amb, const double* weights);
extern "C" int cb_set_tips(cb_ctx* c, int n_taxa, int64_t n_sites, int n_states, int n_cats,
                           const void* codes, int code_bytes, const double* amb_sets, int n_amb,
                           const double* weights) {
  return guarded([&] { return set_tips_impl(c, n_taxa, n_sites, n_states, n_cats, codes, code_bytes, amb_sets, n_amb, weights); });
}
static int set_tips_impl(cb_ctx* c, int n_taxa, int64_t n_sites, int n_states, int n_cats, const void* codes,
                         int code_bytes, const double* amb_sets, int n_amb, const double* weights) {
  REQUIRE(c && codes, "null argument");
  REQUIRE(n_taxa >= 2 && n_sites >= 1 && n_states >= 2, "bad alignment shape %d x %lld x %d", n_taxa,
          (long long)n_sites, n_states);
  REQUIRE(n_cats >= 1 && n_cats <= CB_MAX_CATS, "n_cats must be 1..%d", CB_MAX_CATS);
  REQUIRE(code_bytes == 1 || code_bytes == 2, "code_bytes must be 1 or 2");
  REQUIRE(n_amb >= 1 && amb_sets, "amb_sets must at least hold the all-ones set");
  REQUIRE(n_states + n_amb <= (code_bytes == 1 ? 256 : 65536), "codes do not fit code_bytes");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  free_alignment(c);
  c->n_taxa = n_taxa;
  c->n_sites = n_sites;
  c->n_states = n_states;
  c->n_cats = n_cats;
  c->code_bytes = code_bytes;
  c->n_amb = n_amb;
  c->P = (n_sites + 63) / 64 * 64;
  c->family_s2 = (n_states == 2 && (n_cats == 4 || n_cats == 1));
  c->use_dmma = !c->family_s2 && n_states >= 32 && n_states <= 64 && !getenv("CYBAYES_NO_DMMA");
  {
    // The kernel is fixed per alignment (never per schedule), so every evaluation of an alignment sums in one order.
    // The register-carried variant (always with the walk schedule, cut into parallel subtrees when there are few site
    // tiles) wins from ~1k patterns up: 0.37 vs 0.54 ms at 2048 x 94 taxa, 10.0 vs 18.5 ms at 16384 x 512 taxa (S = 64).
    // CYBAYES_DMMA_RC=1/0 forces it on / off, CYBAYES_RC_MIN_SITES moves the threshold.
    const char* v = getenv("CYBAYES_DMMA_RC");
    const char* m = getenv("CYBAYES_RC_MIN_SITES");
    const char* ms = getenv("CYBAYES_RC_MIN_STATES");
    const bool states_ok = !c->family_s2 && n_states >= (ms ? std::max(9, atoi(ms)) : DMMA_RC_MIN_STATES) && n_states <= 64 &&
                           !getenv("CYBAYES_NO_DMMA");
    c->dmma_rc = states_ok && (v ? atoi(v) != 0 : c->P >= (m ? atoll(m) : DMMA_RC_MIN_SITES));
  }
  const int64_t P = c->P;
  if (dev_alloc(c, &c->d_codes, (size_t)n_taxa * P * code_bytes)) return 1;
  // padding sites carry the all-ones code (a no-op factor) and weight 0
  if (code_bytes == 1) {
    CU(cudaMemsetAsync(c->d_codes, n_states, (size_t)n_taxa * P, c->stream));
  } else {
    std::vector<uint16_t> fill((size_t)P, (uint16_t)n_states);
    for (int t = 0; t < n_taxa; ++t)
      CU(cudaMemcpyAsync((char*)c->d_codes + (size_t)t * P * 2, fill.data(), (size_t)P * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  CU(cudaMemcpy2DAsync(c->d_codes, (size_t)P * code_bytes, codes, (size_t)n_sites * code_bytes,
                       (size_t)n_sites * code_bytes, n_taxa, cudaMemcpyHostToDevice, c->stream));
  c->h2d += (int64_t)n_taxa * n_sites * code_bytes;
  if (dev_alloc(c, (void**)&c->d_weights, (size_t)P * 8)) return 1;
  CU(cudaMemsetAsync(c->d_weights, 0, (size_t)P * 8, c->stream));
  {
    std::vector<double> w((size_t)n_sites, 1.0);
    if (weights) memcpy(w.data(), weights, (size_t)n_sites * 8);
    CU(cudaMemcpyAsync(c->d_weights, w.data(), (size_t)n_sites * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->h2d += n_sites * 8;
  }
  if (dev_alloc(c, (void**)&c->d_amb, (size_t)n_amb * n_states * 8)) return 1;
  CU(cudaMemcpyAsync(c->d_amb, amb_sets, (size_t)n_amb * n_states * 8, cudaMemcpyHostToDevice, c->stream));
  if (dev_alloc(c, (void**)&c->d_pi, (size_t)n_states * 8)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  const size_t scale_ints = c->family_s2 ? (size_t)P : (size_t)n_cats * P;
  c->buffer_bytes = (size_t)n_cats * n_states * P * 8 + scale_ints * 4;
  return 0;
}

// ------------------------------------------------------------------------ P matrices
extern "C" int cb_pmat_reserve(cb_ctx* c, int n_slots) {
  REQUIRE(c && c->n_states > 0, "cb_set_tips must come first");
  if (n_slots <= c->pmat_cap) return 0;
  CU(cudaSetDevice(c->device));
  int cap = std::max(n_slots, std::max(1024, c->pmat_cap * 2));
  const size_t mat = (size_t)c->n_states * c->n_states * 8;
  double* nd = nullptr;
  if (dev_alloc(c, (void**)&nd, (size_t)cap * mat)) return 1;
  if (c->d_pmats) {
    CU(cudaMemcpyAsync(nd, c->d_pmats, (size_t)c->pmat_cap * mat, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dev_free(c, c->d_pmats, (size_t)c->pmat_cap * mat);
  }
  c->d_pmats = nd;
  c->pmat_cap = cap;
  return 0;
}

static int check_slots(cb_ctx* c, int count, const int32_t* slots) {
  REQUIRE(c && slots && count >= 0, "bad argument");
  for (int i = 0; i < count; ++i)
    REQUIRE(slots[i] >= 0 && slots[i] < c->pmat_cap, "P slot %d out of range (reserved %d)", slots[i], c->pmat_cap);
  return 0;
}

extern "C" int cb_pmat_upload(cb_ctx* c, int count, const int32_t* slots, const double* mats) {
  if (check_slots(c, count, slots)) return 1;
  REQUIRE(mats, "null argument");
  CU(cudaSetDevice(c->device));
  const size_t mat = (size_t)c->n_states * c->n_states * 8;
  if (ensure_scratch(c, (size_t)count * mat)) return 1;
  CU(cudaStreamSynchronize(c->stream));  // scratch reuse
  memcpy(c->h_scratch, mats, (size_t)count * mat);
  int i = 0;
  while (i < count) {  // coalesce runs of consecutive slots into one copy
    int j = i + 1;
    while (j < count && slots[j] == slots[j - 1] + 1) ++j;
    CU(cudaMemcpyAsync(c->d_pmats + (size_t)slots[i] * c->n_states * c->n_states, (char*)c->h_scratch + (size_t)i * mat,
                       (size_t)(j - i) * mat, cudaMemcpyHostToDevice, c->stream));
    i = j;
  }
  c->h2d += (int64_t)count * mat;
  return 0;
}

extern "C" int cb_pmat_download(cb_ctx* c, int count, const int32_t* slots, double* out) {
  if (check_slots(c, count, slots)) return 1;
  REQUIRE(out, "null argument");
  CU(cudaSetDevice(c->device));
  const size_t mat = (size_t)c->n_states * c->n_states * 8;
  for (int i = 0; i < count; ++i)
    CU(cudaMemcpyAsync((char*)out + (size_t)i * mat, c->d_pmats + (size_t)slots[i] * c->n_states * c->n_states, mat,
                       cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)count * mat;
  return 0;
}

extern "C" int cb_pmat_build(cb_ctx* c, int model, const double* pi, double beta, const double* gtr,
                             int count, const int32_t* slots, const double* d, const double* x) {
  if (check_slots(c, count, slots)) return 1;
  REQUIRE(d, "null argument");
  REQUIRE(model >= CB_MODEL_JC && model <= CB_MODEL_GTR_EIG, "unknown model %d", model);
  REQUIRE(model == CB_MODEL_JC || pi, "pi required");
  REQUIRE(model != CB_MODEL_GTR_EIG || gtr, "GTR eigensystem required");
  REQUIRE(model != CB_MODEL_F81_BINARY || c->n_states == 2, "binary F81 needs 2 states");
  if (count == 0) return 0;
  CU(cudaSetDevice(c->device));
  const int S = c->n_states;
  const size_t n_gtr = (model == CB_MODEL_GTR_EIG) ? (size_t)S + 2 * (size_t)S * S : 0;
  // scratch layout: pi[S] | gtr[n_gtr] | d[count] | x[count] | slots[count] (int32)
  const size_t doubles = (size_t)S + n_gtr + 2 * (size_t)count;
  const size_t bytes = doubles * 8 + (size_t)count * 4;
  if (ensure_scratch(c, bytes)) return 1;
  CU(cudaStreamSynchronize(c->stream));  // scratch reuse
  double* h = (double*)c->h_scratch;
  if (pi) memcpy(h, pi, (size_t)S * 8); else memset(h, 0, (size_t)S * 8);
  if (n_gtr) memcpy(h + S, gtr, n_gtr * 8);
  memcpy(h + S + n_gtr, d, (size_t)count * 8);
  if (x) memcpy(h + S + n_gtr + count, x, (size_t)count * 8);
  memcpy(h + doubles, slots, (size_t)count * 4);
  CU(cudaMemcpyAsync(c->d_scratch, c->h_scratch, bytes, cudaMemcpyHostToDevice, c->stream));
  c->h2d += (int64_t)bytes;
  const double* dd = (const double*)c->d_scratch;
  const int threads = (S * S >= 256) ? 256 : ((S * S + 31) / 32 * 32);
  pmat_build_kernel<<<count, threads, (size_t)S * 8, c->stream>>>(
      model, S, dd, beta, n_gtr ? dd + S : nullptr, count, (const int32_t*)(dd + doubles), dd + S + n_gtr,
      x ? dd + S + n_gtr + count : nullptr, c->d_pmats);
  CU(cudaGetLastError());
  c->launches += 1;
  return 0;
}

// ------------------------------------------------------------------------ buffers/snapshots
static int buffer_acquire(cb_ctx* c, int* out) {
  if (!c->free_buffers.empty()) {
    *out = c->free_buffers.back();
    c->free_buffers.pop_back();
  } else {
    Buffer b;
    void* p = nullptr;
    if (dev_alloc(c, &p, c->buffer_bytes)) return 1;
    b.data = (double*)p;
    b.scale = (int32_t*)(b.data + (size_t)c->n_cats * c->n_states * c->P);
    c->buffers.push_back(b);
    *out = (int)c->buffers.size() - 1;
  }
  c->buffers[*out].refs = 1;
  return 0;
}
static void buffer_release(cb_ctx* c, int b) {
  if (b < 0) return;
  if (--c->buffers[b].refs == 0) c->free_buffers.push_back(b);
}
static int rec_acquire(cb_ctx* c, int tip0, int tip1, int* out) {
  if (!c->free_recs.empty()) {
    *out = c->free_recs.back();
    c->free_recs.pop_back();
  } else {
    c->recs.emplace_back();
    *out = (int)c->recs.size() - 1;
  }
  c->recs[*out].tip[0] = tip0;
  c->recs[*out].tip[1] = tip1;
  c->recs[*out].refs = 1;
  return 0;
}
static void rec_release(cb_ctx* c, int r) {
  if (r < 0) return;
  if (--c->recs[r].refs == 0) c->free_recs.push_back(r);
}
// the library's P pool must hold every record that exists (old contents are kept when it grows)
static int ensure_lib_pool(cb_ctx* c) {
  const int need = (int)c->recs.size();
  if (need <= c->lib_cap) return 0;
  const int cap = std::max(need, std::max(1024, c->lib_cap * 2));
  const size_t per = (size_t)2 * c->n_cats * c->n_states * c->n_states * 8;
  double* nd = nullptr;
  if (dev_alloc(c, (void**)&nd, (size_t)cap * per)) return 1;
  if (c->d_pmats_lib) {
    CU(cudaMemcpyAsync(nd, c->d_pmats_lib, (size_t)c->lib_cap * per, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    dev_free(c, c->d_pmats_lib, (size_t)c->lib_cap * per);
  }
  c->d_pmats_lib = nd;
  c->lib_cap = cap;
  return 0;
}
static bool snapshot_valid(cb_ctx* c, int s) {
  return s >= 0 && s < (int)c->snaps.size() && c->snaps[s].refs > 0;
}
extern "C" int cb_snapshot_retain(cb_ctx* c, int s) {
  REQUIRE(c && snapshot_valid(c, s), "invalid snapshot %d", s);
  c->snaps[s].refs++;
  return 0;
}
extern "C" int cb_snapshot_release(cb_ctx* c, int s) {
  REQUIRE(c && snapshot_valid(c, s), "invalid snapshot %d", s);
  if (--c->snaps[s].refs == 0) {
    for (int32_t b : c->snaps[s].buf_of_node) buffer_release(c, b);
    for (int32_t r : c->snaps[s].cherry_of_node) rec_release(c, r);
    c->snaps[s].buf_of_node.clear();
    c->snaps[s].cherry_of_node.clear();
    c->free_snaps.push_back(s);
  }
  return 0;
}

static int snapshot_read_impl(cb_ctx* c, int s, int node, double* out, int32_t* scale_out);
extern "C" int cb_snapshot_read(cb_ctx* c, int s, int node, double* out, int32_t* scale_out) {
  return guarded([&] { return snapshot_read_impl(c, s, node, out, scale_out); });
}
static int materialize_cherry(cb_ctx* c, int rec, int* buf_out);
static int snapshot_read_impl(cb_ctx* c, int s, int node, double* out, int32_t* scale_out) {
  REQUIRE(c && out && snapshot_valid(c, s), "invalid snapshot %d", s);
  const Snapshot& sn = c->snaps[s];
  REQUIRE(node >= 0 && node < (int)sn.buf_of_node.size(), "node %d is not in snapshot %d", node, s);
  CU(cudaSetDevice(c->device));
  int bidx = sn.buf_of_node[node], tmp_buf = -1;
  if (bidx < 0 && node < (int)sn.cherry_of_node.size() && sn.cherry_of_node[node] >= 0) {
    // a folded cherry has no stored partial: compute it now, from the P copies the snapshot keeps
    if (materialize_cherry(c, sn.cherry_of_node[node], &tmp_buf)) return 1;
    bidx = tmp_buf;
  }
  REQUIRE(bidx >= 0, "node %d is not in snapshot %d", node, s);
  const Buffer& b = c->buffers[bidx];
  const int C = c->n_cats, S = c->n_states;
  const int64_t P = c->P, n = c->n_sites;
  std::vector<double> tmp((size_t)C * S * P);
  const size_t n_scale = c->family_s2 ? (size_t)P : (size_t)C * P;
  std::vector<int32_t> sc(n_scale);
  CU(cudaMemcpyAsync(tmp.data(), b.data, tmp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(sc.data(), b.scale, n_scale * 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->d2h += (int64_t)tmp.size() * 8 + (int64_t)n_scale * 4;
  for (int k = 0; k < C; ++k)
    for (int i = 0; i < S; ++i)
      for (int64_t p = 0; p < n; ++p) {
        const int e = c->family_s2 ? sc[p] : sc[(size_t)k * P + p];
        double v = tmp[((size_t)k * S + i) * P + p];
        if (!scale_out) v = ldexp(v, e);
        out[((size_t)k * S + i) * n + p] = v;
      }
  if (scale_out) {
    // one exponent per site: categories of the general family are brought to their max
    for (int64_t p = 0; p < n; ++p) {
      if (c->family_s2) {
        scale_out[p] = sc[p];
      } else {
        int emax = INT_MIN;
        for (int k = 0; k < C; ++k) emax = std::max(emax, sc[(size_t)k * P + p]);
        scale_out[p] = emax;
        for (int k = 0; k < C; ++k)
          for (int i = 0; i < S; ++i) {
            double& v = out[((size_t)k * S + i) * n + p];
            v = ldexp(v, sc[(size_t)k * P + p] - emax);
          }
      }
    }
  }
  if (tmp_buf >= 0) buffer_release(c, tmp_buf);
  return 0;
}

// ---------------------------------------------------------------------------- evaluation
static int ensure_staging(cb_ctx* c, int n_ops, int n_ranges, int n_out) {
  if (n_ops > c->ops_cap) {
    int cap = std::max(n_ops, std::max(256, c->ops_cap * 2));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_ops) dev_free(c, c->d_ops, (size_t)c->ops_cap * sizeof(OpDesc));
    if (c->h_ops) cudaFreeHost(c->h_ops);
    c->d_ops = nullptr; c->h_ops = nullptr; c->ops_cap = 0;
    if (dev_alloc(c, (void**)&c->d_ops, (size_t)cap * sizeof(OpDesc))) return 1;
    CU(cudaMallocHost(&c->h_ops, (size_t)cap * sizeof(OpDesc)));
    c->ops_cap = cap;
  }
  if (n_ranges > c->ranges_cap) {
    int cap = std::max(n_ranges, std::max(256, c->ranges_cap * 2));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_ranges) dev_free(c, c->d_ranges, (size_t)c->ranges_cap * sizeof(RangeDesc));
    if (c->h_ranges) cudaFreeHost(c->h_ranges);
    c->d_ranges = nullptr; c->h_ranges = nullptr; c->ranges_cap = 0;
    if (dev_alloc(c, (void**)&c->d_ranges, (size_t)cap * sizeof(RangeDesc))) return 1;
    CU(cudaMallocHost(&c->h_ranges, (size_t)cap * sizeof(RangeDesc)));
    c->ranges_cap = cap;
  }
  if (n_out > c->out_cap) {
    int cap = std::max(n_out, std::max(1, c->out_cap * 2));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_block_sums) dev_free(c, c->d_block_sums, (size_t)c->out_cap * c->max_blocks * 8);
    if (c->d_tickets) dev_free(c, c->d_tickets, (size_t)c->out_cap * 4);
    if (c->d_results) dev_free(c, c->d_results, (size_t)c->out_cap * 8);
    if (c->d_root_dot) dev_free(c, c->d_root_dot, (size_t)c->out_cap * c->n_cats * c->P * 8);
    if (c->d_root_exp) dev_free(c, c->d_root_exp, (size_t)c->out_cap * c->n_cats * c->P * 4);
    if (c->h_results) cudaFreeHost(c->h_results);
    c->d_block_sums = c->d_results = c->d_root_dot = nullptr;
    c->d_tickets = nullptr; c->d_root_exp = nullptr; c->h_results = nullptr; c->out_cap = 0;
    c->max_blocks = (int)(c->P / 32) + 1;
    if (dev_alloc(c, (void**)&c->d_block_sums, (size_t)cap * c->max_blocks * 8)) return 1;
    if (dev_alloc(c, (void**)&c->d_tickets, (size_t)cap * 4)) return 1;
    CU(cudaMemsetAsync(c->d_tickets, 0, (size_t)cap * 4, c->stream));
    if (dev_alloc(c, (void**)&c->d_results, (size_t)cap * 8)) return 1;
    if (!c->family_s2) {
      if (dev_alloc(c, (void**)&c->d_root_dot, (size_t)cap * c->n_cats * c->P * 8)) return 1;
      if (dev_alloc(c, (void**)&c->d_root_exp, (size_t)cap * c->n_cats * c->P * 4)) return 1;
    }
    CU(cudaMallocHost(&c->h_results, (size_t)cap * 8));
    c->out_cap = cap;
  }
  return 0;
}

static LaunchConst make_const(cb_ctx* c) {
  LaunchConst k;
  k.ops = c->d_ops;
  k.ranges = c->d_ranges;
  k.pmats = c->d_pmats;
  k.pmats_lib = c->d_pmats_lib;
  k.staged = c->d_staged;
  k.weights = c->d_weights;
  k.pi = c->d_pi;
  k.amb = c->d_amb;
  k.block_sums = c->d_block_sums;
  k.tickets = c->d_tickets;
  k.results = c->d_results;
  k.root_dot = c->d_root_dot;
  k.root_exp = c->d_root_exp;
  k.n_sites = c->P;
  k.n_states = c->n_states;
  k.n_cats = c->n_cats;
  k.code_bytes = c->code_bytes;
  k.max_blocks = c->max_blocks;
  k.cats = (double)c->n_cats;
  k.rc_stagger = c->rc_stagger;
  k.n_amb = c->n_amb;
  return k;
}

constexpr int MAX_CHAIN_OPS = 4096;          // longer dirty paths fall back to the level schedule
constexpr int64_t WALK_MIN_SITES_S2 = 32768; // below this a site tile cannot fill the GPU: use levels
constexpr int64_t WALK_MIN_SITES_GENERAL = 32768;

static int general_rows_per_chunk(int S) {
  // rows of the two P matrices staged per pass (multiple of 4, <= 64).  Prefer a footprint
  // <= 100 KB (two blocks per SM) as long as that keeps >= 16 rows; else use up to 226 KB.
  const int rmax = std::min((S + 3) / 4 * 4, 64);
  for (int R = rmax; R >= std::min(16, rmax); R -= 4)
    if (gen_smem_bytes(S, R) <= (size_t)100 * 1024) return R;
  for (int R = rmax; R >= 4; R -= 4)
    if (gen_smem_bytes(S, R) <= (size_t)226 * 1024) return R;
  return 0;
}

// Launch the ranges [r_begin, r_end) (all independent) as one kernel.
static int launch_ranges(cb_ctx* c, const LaunchConst& k, int r_begin, int r_end, int max_ops_in_range) {
  const int n_r = r_end - r_begin;
  LaunchConst kk = k;
  kk.ranges = k.ranges + r_begin;
  if (c->family_s2) {
    // Fixed per alignment (independent of the schedule) so the reduction order never changes:
    // big alignments: 256 threads x V sites (V from CYBAYES_S2_V, default 1 = more resident warps);
    // small alignments: 64-thread blocks, one site per thread, to spread over the SMs.
    const size_t smem = s2_smem_bytes(max_ops_in_range, c->n_cats);
    const bool small = c->P < (int64_t)64 * 2 * c->sm_count * 4;
    const int V = small ? 1 : c->s2_vec;
    const int threads = small ? 64 : 256;
    dim3 grid((unsigned)((c->P + (int64_t)threads * V - 1) / ((int64_t)threads * V)), (unsigned)n_r);
#define CB_LAUNCH_S2(CC, VV, TT, MB) prune_s2_kernel<CC, VV, TT, MB><<<grid, TT, smem, c->stream>>>(kk)
    if (c->n_cats == 4) {
      if (small) CB_LAUNCH_S2(4, 1, 64, 8);
      else if (V == 1 && c->s2_minb == 4) CB_LAUNCH_S2(4, 1, 256, 4);
      else if (V == 1) CB_LAUNCH_S2(4, 1, 256, 3);
      else CB_LAUNCH_S2(4, 2, 256, 2);
    } else {
      if (small) CB_LAUNCH_S2(1, 1, 64, 8); else if (V == 1) CB_LAUNCH_S2(1, 1, 256, 4); else CB_LAUNCH_S2(1, 2, 256, 4);
    }
#undef CB_LAUNCH_S2
  } else if (c->dmma_rc) {
    dim3 grid((unsigned)((c->P + RC_T - 1) / RC_T), (unsigned)n_r, (unsigned)c->n_cats);
    switch ((c->n_states + 7) / 8 * 8) {  // compiled per padded state count; the real one is a run-time value
#define CB_RC_CASE(SS)                                                                                          \
  case SS:                                                                                                      \
    if (c->n_states == SS) prune_dmma_rc_kernel<SS, true><<<grid, RC_THREADS, RcCfg<SS>::SMEM, c->stream>>>(kk); \
    else prune_dmma_rc_kernel<SS, false><<<grid, RC_THREADS, RcCfg<SS>::SMEM, c->stream>>>(kk);                  \
    break
      CB_RC_CASE(16); CB_RC_CASE(24); CB_RC_CASE(32); CB_RC_CASE(40); CB_RC_CASE(48); CB_RC_CASE(56); CB_RC_CASE(64);
#undef CB_RC_CASE
      default: return fail("internal error: no register-carried DMMA kernel for %d states", c->n_states);
    }
  } else if (c->use_dmma) {
    dim3 grid((unsigned)(c->P / DM_T), (unsigned)n_r, (unsigned)c->n_cats);
    const size_t smem = dm_smem_bytes(c->n_states);
    switch (c->n_states) {  // compile-time state counts for the common sizes, run-time S otherwise
#define CB_DMMA_CASE(SS) case SS: prune_dmma_kernel<SS><<<grid, DM_THREADS, smem, c->stream>>>(kk); break
      CB_DMMA_CASE(32); CB_DMMA_CASE(40); CB_DMMA_CASE(47); CB_DMMA_CASE(48); CB_DMMA_CASE(56); CB_DMMA_CASE(64);
#undef CB_DMMA_CASE
      default: prune_dmma_kernel<0><<<grid, DM_THREADS, smem, c->stream>>>(kk);
    }
  } else {
    const int R = general_rows_per_chunk(c->n_states);
    REQUIRE(R > 0, "n_states = %d does not fit the shared-memory tiling", c->n_states);
    dim3 grid((unsigned)(c->P / GEN_T), (unsigned)n_r, (unsigned)c->n_cats);
    prune_general_kernel<<<grid, GEN_THREADS, gen_smem_bytes(c->n_states, R), c->stream>>>(kk, R);
  }
  CU(cudaGetLastError());
  c->launches += 1;
  return 0;
}

// Compute the partial of a folded cherry into a fresh buffer (debug / read-back path): one ordinary op
// with two tip children whose P matrices come from the library's pool.
static int materialize_cherry(cb_ctx* c, int rec, int* buf_out) {
  REQUIRE(c->family_s2 && rec >= 0 && rec < (int)c->recs.size(), "bad cherry record %d", rec);
  if (ensure_staging(c, 1, 1, 1)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  int bi;
  if (buffer_acquire(c, &bi)) return 1;
  OpDesc& op = c->h_ops[0];
  memset(&op, 0, sizeof op);
  op.dst = c->buffers[bi].data;
  op.dst_scale = c->buffers[bi].scale;
  for (int kx = 0; kx < 2; ++kx) {
    op.kind[kx] = SRC_TIP;
    op.src[kx] = (const char*)c->d_codes + (size_t)(c->recs[rec].tip[kx] - 1) * c->P * c->code_bytes;
    for (int q = 0; q < c->n_cats; ++q) op.pslot[kx][q] = CB_LIB_SLOT | (rec * 2 * c->n_cats + kx * c->n_cats + q);
    op.crec_out[kx] = -1;
  }
  c->h_ranges[0] = RangeDesc{0, 1, -1, 0};
  CU(cudaMemcpyAsync(c->d_ops, c->h_ops, sizeof(OpDesc), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_ranges, c->h_ranges, sizeof(RangeDesc), cudaMemcpyHostToDevice, c->stream));
  const LaunchConst k = make_const(c);
  if (launch_ranges(c, k, 0, 1, 1)) return 1;
  CU(cudaStreamSynchronize(c->stream));
  *buf_out = bi;
  return 0;
}

// Shared implementation of cb_eval (n_lists = 1) and cb_eval_batch.
//
// Schedules (all run the same per-node arithmetic, so results are bit-identical):
//   CHAIN  a dirty path: one launch, one block range walks the ops, the on-path partial is
//          carried in registers / shared memory                        (ML_gamma.pyx:99-114)
//   WALK   a whole (sub)tree on a large alignment: ONE launch; every block walks all ops for
//          its site tile in depth-first order, heavier subtree first.  The child finished
//          last is carried on chip, so only nodes with two internal children are ever read
//          back (about a third of them), and those reads come from this block's own recent
//          writes (L2) when the lighter sibling subtree is small.
//   LEVELS one launch per tree level, grid.y = nodes of the level: the latency schedule for
//          small alignments where a site tile alone cannot fill the GPU.
enum Schedule { SCHED_CHAIN = 0, SCHED_WALK = 1, SCHED_LEVELS = 2 };

static int eval_impl(cb_ctx* c, int snapshot_in, int n_lists, const int32_t* offsets_in, const int32_t* nodes_in,
                     const int32_t* children_in, const int32_t* pslots_in, const double* pi, int flags,
                     int* snapshot_out, double* lnl_out) {
  REQUIRE(c && c->n_states > 0, "cb_set_tips must come first");
  REQUIRE(offsets_in && nodes_in && children_in && pslots_in && pi, "null argument");
  REQUIRE(snapshot_in < 0 || snapshot_valid(c, snapshot_in), "invalid snapshot %d", snapshot_in);
  const int C = c->n_cats, N = c->n_taxa, n_nodes = 2 * N;  // ids 1 .. 2N-1
  REQUIRE(offsets_in[n_lists] > 0, "empty op list");

  // ---- cherry folding (2-state family) ---------------------------------------------------------
  // An op whose two children are tips is not run: its parent (which is always in the list) takes it
  // as a SRC_CHERRY child and looks the cherry's 3 x 3 possible partials up.  The compacted op list
  // replaces the caller's; fold_child remembers, per (compacted op, child), the caller's op index
  // of the folded cherry.
  std::vector<int32_t> offsets_v(offsets_in, offsets_in + n_lists + 1), nodes_v, children_v, pslots_v, fold_child;
  const bool fold = c->family_s2 && !(flags & CB_EVAL_NO_FOLD);
  if (fold) {
    std::vector<int32_t> parent_op(n_nodes);
    for (int li = 0; li < n_lists; ++li) {
      const int b = offsets_in[li], e = offsets_in[li + 1];
      std::fill(parent_op.begin(), parent_op.end(), -1);
      for (int i = b; i < e; ++i)
        for (int kx = 0; kx < 2; ++kx) {
          const int ch = children_in[2 * i + kx];
          if (ch > N && ch < n_nodes) parent_op[ch] = i;
        }
      offsets_v[li] = (int32_t)nodes_v.size();
      std::vector<int32_t> folded_at(n_nodes, -1);  // node -> caller's op index of its folded cherry op
      for (int i = b; i < e; ++i) {
        const int node = nodes_in[i], c0 = children_in[2 * i], c1 = children_in[2 * i + 1];
        const bool cherry = c0 >= 1 && c0 <= N && c1 >= 1 && c1 <= N && i != e - 1 && node > N && node < n_nodes &&
                            parent_op[node] > i;
        if (cherry) {
          folded_at[node] = i;
          continue;
        }
        nodes_v.push_back(node);
        for (int kx = 0; kx < 2; ++kx) {
          const int ch = children_in[2 * i + kx];
          children_v.push_back(ch);
          fold_child.push_back((ch > N && ch < n_nodes) ? folded_at[ch] : -1);
          for (int q = 0; q < C; ++q) pslots_v.push_back(pslots_in[(size_t)(2 * i + kx) * C + q]);
        }
      }
    }
    offsets_v[n_lists] = (int32_t)nodes_v.size();
  }
  const int32_t* offsets = fold ? offsets_v.data() : offsets_in;
  const int32_t* nodes = fold ? nodes_v.data() : nodes_in;
  const int32_t* children = fold ? children_v.data() : children_in;
  const int32_t* pslots = fold ? pslots_v.data() : pslots_in;
  const int total_ops = offsets[n_lists];
  REQUIRE(total_ops > 0, "empty op list");
  std::vector<int> new_cherry_nodes, new_cherry_recs;  // folded cherries that stay in the returned snapshot
  const bool want_snap = (flags & CB_EVAL_WANT_SNAPSHOT) != 0;
  const bool store_root = (flags & CB_EVAL_STORE_ROOT) != 0;
  REQUIRE(!(want_snap && n_lists != 1), "snapshots are only kept for single evaluations");
  CU(cudaSetDevice(c->device));
  if (ensure_staging(c, total_ops, total_ops + n_lists, n_lists)) return 1;
  CU(cudaEventSynchronize(c->ev_stage));  // previous H2D of the staging area finished

  const Snapshot* sin = snapshot_in >= 0 ? &c->snaps[snapshot_in] : nullptr;
  c->op_of_node.assign(n_nodes, -1);
  std::vector<int> new_bufs, new_nodes;  // buffers written by this evaluation, and their nodes
  std::vector<int> tmp_bufs;             // temporaries of an evaluation that keeps no snapshot
  int n_ranges = 0;
  std::vector<std::pair<int, int>> launches;  // [range begin, range end) per launch
  std::vector<int> launch_maxops;
  bool single_launch = true;
  std::vector<int>& order = c->order;    // position -> original op index (per list)
  std::vector<int> pos_of, n_sub, kid_op[2], range_id;
  std::vector<std::pair<int, int>> walk_segs;
  std::vector<char> read_back;

  for (int li = 0; li < n_lists; ++li) {
    const int b = offsets[li], e = offsets[li + 1], n = e - b;
    REQUIRE(n > 0, "empty op list %d", li);
    for (int i = b; i < e; ++i) {
      const int node = nodes[i];
      REQUIRE(node > N && node < n_nodes, "op %d: node %d is not an internal node", i, node);
      REQUIRE(c->op_of_node[node] < 0, "op %d: node %d is computed twice", i, node);
      c->op_of_node[node] = i;
    }
    // producers of the children inside this list (-1: tip or snapshot)
    kid_op[0].assign(n, -1);
    kid_op[1].assign(n, -1);
    for (int i = b; i < e; ++i)
      for (int kx = 0; kx < 2; ++kx) {
        const int ch = children[2 * i + kx];
        REQUIRE((ch >= 1 && ch <= N) || (ch > N && ch < n_nodes), "op %d: bad child id %d", i, ch);
        if (ch > N && c->op_of_node[ch] >= b) {
          REQUIRE(c->op_of_node[ch] < i, "op %d: child %d is computed after its parent", i, ch);
          kid_op[kx][i - b] = c->op_of_node[ch] - b;
        }
      }
    // chain: every op after the first consumes exactly the previous op and nothing else of the list
    bool chain = !(flags & CB_EVAL_FORCE_LEVELS) && n <= MAX_CHAIN_OPS;
    for (int i = 0; i < n && chain; ++i) {
      const int a0 = kid_op[0][i], a1 = kid_op[1][i];
      if (i == 0) chain = (a0 < 0 && a1 < 0);
      else chain = (a0 == i - 1) != (a1 == i - 1) && (a0 < 0 || a0 == i - 1) && (a1 < 0 || a1 == i - 1);
    }
    REQUIRE(chain || n_lists == 1, "cb_eval_batch: candidate %d is not a chain of at most %d ops", li, MAX_CHAIN_OPS);
    Schedule sched = SCHED_CHAIN;
    if (!chain) {
      // the register-carried kernel only pays off when the partial is carried: its alignments always walk
      const bool big = c->dmma_rc || c->P >= (c->family_s2 ? WALK_MIN_SITES_S2 : WALK_MIN_SITES_GENERAL);
      sched = ((big || (flags & CB_EVAL_FORCE_WALK)) && !(flags & CB_EVAL_FORCE_LEVELS)) ? SCHED_WALK : SCHED_LEVELS;
    }

    order.resize(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    if (sched == SCHED_WALK) {
      // subtree sizes in ops (children precede parents in the caller's order)
      n_sub.assign(n, 1);
      for (int i = 0; i < n; ++i)
        for (int kx = 0; kx < 2; ++kx)
          if (kid_op[kx][i] >= 0) n_sub[i] += n_sub[kid_op[kx][i]];
      // Post-order from the root op, heavier child subtree first.  When the alignment gives too few
      // site tiles to fill the GPU for a whole sequential walk (small shards), the tree is cut into
      // independent subtrees of at most `limit` ops that run as parallel ranges of a first launch;
      // the ops above the cut follow in a second launch.
      int limit = n + 1;
      {
        const int64_t tile_sites = c->family_s2 ? (int64_t)256 * c->s2_vec : (c->dmma_rc ? RC_T : (c->use_dmma ? DM_T : GEN_T));
        const int64_t blocks = (c->P + tile_sites - 1) / tile_sites * (c->family_s2 ? 1 : c->n_cats);
        const int64_t slots = (int64_t)c->sm_count * (c->family_s2 ? (c->s2_vec == 1 ? 3 : 2) : (c->dmma_rc ? 1 : 2));
        if (blocks < 4 * slots && n >= 64 && !(flags & CB_EVAL_FORCE_WALK))
          limit = std::max(16, std::min(256, (int)(n * blocks / (8 * slots)) + 1));
        if (getenv("CYBAYES_WALK_SPLIT")) limit = std::max(2, atoi(getenv("CYBAYES_WALK_SPLIT")));
        if (limit >= n) limit = n + 1;  // nothing to cut
      }
      std::vector<int> out, top, stack;
      std::vector<char> expanded(n, 0);
      std::vector<std::pair<int, int>> segs;
      out.reserve(n);
      auto post_order = [&](int root_op) {
        stack.clear();
        stack.push_back(root_op);
        while (!stack.empty()) {
          const int i = stack.back();
          if (expanded[i]) {
            stack.pop_back();
            out.push_back(i);
            continue;
          }
          expanded[i] = 1;
          int a0 = kid_op[0][i], a1 = kid_op[1][i];
          if (a0 >= 0 && a1 >= 0 && n_sub[a1] > n_sub[a0]) std::swap(a0, a1);  // a0 = heavier
          if (a1 >= 0) stack.push_back(a1);  // lighter: visited second (pushed first)
          if (a0 >= 0) stack.push_back(a0);
        }
      };
      if (limit > n) {
        post_order(n - 1);
      } else {
        // walk down from the root: ops with a subtree above the limit stay in `top`
        std::vector<std::pair<int, int>> st2;  // (op, phase)
        std::vector<int> cut;
        st2.push_back({n - 1, 0});
        while (!st2.empty()) {
          auto [i, phase] = st2.back();
          st2.pop_back();
          if (phase == 1) { top.push_back(i); continue; }
          if (n_sub[i] <= limit) { cut.push_back(i); continue; }
          st2.push_back({i, 1});
          int a0 = kid_op[0][i], a1 = kid_op[1][i];
          if (a0 >= 0 && a1 >= 0 && n_sub[a1] > n_sub[a0]) std::swap(a0, a1);
          if (a1 >= 0) st2.push_back({a1, 0});
          if (a0 >= 0) st2.push_back({a0, 0});
        }
        std::stable_sort(cut.begin(), cut.end(), [&](int x, int y) { return n_sub[x] > n_sub[y]; });  // big first
        for (int r : cut) {
          const int b0 = (int)out.size();
          post_order(r);
          segs.push_back({b0, (int)out.size()});
        }
        for (int i : top) out.push_back(i);
      }
      if ((int)out.size() == n) {
        order = out;
        range_id.assign(n, 0);
        for (size_t si = 0; si < segs.size(); ++si)
          for (int p = segs[si].first; p < segs[si].second; ++p) range_id[p] = (int)si + 1;
        walk_segs = segs;
      } else {
        sched = SCHED_LEVELS;  // ops not under the root
      }
    }
    if (sched != SCHED_WALK) { range_id.assign(n, 0); walk_segs.clear(); }
    pos_of.assign(n, 0);
    for (int p = 0; p < n; ++p) pos_of[order[p]] = p;
    // which ops are read back from memory by a later op (as opposed to carried on chip)?
    read_back.assign(n, 0);
    for (int i = 0; i < n; ++i)
      for (int kx = 0; kx < 2; ++kx) {
        const int k0 = kid_op[kx][i];
        if (k0 >= 0 && !(sched != SCHED_LEVELS && pos_of[k0] == pos_of[i] - 1 && range_id[pos_of[k0]] == range_id[pos_of[i]]))
          read_back[k0] = 1;
      }

    for (int p = 0; p < n; ++p) {
      const int i0 = order[p], i = b + i0;
      OpDesc& op = c->h_ops[b + p];
      const bool is_root = (i0 == n - 1);
      REQUIRE(!is_root || p == n - 1, "internal error: root is not last");
      op.is_root = is_root ? 1 : 0;
      op.pad_ = (read_back[i0] || !c->s2_stream_stores) ? 0 : 1;  // 1: nobody reads this partial back in this evaluation -> streaming stores
      op.dst = nullptr;
      op.dst_scale = nullptr;
      const bool keep = is_root ? (want_snap && store_root) : (want_snap || read_back[i0]);
      if (keep) {
        int bi;
        if (buffer_acquire(c, &bi)) return 1;
        op.dst = c->buffers[bi].data;
        op.dst_scale = c->buffers[bi].scale;
        if (want_snap) {
          new_bufs.push_back(bi);
          new_nodes.push_back(nodes[i]);
        } else {
          tmp_bufs.push_back(bi);
        }
      }
      for (int kx = 0; kx < 2; ++kx) {
        const int ch = children[2 * i + kx];
        op.src[kx] = nullptr;
        op.src_scale[kx] = nullptr;
        op.ctip[kx][0] = op.ctip[kx][1] = nullptr;
        op.crec_out[kx] = -1;
        for (int t = 0; t < 2; ++t)
          for (int q = 0; q < CB_S2_MAX_CATS; ++q) op.cslot[kx][t][q] = 0;
        const int folded = fold ? fold_child[(size_t)2 * i + kx] : -1;
        const int snap_rec = (fold && folded < 0 && ch > N && kid_op[kx][i0] < 0 && sin &&
                              ch < (int)sin->cherry_of_node.size()) ? sin->cherry_of_node[ch] : -1;
        if (ch <= N) {
          op.kind[kx] = SRC_TIP;
          op.src[kx] = (const char*)c->d_codes + (size_t)(ch - 1) * c->P * c->code_bytes;
        } else if (folded >= 0) {
          // the cherry was in the caller's list: its P matrices are the caller's slots
          op.kind[kx] = SRC_CHERRY;
          for (int t = 0; t < 2; ++t) {
            const int tip = children_in[2 * folded + t];
            op.ctip[kx][t] = (const char*)c->d_codes + (size_t)(tip - 1) * c->P * c->code_bytes;
            for (int q = 0; q < C; ++q) {
              const int sl = pslots_in[(size_t)(2 * folded + t) * C + q];
              REQUIRE(sl >= 0 && sl < c->pmat_cap, "op %d: P slot %d out of range", folded, sl);
              op.cslot[kx][t][q] = sl;
            }
          }
          if (want_snap) {
            int r;
            if (rec_acquire(c, children_in[2 * folded], children_in[2 * folded + 1], &r)) return 1;
            op.crec_out[kx] = r;
            new_cherry_nodes.push_back(ch);
            new_cherry_recs.push_back(r);
          }
        } else if (snap_rec >= 0) {
          // a cherry kept by the input snapshot: its P matrices live in the library's pool
          op.kind[kx] = SRC_CHERRY;
          for (int t = 0; t < 2; ++t) {
            op.ctip[kx][t] = (const char*)c->d_codes + (size_t)(c->recs[snap_rec].tip[t] - 1) * c->P * c->code_bytes;
            for (int q = 0; q < C; ++q) op.cslot[kx][t][q] = CB_LIB_SLOT | (snap_rec * 2 * C + t * C + q);
          }
        } else if (kid_op[kx][i0] >= 0) {
          const int pp = pos_of[kid_op[kx][i0]];
          if (sched != SCHED_LEVELS && pp == p - 1 && range_id[pp] == range_id[p]) {
            op.kind[kx] = SRC_CARRIED;
          } else {
            const OpDesc& prod = c->h_ops[b + pp];
            REQUIRE(pp < p && prod.dst, "internal error: producer of node %d has no buffer", ch);
            op.kind[kx] = SRC_BUFFER;
            op.src[kx] = prod.dst;
            op.src_scale[kx] = prod.dst_scale;
          }
        } else {
          REQUIRE(sin && ch < (int)sin->buf_of_node.size() && sin->buf_of_node[ch] >= 0,
                  "op %d: child %d is neither recomputed nor in the input snapshot", i, ch);
          const Buffer& bf = c->buffers[sin->buf_of_node[ch]];
          op.kind[kx] = SRC_BUFFER;
          op.src[kx] = bf.data;
          op.src_scale[kx] = bf.scale;
        }
        for (int q = 0; q < C; ++q) {
          const int sl = pslots[(size_t)(2 * i + kx) * C + q];
          REQUIRE(sl >= 0 && sl < c->pmat_cap, "op %d: P slot %d out of range", i, sl);
          op.pslot[kx][q] = sl;
        }
        for (int q = C; q < CB_MAX_CATS; ++q) op.pslot[kx][q] = 0;
      }
    }

    if (sched == SCHED_WALK && !walk_segs.empty()) {
      single_launch = false;
      const int start = n_ranges;
      int mx = 1;
      for (auto& sg : walk_segs) {
        RangeDesc& r = c->h_ranges[n_ranges++];
        r.begin = b + sg.first; r.end = b + sg.second; r.out_index = -1; r.pad_ = 0;
        mx = std::max(mx, sg.second - sg.first);
      }
      launches.push_back({start, n_ranges});
      launch_maxops.push_back(mx);
      RangeDesc& r = c->h_ranges[n_ranges++];
      r.begin = b + walk_segs.back().second; r.end = e; r.out_index = li; r.pad_ = 0;
      launches.push_back({n_ranges - 1, n_ranges});
      launch_maxops.push_back(r.end - r.begin);
    } else if (sched != SCHED_LEVELS) {
      RangeDesc& r = c->h_ranges[n_ranges++];
      r.begin = b; r.end = e; r.out_index = li; r.pad_ = 0;
    } else {
      single_launch = false;
      c->level_of_op.assign(n, 1);
      int max_level = 1;
      for (int i = 0; i < n; ++i) {
        int lv = 1;
        for (int kx = 0; kx < 2; ++kx)
          if (kid_op[kx][i] >= 0) lv = std::max(lv, c->level_of_op[kid_op[kx][i]] + 1);
        c->level_of_op[i] = lv;
        max_level = std::max(max_level, lv);
      }
      std::vector<int> by_level(n);
      for (int i = 0; i < n; ++i) by_level[i] = i;
      std::stable_sort(by_level.begin(), by_level.end(),
                       [&](int x, int y) { return c->level_of_op[x] < c->level_of_op[y]; });
      int pos = 0;
      for (int lv = 1; lv <= max_level; ++lv) {
        const int start = n_ranges;
        while (pos < n && c->level_of_op[by_level[pos]] == lv) {
          const int i = b + by_level[pos++];
          RangeDesc& r = c->h_ranges[n_ranges++];
          r.begin = i; r.end = i + 1; r.out_index = (i == e - 1) ? li : -1; r.pad_ = 0;
        }
        if (n_ranges > start) {
          launches.push_back({start, n_ranges});
          launch_maxops.push_back(1);
        }
      }
    }
    for (int i = b; i < e; ++i) c->op_of_node[nodes[i]] = -1;
  }
  if (single_launch) {
    int mx = 0;
    for (int li = 0; li < n_lists; ++li) mx = std::max(mx, offsets[li + 1] - offsets[li]);
    launches.push_back({0, n_ranges});
    launch_maxops.push_back(mx);
  }

  if (ensure_lib_pool(c)) return 1;
  // upload descriptors + pi, launch
  CU(cudaMemcpyAsync(c->d_ops, c->h_ops, (size_t)total_ops * sizeof(OpDesc), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_ranges, c->h_ranges, (size_t)n_ranges * sizeof(RangeDesc), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_pi, pi, (size_t)c->n_states * 8, cudaMemcpyHostToDevice, c->stream));
  CU(cudaEventRecord(c->ev_stage, c->stream));
  c->h2d += (int64_t)total_ops * sizeof(OpDesc) + (int64_t)n_ranges * sizeof(RangeDesc) + c->n_states * 8;

  if (c->dmma_rc) {
    // every P matrix the ops use, re-laid out into the shared-memory stage image, in consumption order
    const int S8 = (c->n_states + 7) / 8 * 8;
    const size_t need = (size_t)total_ops * 2 * C * (S8 + RC_EXTRA_ROWS) * S8 * 8;
    if (need > c->staged_bytes) {
      CU(cudaStreamSynchronize(c->stream));
      if (c->d_staged) dev_free(c, c->d_staged, c->staged_bytes);
      c->d_staged = nullptr;
      c->staged_bytes = 0;
      const size_t cap = std::max(need, (size_t)64 << 20);
      if (dev_alloc(c, (void**)&c->d_staged, cap)) return 1;
      c->staged_bytes = cap;
    }
  }
  const LaunchConst k = make_const(c);
  CU(cudaEventRecord(c->ev0, c->stream));
  if (c->dmma_rc) {
    const unsigned jobs = (unsigned)(total_ops * 2 * C);
    switch ((c->n_states + 7) / 8 * 8) {
#define CB_RS_CASE(SS) case SS: rc_restage_kernel<SS><<<jobs, 256, 0, c->stream>>>(k, total_ops, c->d_staged); break
      CB_RS_CASE(16); CB_RS_CASE(24); CB_RS_CASE(32); CB_RS_CASE(40); CB_RS_CASE(48); CB_RS_CASE(56); CB_RS_CASE(64);
#undef CB_RS_CASE
      default: return fail("internal error: no register-carried DMMA kernel for %d states", c->n_states);
    }
    CU(cudaGetLastError());
    c->launches += 1;
  }
  for (size_t li = 0; li < launches.size(); ++li)
    if (launch_ranges(c, k, launches[li].first, launches[li].second, launch_maxops[li])) return 1;
  if (!c->family_s2) {
    dim3 grid((unsigned)((c->P + 255) / 256), (unsigned)n_lists);
    root_combine_kernel<<<grid, 256, 0, c->stream>>>(k);
    CU(cudaGetLastError());
    c->launches += 1;
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  c->timing_valid = true;

  if (c->comm) {
    int r = g_nccl.AllReduce(c->d_results, c->d_results, (size_t)n_lists, NCCL_DOUBLE, NCCL_SUM, c->comm, c->stream);
    REQUIRE(r == 0, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
  }
  CU(cudaMemcpyAsync(c->h_results, c->d_results, (size_t)n_lists * 8, cudaMemcpyDeviceToHost, c->stream));
  c->d2h += (int64_t)n_lists * 8;
  c->last_n_out = n_lists;

  // bookkeeping (stream-ordered: temporaries may be recycled by later launches on this stream)
  for (int bi : tmp_bufs) buffer_release(c, bi);
  if (want_snap) {
    int sid;
    if (!c->free_snaps.empty()) {
      sid = c->free_snaps.back();
      c->free_snaps.pop_back();
    } else {
      c->snaps.emplace_back();
      sid = (int)c->snaps.size() - 1;
    }
    Snapshot& sn = c->snaps[sid];
    sin = snapshot_in >= 0 ? &c->snaps[snapshot_in] : nullptr;  // vector may have moved
    if (sin) sn.buf_of_node = sin->buf_of_node; else sn.buf_of_node.assign(n_nodes, -1);
    if (sin) sn.cherry_of_node = sin->cherry_of_node; else sn.cherry_of_node.assign(n_nodes, -1);
    sn.buf_of_node.resize(n_nodes, -1);
    sn.cherry_of_node.resize(n_nodes, -1);
    sn.refs = 1;
    // a recomputed node replaces whatever the input snapshot held for it (buffer or folded cherry)
    for (size_t i = 0; i < new_nodes.size(); ++i) {
      sn.buf_of_node[new_nodes[i]] = -2 - (int)i;
      sn.cherry_of_node[new_nodes[i]] = -1;
    }
    for (size_t i = 0; i < new_cherry_nodes.size(); ++i) {
      sn.buf_of_node[new_cherry_nodes[i]] = -1;
      sn.cherry_of_node[new_cherry_nodes[i]] = -2 - (int)i;
    }
    for (int nd = 0; nd < n_nodes; ++nd) {
      int32_t& bi = sn.buf_of_node[nd];
      if (bi >= 0) c->buffers[bi].refs++;
      else if (bi <= -2) bi = new_bufs[-2 - bi];  // ownership moves from this evaluation to the snapshot
      int32_t& ri = sn.cherry_of_node[nd];
      if (ri >= 0) c->recs[ri].refs++;
      else if (ri <= -2) ri = new_cherry_recs[-2 - ri];
    }
    if (snapshot_out) *snapshot_out = sid;
  } else if (snapshot_out) {
    *snapshot_out = -1;
  }

  if (!(flags & CB_EVAL_NO_SYNC)) {
    CU(cudaStreamSynchronize(c->stream));
    if (lnl_out) memcpy(lnl_out, c->h_results, (size_t)n_lists * 8);
  }
  return 0;
}

extern "C" int cb_eval(cb_ctx* c, int snapshot_in, int n_ops, const int32_t* nodes, const int32_t* children,
                       const int32_t* pslots, const double* pi, int flags, int* snapshot_out, double* lnl_out) {
  const int32_t offsets[2] = {0, n_ops};
  return guarded([&] { return eval_impl(c, snapshot_in, 1, offsets, nodes, children, pslots, pi, flags, snapshot_out, lnl_out); });
}

extern "C" int cb_eval_batch(cb_ctx* c, int snapshot_in, int n_batch, const int32_t* op_offsets, const int32_t* nodes,
                             const int32_t* children, const int32_t* pslots, const double* pi, double* lnl_out) {
  REQUIRE(n_batch >= 1, "empty batch");
  return guarded([&] { return eval_impl(c, snapshot_in, n_batch, op_offsets, nodes, children, pslots, pi, 0, nullptr, lnl_out); });
}

extern "C" int cb_result_wait(cb_ctx* c, double* lnl_out) {
  REQUIRE(c, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  if (lnl_out && c->h_results) memcpy(lnl_out, c->h_results, (size_t)c->last_n_out * 8);
  return 0;
}

// --------------------------------------------------------------------------- introspection
extern "C" int cb_stats(cb_ctx* c, int64_t* launches, int64_t* h2d, int64_t* d2h, int64_t* dev_bytes) {
  REQUIRE(c, "null argument");
  if (launches) *launches = c->launches;
  if (h2d) *h2d = c->h2d;
  if (d2h) *d2h = c->d2h;
  if (dev_bytes) *dev_bytes = c->dev_bytes;
  return 0;
}
extern "C" int cb_last_eval_ms(cb_ctx* c, float* ms) {
  REQUIRE(c && ms && c->timing_valid, "no evaluation has been timed");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ev1));
  CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return 0;
}
extern "C" int cb_mark(cb_ctx* c, int which) {
  REQUIRE(c && (which == 0 || which == 1), "bad argument");
  CU(cudaSetDevice(c->device));
  CU(cudaEventRecord(c->ev_mark[which], c->stream));
  return 0;
}
extern "C" int cb_mark_elapsed_ms(cb_ctx* c, float* ms) {
  REQUIRE(c && ms, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaEventSynchronize(c->ev_mark[1]));
  CU(cudaEventElapsedTime(ms, c->ev_mark[0], c->ev_mark[1]));
  return 0;
}
extern "C" int cb_sync(cb_ctx* c) {
  REQUIRE(c, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int cb_flush_l2(cb_ctx* c) {
  REQUIRE(c, "null argument");
  CU(cudaSetDevice(c->device));
  if (!c->d_flush) {
    c->flush_bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    if (dev_alloc(c, &c->d_flush, c->flush_bytes)) return 1;
  }
  CU(cudaMemsetAsync(c->d_flush, 0x5a, c->flush_bytes, c->stream));
  return 0;
}
