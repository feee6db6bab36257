// Device-visible descriptors shared by the host scheduler and the pruning kernels.
#pragma once
#include <stdint.h>
#include "../../include/cybayes_b200.h"

namespace cb {

enum SrcKind : int32_t { SRC_TIP = 0, SRC_BUFFER = 1, SRC_CARRIED = 2 };

// One node operation: L_node = (P_a . L_a) * (P_b . L_b), ML_gamma.pyx:24-36.
// 128 bytes, read by every block of a launch through the read-only path.
struct OpDesc {
  double* dst;                  // [C][S][P] or nullptr (not stored)
  int32_t* dst_scale;           // [P] (2-state family) or [C][P] (general family)
  const void* src[2];           // tip code row, partial buffer, or nullptr when carried
  const int32_t* src_scale[2];  // nullptr for tips
  int32_t kind[2];              // SrcKind
  int32_t pslot[2][CB_MAX_CATS];
  int32_t is_root;
  int32_t pad_;
};
static_assert(sizeof(OpDesc) == 128, "OpDesc must stay 128 bytes");

// A block walks ops [begin, end) in order for its site tile; results of a root op go to
// results[out_index].
struct RangeDesc {
  int32_t begin, end, out_index, pad_;
};

struct LaunchConst {
  const OpDesc* ops;
  const RangeDesc* ranges;
  const double* pmats;     // slot pool, S*S doubles per slot
  const double* weights;   // [P]
  const double* pi;        // [S]
  const double* amb;       // [n_amb][S] 0/1
  double* block_sums;      // [n_out][max_blocks]
  unsigned int* tickets;   // [n_out]
  double* results;         // [n_out]
  double* root_dot;        // general family: [n_out][C][P] per-category pi.L_root
  int32_t* root_exp;       // general family: [n_out][C][P]
  int64_t n_sites;         // padded pattern count P (multiple of 64)
  int32_t n_states, n_cats, code_bytes, max_blocks;
  double cats;            // n_cats as a double (the reference divides, ML_gamma.pyx:38)
};

}  // namespace cb
