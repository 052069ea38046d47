// batch.cu - batched node-LP relaxations sharing one matrix.
#include "internal.h"

namespace mipcu {

void solve_lp_batch(mipcu_ctx*, int32_t, const double*, const double*, mipcu_stats*, double*) {
  throw Error("mipcu_solve_lp_batch is not built yet");
}

}  // namespace mipcu
