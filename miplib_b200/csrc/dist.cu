// dist.cu - row-sharded solve across GPUs (one process per GPU, NCCL over NVLink).
#include "internal.h"

namespace mipcu {

void dist_unique_id(void*) { throw Error("row-sharded contexts are not built yet"); }
void dist_init(mipcu_ctx*, int, int, const void*) { throw Error("row-sharded contexts are not built yet"); }
void dist_destroy(mipcu_ctx*) {}

}  // namespace mipcu
