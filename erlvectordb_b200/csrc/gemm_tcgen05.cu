// gemm_tcgen05.cu -- batched cosine as a tcgen05/TMEM GEMM (family B).  Placeholder until the
// kernel lands: the planner never selects it.
#include "internal.h"

namespace evdb {

bool gemm_plan_supported(evdb_store *, int, int, int) { return false; }

int launch_gemm_topk(evdb_store *, int, int, int, uint64_t *, int *, cudaStream_t) {
    return EVDB_E_UNSUPPORTED;
}

}  // namespace evdb
