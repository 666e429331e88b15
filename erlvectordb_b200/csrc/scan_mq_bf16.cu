// scan_mq_bf16.cu -- instantiates the multi-query scan for bf16 rows (see scan_mq.cuh).
#include "scan_mq.cuh"

namespace evdb {
scan_fn_t pick_float_mq_bf16(int metric, int tpr, int Q) { return pick_float_mq_m<EVDB_BF16>(metric, tpr, Q); }
}  // namespace evdb
