// scan_mq_f32.cu -- instantiates the multi-query scan for fp32 rows (see scan_mq.cuh).
#include "scan_mq.cuh"

namespace evdb {
scan_fn_t pick_float_mq_f32(int metric, int tpr, int Q) { return pick_float_mq_m<EVDB_F32>(metric, tpr, Q); }
}  // namespace evdb
