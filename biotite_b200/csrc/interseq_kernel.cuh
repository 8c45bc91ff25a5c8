// placeholder, replaced by the packed 16-bit inter-sequence kernels
#pragma once
#include <vector>
#include "../../include/b200align.h"
#include "common.cuh"
namespace bt {
struct InterseqPlan {};
struct InterseqIO {
    const void *codes1, *codes2; const int64_t *off1, *off2; const int32_t *matrix;
    int32_t *score; int64_t *trace_len; int32_t *first; uint8_t *ops;
};
inline bool interseq_eligible(const BtParams &, const std::vector<int64_t> &, const std::vector<int64_t> &,
                              int64_t, int32_t, int32_t) { return false; }
inline int interseq_prepare(InterseqPlan &, const BtParams &, const std::vector<int64_t> &,
                            const std::vector<int64_t> &, int64_t, int, int32_t, int32_t, int64_t *,
                            cudaStream_t) { return BT_ERR_INVALID_ARG; }
inline cudaError_t interseq_run_fill(InterseqPlan &, const InterseqIO &, cudaStream_t, int64_t *) {
    return cudaErrorNotSupported;
}
inline cudaError_t interseq_run_traceback(InterseqPlan &, const InterseqIO &, cudaStream_t, int64_t *) {
    return cudaErrorNotSupported;
}
}  // namespace bt
