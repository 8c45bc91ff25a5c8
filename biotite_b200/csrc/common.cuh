// Shared device/host definitions of libb200align.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <nvtx3/nvToolsExt.h>

namespace bt {

// NVTX range of one host-side phase (plan / pack + fill / traceback / copies): header-only NVTX 3,
// a no-op unless a profiler has injected its library.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// Direction bits, identical to the reference's bitflags
// (src/rust/sequence/align/cell.rs:44-80) so that direction tables can be
// traced with the reference's rules.
enum : uint8_t {
    L_MATCH = 1, L_GAP1 = 2, L_GAP2 = 4,
    A_MM = 1, A_G1M = 2, A_G2M = 4, A_MG1 = 8, A_G1G1 = 16, A_MG2 = 32, A_G2G2 = 64,
    DIR_MARK = 128  // "M of this cell equals the optimum" (local start enumeration)
};
enum : int { ST_MATCH = 0, ST_GAP1 = 1, ST_GAP2 = 2 };
enum : uint8_t { OP_DIAG = 0, OP_TO1 = 1, OP_TO2 = 2 };

// Mode of one launch (device copy of BtParams plus derived values).
struct DevParams {
    int affine, local, terminal, use_matrix, banded, score_only;
    int32_t gap_open, gap_ext, match, mismatch, neg_inf;
    int n_rows, n_cols, code_bytes;
    int mark;            // second pass of local enumeration: tag cells whose M == mark_value
    int32_t mark_value;
    int matrix_in_smem;  // matrix staged in shared memory
    int stats;           // traceback emits GapStats (in `first`) instead of step bytes
    int dir_pad;         // direction tables have their row stride rounded up to 16 (wavefront kernel)
    int lean;            // first-optimal directions suffice and the mode fits strip_kernel.cuh
};

// row stride of a direction table with `cols` columns
__host__ __device__ __forceinline__ int64_t table_stride(int dir_pad, int64_t cols) {
    return dir_pad ? (cols + 15) & ~(int64_t)15 : cols;
}

// Which sequences pair p aligns.  off1/off2 are the offsets of two sequence pools; idx1/idx2
// map a pair to its sequences (nullptr = identity: pair p aligns sequence p of either pool,
// the plain many-pairs batch).  Usable with host or device pointers.
struct PairMap {
    const int64_t *off1 = nullptr, *off2 = nullptr;
    const int32_t *idx1 = nullptr, *idx2 = nullptr;
    const int64_t *ops_off = nullptr;   // pair -> first step byte (nullptr: off1[p] + off2[p])
    __host__ __device__ __forceinline__ int64_t beg1(int64_t p) const { return off1[idx1 ? idx1[p] : p]; }
    __host__ __device__ __forceinline__ int64_t beg2(int64_t p) const { return off2[idx2 ? idx2[p] : p]; }
    __host__ __device__ __forceinline__ int64_t len1(int64_t p) const {
        const int64_t q = idx1 ? idx1[p] : p;
        return off1[q + 1] - off1[q];
    }
    __host__ __device__ __forceinline__ int64_t len2(int64_t p) const {
        const int64_t q = idx2 ? idx2[p] : p;
        return off2[q + 1] - off2[q];
    }
    __host__ __device__ __forceinline__ int64_t ops(int64_t p) const { return ops_off ? ops_off[p] : off1[p] + off2[p]; }
};

// Device views of one batch (or one chunk of it).
struct BatchView {
    const void *codes1, *codes2;
    PairMap map;                  // pair -> sequences
    const int64_t *bands;         // 2*n_pairs (lower, upper) or nullptr
    const int32_t *matrix;        // n_rows*n_cols or nullptr
    uint8_t *dirs;                // direction table storage of this chunk
    const int64_t *dir_off;       // per pair byte offset into dirs (relative to chunk)
    int32_t *cand;                // banded semi-global end candidates
    const int64_t *cand_off;      // per pair offset into cand (ints)
    int32_t *gscratch;            // per block wavefront buffers when they exceed smem
    int64_t gscratch_stride;      // ints per block
    // results
    int32_t *score;
    int64_t *start_idx;
    int32_t *start_state;
    int32_t *end_vals;            // 3 per pair: M, G1, G2 of the global end cell
    int64_t *trace_len;
    int32_t *first;               // 2 per pair
    uint8_t *ops;                 // compact traces, pair k at off1[k]+off2[k]
    int64_t pair_begin, pair_end; // pairs handled by this launch
};

// Alignment statistics accumulated during a traceback instead of storing the step bytes:
// number of columns and of gap-opening / gap-extending columns as counted by the
// reference's guide-tree code (src/biotite/sequence/align/multiple.pyx:403-455 `_count_gaps`):
// a gap column opens a gap if the previous column had no gap in that sequence, else it
// extends one; without terminal penalty the columns before the first and after the last
// gap-free column are ignored.  The counts do not depend on the direction of the walk.
struct GapStats {
    int len = 0, opens = 0, exts = 0;
    int pend_opens = 0, pend_exts = 0;
    int prev = -1;
    bool inner = false;       // a gap-free column has been seen
    bool count_terminal;
    __host__ __device__ explicit GapStats(bool terminal) : count_terminal(terminal) {}
    __host__ __device__ __forceinline__ void push(int op) {
        len++;
        if (op == OP_DIAG) {
            opens += pend_opens; exts += pend_exts;
            pend_opens = pend_exts = 0;
            inner = true;
        } else if (count_terminal || inner) {
            if (op == prev) pend_exts++; else pend_opens++;
        }
        prev = op;
    }
    __host__ __device__ __forceinline__ void finish() {
        if (count_terminal) { opens += pend_opens; exts += pend_exts; }
    }
};

__host__ __device__ inline int32_t wadd(int32_t a, int32_t b) {
    return (int32_t)((uint32_t)a + (uint32_t)b);  // i32 wraps in the reference
}
__host__ __device__ inline int32_t wsub(int32_t a, int32_t b) {
    return (int32_t)((uint32_t)a - (uint32_t)b);
}

__device__ __forceinline__ uint64_t load_code(const void *base, int64_t idx, int bytes) {
    switch (bytes) {
    case 1: return ((const uint8_t *)base)[idx];
    case 2: return ((const uint16_t *)base)[idx];
    case 4: return ((const uint32_t *)base)[idx];
    default: return ((const uint64_t *)base)[idx];
    }
}

}  // namespace bt
