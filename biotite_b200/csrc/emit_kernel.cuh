// Wire-format emitters over the compact device-resident traces:
// CIGAR operations (reference: src/biotite/sequence/align/cigar.py:196-392
// `write_alignment_to_cigar`) and gapped sequence strings (alignment.py:233-243
// `get_gapped_sequences`, the payload of io/fasta/convert.py:236-262 `set_alignment`),
// produced straight from the step bytes the traceback kernels left in HBM, so that a batch of
// 10^6 alignments never materialises its (L, 2) int64 trace arrays.
//
// Compact trace of pair p (all routes): `first` = table cell (i, j) of the first alignment
// column, `ops` = one step byte per column stored end -> start (ops[L-1] belongs to the first
// column and is not needed), see host_trace.hpp `expand_ops`.
#pragma once
#include "common.cuh"

namespace bt {

// BAM numbering of cigar.py:20-35
enum : uint32_t { CG_MATCH = 0, CG_INS = 1, CG_DEL = 2, CG_SOFT = 4, CG_HARD = 5, CG_EQUAL = 7, CG_DIFF = 8 };
enum : int { EMIT_DISTINGUISH = 1, EMIT_HARD_CLIP = 2, EMIT_TERMINAL_GAPS = 4 };

struct EmitArgs {
    const void *codes1, *codes2;
    PairMap map;
    const int64_t *bands;       // banded batches: (lower, upper) per pair, else nullptr
    const uint8_t *swapped;     // per pair: the caller swapped the two sequences (align_banded), or nullptr
    const int64_t *trace_len;
    const int32_t *first;
    const uint8_t *ops;
    int code_bytes;
    int64_t n_pairs;
};

// Walks the columns of pair p from start to end; f(column index, position in sequence 1 or -1,
// position in sequence 2 or -1).
template <typename F>
__device__ __forceinline__ void emit_walk(const EmitArgs &A, int64_t p, F f) {
    const int64_t len = A.trace_len[p];
    if (len <= 0) return;
    const uint8_t *ops = A.ops + A.map.ops(p);
    int64_t s0 = (int64_t)A.first[2 * p] - 1;
    int64_t s1 = A.bands ? (int64_t)A.first[2 * p + 1] + s0 + A.bands[2 * p] - 1 : (int64_t)A.first[2 * p + 1] - 1;
    int64_t col = 0;
    if (!(s0 == -1 && s1 == -1)) f(col++, s0, s1);
    int64_t a = s0, b = s1;  // last consumed positions
    for (int64_t k = len - 2; k >= 0; k--) {
        const uint8_t op = ops[k];
        if (op == OP_DIAG) { a++; b++; f(col++, a, b); }
        else if (op == OP_TO1) { b++; f(col++, (int64_t)-1, b); }
        else { a++; f(col++, a, (int64_t)-1); }
    }
}

// One thread per pair.  Output: BAM-encoded operations (length << 4 | op) of pair p at
// cigar[ops offset of p ...], their number in n_ops[p].  Reference sequence = sequence 1,
// segment = sequence 2 (cigar.py defaults), exchanged for pairs flagged in `swapped`.
__global__ void emit_cigar_kernel(const EmitArgs A, int flags, int32_t *__restrict__ n_ops,
                                  uint32_t *__restrict__ cigar) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pairs) return;
    const bool swap = A.swapped && A.swapped[p];
    const int64_t seg_len = swap ? A.map.len1(p) : A.map.len2(p);
    // pass 1: columns of the first / last aligned segment symbol (cigar.py:396-421)
    int64_t col_a = -1, col_b = -1, seg_a = 0, seg_b = 0;
    emit_walk(A, p, [&](int64_t col, int64_t s0, int64_t s1) {
        const int64_t seg = swap ? s0 : s1;
        if (seg >= 0) {
            if (col_a < 0) { col_a = col; seg_a = seg; }
            col_b = col; seg_b = seg;
        }
    });
    uint32_t *out = cigar + A.map.ops(p);
    int32_t count = 0;
    if (col_a < 0) { n_ops[p] = 0; return; }  // no segment symbol aligned (the reference would raise)
    const uint32_t clip = (flags & EMIT_HARD_CLIP) ? CG_HARD : CG_SOFT;
    if (seg_a > 0) out[count++] = ((uint32_t)seg_a << 4) | clip;
    const bool terminal = (flags & EMIT_TERMINAL_GAPS) != 0;
    const int64_t b1 = A.map.beg1(p), b2 = A.map.beg2(p);
    uint32_t run_op = 0xffffffffu, run_len = 0;
    emit_walk(A, p, [&](int64_t col, int64_t s0, int64_t s1) {
        if (!terminal && (col < col_a || col > col_b)) return;
        const int64_t ref = swap ? s1 : s0, seg = swap ? s0 : s1;
        uint32_t op;
        if (ref < 0) op = CG_INS;
        else if (seg < 0) op = CG_DEL;
        else if (flags & EMIT_DISTINGUISH)
            op = load_code(A.codes1, b1 + s0, A.code_bytes) == load_code(A.codes2, b2 + s1, A.code_bytes) ? CG_EQUAL : CG_DIFF;
        else op = CG_MATCH;
        if (op == run_op) { run_len++; return; }
        if (run_len) out[count++] = (run_len << 4) | run_op;
        run_op = op; run_len = 1;
    });
    if (run_len) out[count++] = (run_len << 4) | run_op;
    const int64_t end_clip = seg_len - seg_b - 1;
    if (end_clip > 0) out[count++] = ((uint32_t)end_clip << 4) | clip;
    n_ops[p] = count;
}

// One thread per pair.  Writes the two gapped strings of pair p (trace_len[p] characters each)
// at gapped1 / gapped2 + ops offset of p; symbols are looked up in the alphabets' byte tables.
__global__ void emit_gapped_kernel(const EmitArgs A, const uint8_t *__restrict__ symbols1, int n_sym1,
                                   const uint8_t *__restrict__ symbols2, int n_sym2,
                                   uint8_t *__restrict__ gapped1, uint8_t *__restrict__ gapped2,
                                   int64_t *__restrict__ n_cols) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pairs) return;
    const bool swap = A.swapped && A.swapped[p];
    const int64_t off = A.map.ops(p), b1 = A.map.beg1(p), b2 = A.map.beg2(p);
    uint8_t *o1 = (swap ? gapped2 : gapped1) + off, *o2 = (swap ? gapped1 : gapped2) + off;
    int64_t cols = 0;
    emit_walk(A, p, [&](int64_t col, int64_t s0, int64_t s1) {
        uint8_t c1 = '-', c2 = '-';
        if (s0 >= 0) { const uint64_t c = load_code(A.codes1, b1 + s0, A.code_bytes); c1 = c < (uint64_t)n_sym1 ? symbols1[c] : (uint8_t)'?'; }
        if (s1 >= 0) { const uint64_t c = load_code(A.codes2, b2 + s1, A.code_bytes); c2 = c < (uint64_t)n_sym2 ? symbols2[c] : (uint8_t)'?'; }
        o1[col] = c1; o2[col] = c2;
        cols = col + 1;
    });
    n_cols[p] = cols;
}

}  // namespace bt
