/*
 * b200align.h - C ABI of libb200align.so, the B200 (sm_100a) implementation of
 * the pairwise dynamic-programming alignment hot path of biotite.
 *
 * This is the drop-in boundary.  It replaces the two native entry points the
 * reference exports from its PyO3 module `biotite.rust.sequence.align`
 * (registered at src/rust/sequence/align/mod.rs:23-24):
 *
 *   align_optimal(code1, code2, scoring, gap_penalty, terminal_penalty, local,
 *                 score_only, max_number)        src/rust/sequence/align/pairwise.rs:57-70
 *   align_banded (code1, code2, scoring, gap_penalty, lower_diag, upper_diag,
 *                 local, score_only, max_number) src/rust/sequence/align/banded.rs:63-77
 *
 * and adds the batched many-pairs entry points the north star asks for.  All
 * functions take plain pointers and sizes; the caller owns every buffer it
 * passes in.  All functions return BT_OK (0) or a negative BtStatus; the
 * message of the last failure on the calling thread is bt_last_error().
 * There is no CPU fallback: without a CUDA device every compute entry point
 * fails with BT_ERR_CUDA.
 */
#ifndef B200ALIGN_H
#define B200ALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum BtStatus {
    BT_OK = 0,
    BT_ERR_INVALID_ARG = -1,  /* -> ValueError  */
    BT_ERR_INVALID_CODE = -2, /* a symbol code indexes outside the matrix -> ValueError */
    BT_ERR_CUDA = -3,         /* -> RuntimeError */
    BT_ERR_OOM = -4,          /* -> MemoryError  */
    BT_ERR_RANGE = -5         /* scores may leave the exact integer range -> OverflowError */
} BtStatus;

/*
 * Alignment mode.  The fields mirror the reference's run-time switches and
 * const generics: scoring scheme (scoring.rs:130-147), gap model
 * (GapPenalty::{Linear,Affine}), LOCAL / TERMINAL_PENALTY (pairwise.rs:348,563)
 * and full table vs. band (pairwise.rs vs. banded.rs).
 */
typedef struct BtParams {
    int32_t affine;           /* 0: linear gap penalty (gap_open), 1: affine (gap_open, gap_ext) */
    int32_t gap_open;         /* <= 0 */
    int32_t gap_ext;          /* <= 0; ignored when !affine */
    int32_t local;            /* Smith-Waterman instead of Needleman-Wunsch */
    int32_t terminal_penalty; /* 0: semi-global (terminal gaps free); ignored when local or banded */
    int32_t use_matrix;       /* 1: substitution matrix, 0: (match, mismatch) */
    int32_t match_score;
    int32_t mismatch_score;
    int32_t n_rows;           /* matrix shape = (alphabet1 size, alphabet2 size) */
    int32_t n_cols;
    int32_t banded;           /* 0: align_optimal, 1: align_banded */
    int32_t code_bytes;       /* element size of the code arrays: 1, 2, 4 or 8 (util.rs:37-52) */
} BtParams;

/* ---- library / device ------------------------------------------------------------ */

const char *bt_last_error(void);
const char *bt_version(void);
int bt_device_count(void);          /* number of CUDA devices, 0 if none / no driver */
int bt_set_device(int device);      /* device used by the calling thread */
int bt_get_device(void);

/* page-locked host memory for full-speed asynchronous copies (cudaHostAlloc / cudaFreeHost) */
void *bt_pinned_alloc(size_t bytes);
void bt_pinned_free(void *ptr);

/* Device buffers of finished batches are cached per device for reuse; this returns them to
 * the driver (cudaFree). */
void bt_pool_trim(void);

/* ---- single pair: the reference's FFI contract -------------------------------------
 *
 * code1/code2: contiguous arrays of `code_bytes`-wide unsigned codes (len1, len2
 * elements).  matrix: row-major int32 (n_rows x n_cols) or NULL when
 * !use_matrix.  lower_diag/upper_diag: the (already cropped) band, banded only;
 * len1 <= len2 is the caller's job as in banded.py:215-236.
 * score_out receives the optimal score.  Unless score_only, *traces_out
 * receives a trace list holding at most max_number traces in the reference's
 * emission order (trace.rs:46-92); release it with bt_tracelist_free().
 */
typedef struct BtTraceList BtTraceList;

int bt_align_pair(const BtParams *params, const void *code1, int64_t len1,
                  const void *code2, int64_t len2, const int32_t *matrix,
                  int64_t lower_diag, int64_t upper_diag, int32_t score_only,
                  int64_t max_number, int32_t *score_out, BtTraceList **traces_out);

int64_t bt_tracelist_count(const BtTraceList *list);
int64_t bt_tracelist_length(const BtTraceList *list, int64_t k); /* rows L of trace k */
/* copies trace k as C-order int64 (L, 2) with -1 gaps (trace.rs:103-130) */
int bt_tracelist_copy(const BtTraceList *list, int64_t k, int64_t *out);
void bt_tracelist_free(BtTraceList *list);

/* ---- batch of independent pairs (max_number = 1 semantics) --------------------------
 *
 * Pair k aligns codes1[off1[k] .. off1[k+1]) with codes2[off2[k] .. off2[k+1]).
 * bands: NULL, or 2*n_pairs int64 (lower, upper) per pair, cropped as above.
 * A batch object keeps its inputs and results resident in HBM:
 *   create  : validate, H2D copies
 *   run     : fill + traceback kernels; *device_ms (may be NULL) receives the
 *             CUDA-event time of exactly those kernels on the batch's stream
 *   fetch_* : D2H of results
 * score_only: 0 = scores + traces, 1 = scores only, 2 = scores + trace statistics: the
 * traceback runs on the GPU but emits, instead of step bytes, trace_len = number of alignment
 * columns and first[2k], first[2k+1] = number of gap-opening / gap-extending columns as the
 * reference's guide-tree code counts them (multiple.pyx:403-455; terminal gap columns are
 * skipped when !terminal_penalty).  Not available for banded batches.
 * Results: scores (int32 per pair) and, unless score_only, one first-optimal
 * trace per pair in compact form: trace_len[k] = L, first[2k..2k+1] = table
 * coordinates of the first alignment column's cell, ops = L step bytes
 * (0 = both advance, 1 = gap in sequence 1, 2 = gap in sequence 2) in
 * end-to-start order starting at ops_off[k] = off1[k] + off2[k]
 * (capacity len1 + len2).  bt_expand_traces() turns them into (L,2) int64.
 */
typedef struct BtBatch BtBatch;

int bt_batch_create(const BtParams *params, const void *codes1, const int64_t *off1,
                    const void *codes2, const int64_t *off2, int64_t n_pairs,
                    const int32_t *matrix, const int64_t *bands, int32_t score_only,
                    BtBatch **batch_out);
/*
 * Indexed batch: the sequences live in two pools (pool k: n_seqk sequences, offk has
 * n_seqk + 1 entries) and pair p aligns sequence idx1[p] of pool 1 with sequence idx2[p] of
 * pool 2.  This is how one-vs-many searches (BASELINE config 3) and all-vs-all score
 * matrices (config 5; the guide-tree loop of the reference, multiple.pyx:322-333) are run
 * without materialising the pairs.  The step bytes of pair p start at the prefix sum of
 * len1 + len2 over the pairs before it (bt_batch_ops_bytes() bytes in total).
 */
int bt_batch_create_indexed(const BtParams *params, const void *codes1, const int64_t *off1,
                            int64_t n_seq1, const void *codes2, const int64_t *off2, int64_t n_seq2,
                            const int32_t *idx1, const int32_t *idx2, int64_t n_pairs,
                            const int32_t *matrix, const int64_t *bands, int32_t score_only,
                            BtBatch **batch_out);
/*
 * Generated batch: the pairs are not listed by the caller but enumerated on the device, which
 * also sorts, groups and packs them (csrc/device_plan.cuh) - the host never holds a per-pair
 * array.  BT_PAIRS_CROSS: pair p = i * n_seq2 + j aligns sequence i of pool 1 with sequence j of
 * pool 2 (one-vs-many searches, BASELINE config 3: the loop over align_optimal(q, d, ...,
 * score_only=True) of a database scan).  BT_PAIRS_TRIANGLE: both pools are the same n sequences
 * and pair p is the p-th (i, j), i <= j, in row-major order - the guide-tree loop of the reference
 * (src/biotite/sequence/align/multiple.pyx:322-333); every pair is aligned shorter sequence first,
 * so the matrix must be symmetric.  score_only must be 1.  Scores come back in pair order
 * (bt_batch_fetch_scores) or, for a triangle, as the symmetric (n, n) matrix
 * (bt_batch_fetch_score_matrix).  Modes outside the packed 16-bit kernels (no matrix, banded,
 * wide codes, linear semi-global) are refused with BT_ERR_INVALID_ARG: use the indexed form.
 */
typedef enum BtPairKind { BT_PAIRS_CROSS = 0, BT_PAIRS_TRIANGLE = 1 } BtPairKind;
int bt_batch_create_generated(const BtParams *params, const void *codes1, const int64_t *off1,
                              int64_t n_seq1, const void *codes2, const int64_t *off2, int64_t n_seq2,
                              int32_t pair_kind, const int32_t *matrix, int32_t score_only,
                              BtBatch **batch_out);
int bt_batch_fetch_score_matrix(BtBatch *batch, int32_t *matrix_out);
/*
 * The k best hits of every sequence of pool 1 of a BT_PAIRS_CROSS batch, selected on the device
 * (a database scan keeps its best hits, not 4 bytes per database sequence): scores_out and
 * index_out receive n_seq1 * k entries, row i holding the k highest scores of sequence i against
 * pool 2 and the pool-2 indices they belong to - higher score first is the caller's sort; among
 * equal scores the lower indices are kept; rows are in ascending index order, padded with
 * (INT32_MIN, -1) when n_seq2 < k.
 */
int bt_batch_fetch_topk(BtBatch *batch, int32_t k, int32_t *scores_out, int32_t *index_out);
int64_t bt_batch_ops_bytes(const BtBatch *batch);
int bt_batch_run(BtBatch *batch, float *device_ms);
int bt_batch_fetch_scores(BtBatch *batch, int32_t *scores_out);
int bt_batch_fetch_traces(BtBatch *batch, int64_t *trace_len_out, int32_t *first_out,
                          uint8_t *ops_out);
/*
 * Wire formats emitted on the device from the compact traces of a traced batch, without
 * expanding them to (L,2) arrays first.
 *
 * bt_batch_fetch_cigar replaces a loop over write_alignment_to_cigar
 * (src/biotite/sequence/align/cigar.py:196-392; reference = sequence 1, segment = sequence 2):
 * cigar_out has bt_batch_ops_bytes() uint32 entries, the operations of pair k start at its
 * step-byte offset and are BAM-encoded (length << 4 | op, op numbering of cigar.py:20-35),
 * n_ops_out[k] is their number.  flags: 1 = distinguish matches ('=' / 'X'), 2 = hard clip,
 * 4 = include terminal gaps.  swapped (NULL or one byte per pair) marks pairs whose two
 * sequences were exchanged by the caller (align_banded, banded.py:217-224).
 *
 * bt_batch_fetch_gapped replaces Alignment.get_gapped_sequences (alignment.py:233-243, the
 * payload of fasta.set_alignment, io/fasta/convert.py:236-262): the two gapped strings of
 * pair k ('-' = gap, symbolsN[code] otherwise) start at its step-byte offset in gapped1_out /
 * gapped2_out (bt_batch_ops_bytes() bytes each) and have n_columns_out[k] characters.
 */
int bt_batch_fetch_cigar(BtBatch *batch, int32_t flags, const uint8_t *swapped, int32_t *n_ops_out,
                         uint32_t *cigar_out);
int bt_batch_fetch_gapped(BtBatch *batch, const uint8_t *symbols1, int32_t n_symbols1,
                          const uint8_t *symbols2, int32_t n_symbols2, const uint8_t *swapped,
                          int64_t *n_columns_out, uint8_t *gapped1_out, uint8_t *gapped2_out);
/* device time of the last run split into DP fill and traceback kernels (CUDA events) */
int bt_batch_timings(const BtBatch *batch, float *fill_ms, float *traceback_ms);
/* DP cells the batch updates per run (sum of len1*len2, or in-band in-sequence cells) */
int64_t bt_batch_cells(const BtBatch *batch);
/* launches issued by the last bt_batch_run (this library's own kernels only) */
int64_t bt_batch_launches(const BtBatch *batch);
/* fill launches of the last bt_batch_run: one per window (packed routes) or chunk (general route) */
int64_t bt_batch_windows(const BtBatch *batch);
/* name of the fill kernel family the batch was routed to ("general_i32", "interseq_s16x2", ...) */
const char *bt_batch_kernel(const BtBatch *batch);
void bt_batch_destroy(BtBatch *batch);

/* one-shot convenience: create + run + fetch + destroy; any output may be NULL */
int bt_align_batch(const BtParams *params, const void *codes1, const int64_t *off1,
                   const void *codes2, const int64_t *off2, int64_t n_pairs,
                   const int32_t *matrix, const int64_t *bands, int32_t score_only,
                   int32_t *scores_out, int64_t *trace_len_out, int32_t *first_out,
                   uint8_t *ops_out);

/*
 * The same call with the traces in compact form, which is what should cross PCIe when the
 * batch is large: trace_len as int32, and the steps as 2 bits each (0 = both advance, 1 = gap
 * in sequence 1, 2 = gap in sequence 2; end-to-start order like the byte form).  Step s of pair
 * k is bits 2 (s % 16) .. 2 (s % 16) + 1 of word ops2_out[ops2_off_out[k] + s / 16].  The
 * (L, 2) int64 trace the reference's FFI returns (src/rust/sequence/align/trace.rs:103-130) is
 * rebuilt from first + steps by bt_expand_traces2().  ops2_out must hold
 * bt_ops2_capacity(off1, off2, n_pairs) words (a bound from the sequence lengths); only the
 * words in use are copied from the device (about L / 4 bytes per pair); *ops2_words_out
 * (optional) receives the end of the used range.  score_only: 0 or 1.  Page-locked output
 * arrays (bt_pinned_alloc) let the copies of a window overlap the kernels of the next.
 */
int64_t bt_ops2_capacity(const int64_t *off1, const int64_t *off2, int64_t n_pairs);
int bt_align_batch_compact(const BtParams *params, const void *codes1, const int64_t *off1,
                           const void *codes2, const int64_t *off2, int64_t n_pairs,
                           const int32_t *matrix, const int64_t *bands, int32_t score_only,
                           int32_t *scores_out, int32_t *trace_len_out, int32_t *first_out,
                           uint32_t *ops2_out, int64_t ops2_capacity, int64_t *ops2_off_out,
                           int64_t *ops2_words_out);
/*
 * The same batch over several GPUs of the node from ONE process: `devices` lists n_devices CUDA
 * device ordinals (NULL: all of bt_device_count()).  The pairs are cut into contiguous shards of
 * equal DP cells; one host thread per device runs its shard with its own streams and writes
 * straight into disjoint slices of the caller's arrays - no collective, pairs are independent
 * (src/rust/sequence/align/pairwise.rs:161 aligns one pair).  Results as in
 * bt_align_batch_compact; ops2_out needs bt_ops2_capacity() + 16 * n_devices words.
 * shard_begin_out (optional, n_devices + 1 entries) receives the first pair of every shard.
 */
int bt_align_batch_multi(const int32_t *devices, int32_t n_devices, const BtParams *params,
                         const void *codes1, const int64_t *off1, const void *codes2,
                         const int64_t *off2, int64_t n_pairs, const int32_t *matrix,
                         const int64_t *bands, int32_t score_only, int32_t *scores_out,
                         int32_t *trace_len_out, int32_t *first_out, uint32_t *ops2_out,
                         int64_t ops2_capacity, int64_t *ops2_off_out, int64_t *ops2_words_out,
                         int64_t *shard_begin_out);
int bt_expand_traces2(int64_t n_pairs, const int32_t *trace_len, const int32_t *first,
                      const uint32_t *ops2, const int64_t *ops2_off, const int64_t *bands,
                      const int64_t *rows_off, int64_t *out);

/*
 * Local X-drop extension of regions: the native part of align_local_gapped, replaces
 * `align_region` (src/rust/sequence/align/localgapped.rs:57-70; called from
 * src/biotite/sequence/align/localgapped.py:235-266 once per side of the seed).  Region r
 * aligns codes1[off1[r]..off1[r+1]) with codes2[off2[r]..off2[r+1]) starting at their first
 * symbols; all regions of a call run concurrently (one thread block each).  scores_out[r] is
 * the region's score, status_out[r] = 1 where the reference would raise MemoryError
 * ("Maximum table size exceeded").  Unless score_only, lists_out[r] receives up to max_number
 * co-optimal traces in the reference's order (free with bt_tracelist_free; NULL for a region
 * whose status is 1).  params: affine / gap_open / gap_ext, use_matrix or match / mismatch,
 * n_rows / n_cols, code_bytes; local, terminal_penalty and banded are ignored.
 */
int bt_xdrop_regions(const BtParams *params, const void *codes1, const int64_t *off1,
                     const void *codes2, const int64_t *off2, int64_t n_regions,
                     const int32_t *matrix, int32_t threshold, int64_t max_table_size,
                     int32_t score_only, int64_t max_number, int32_t *scores_out,
                     int32_t *status_out, BtTraceList **lists_out);

/*
 * Host-side expansion of compact traces into the reference's (L,2) int64 layout.
 * rows_off[k] (n_pairs+1 entries) = prefix sum of trace_len; out has 2*rows_off[n_pairs]
 * entries.  bands as in bt_batch_create (NULL for align_optimal).
 */
int bt_expand_traces(int64_t n_pairs, const int64_t *trace_len, const int32_t *first,
                     const uint8_t *ops, const int64_t *ops_off, const int64_t *bands,
                     const int64_t *rows_off, int64_t *out);

/*
 * Test hook: the host-side plan of a resident packed batch (pair order in groups of 64 slots,
 * -1 = padding; groups and direction-table bytes per window) without any device work.  Pair k
 * has lengths off1[k+1]-off1[k] and off2[k+1]-off2[k].  `layout_hash_out` (optional) receives a
 * hash of the group descriptors, window records and scratch sizes.  Used by the CPU test-suite to
 * check the planner's invariants; not needed by callers.
 */
int bt_debug_plan(const BtParams *params, const int64_t *off1, const int64_t *off2, int64_t n_pairs,
                  int32_t score_only, int32_t waves_per_window, int32_t *order_out, int64_t order_cap,
                  int64_t *n_groups_out, int64_t *window_groups_out, int64_t *window_dir_bytes_out,
                  int64_t window_cap, int64_t *n_windows_out, int64_t *cells_out,
                  uint64_t *layout_hash_out);

/*
 * Test hook: the device planner's plan of the pairs (off1, off2) as one window: order_out receives
 * ceil(n_pairs / 64) * 64 slot -> pair ids (-1 = padding), group_nm_out (optional) nmax << 16 |
 * mmax per group, dir_off_out (optional) the first direction word of every group.  The GPU tests
 * compare it with the host planner (bt_debug_plan) and a numpy restatement.
 */
int bt_debug_device_plan(const int64_t *off1, const int64_t *off2, int64_t n_pairs, int32_t score_only,
                         int32_t affine, int32_t *order_out, uint32_t *group_nm_out, int64_t *dir_off_out);

/*
 * Roofline denominators measured on the current device: register-only loops of
 * the instructions the fill kernels are made of.
 *   kind 0: packed 16-bit DPX (VIADDMNMX.S16x2 / VIMNMX3.S16x2), result in
 *           lane-instructions per second (one lane-instruction updates two
 *           16-bit values)
 *   kind 1: 32-bit DPX (VIADDMNMX / VIMNMX3), lane-instructions per second
 *   kind 2: mixed ALU + FMA-pipe integer stream (VIADDMNMX.S16x2 + IMAD), lane-
 *           instructions per second - the issue-rate ceiling
 */
int bt_microbench(int kind, double *lane_instr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* B200ALIGN_H */
