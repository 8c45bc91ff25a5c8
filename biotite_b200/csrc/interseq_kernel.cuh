// Inter-sequence packed 16-bit kernels: the throughput path for batches of short
// pairs (BASELINE config 2: 10^6 pairs of ~300 aa, global affine, score + trace).
//
// Mapping.  Pairs are sorted by (len1, len2) and dealt to warps in groups of 64.
// Every lane owns TWO pairs, one in each 16-bit half of its registers, and sweeps
// their DP tables in stripes of R rows, column by column: the stripe's left
// neighbours (M, F1, H of the previous column) live in registers, the row above
// the stripe comes from a lane-interleaved row buffer, the substitution scores
// come from a bank-conflict-free shared-memory table (one private copy per lane),
// and the recurrence runs on the packed DPX instructions VIADD.16x2 / VIMNMX.S16x2
// / VIADDMNMX.S16x2.  There is no intra-warp communication.
//
// Recurrence (reference: src/rust/sequence/align/pairwise.rs:608-681).  The
// reference keeps three states M, G1, G2 per cell, with G1 fed from M and G1
// only.  The kernel carries M, F1 = G1 - open, F2 = G2 - open and
// H = max(M, G1, G2):
//     M  = H(i-1,j-1) + sim
//     F1 = max(M(i,j-1), F1(i,j-1) + ext)        o1 = [M(i,j-1) >= F1(i,j-1) + ext]
//     F2 = max(M(i-1,j), F2(i-1,j) + ext)        o2 = [M(i-1,j) >= F2(i-1,j) + ext]
//     T  = max(F1, F2)                           hG = [F1 >= F2]
//     H  = max(M, T + open)                      hM = [M >= T + open]
// The four predicates are exactly the reference's first-priority choices
// (cell.rs:219-254, :331-394: Match <- M, G1, G2 in that order; Gap1 <- M before
// G1; Gap2 <- M before G2) and are stored as one nibble per cell; the traceback
// kernel walks them.  Sixteen-bit lanes are used only when the host has proven
// that no finite value can leave the exact range (interseq_eligible).
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/b200align.h"
#include "common.cuh"
#include "pool.cuh"

namespace bt {

constexpr int IS_R = 16;            // rows per stripe
constexpr int IS_THREADS = 128;     // 4 warps per block
constexpr int IS_MAX_LEN = 8000;    // longest sequence routed here
constexpr int64_t IS_DIR_BUDGET = 40ll << 30;  // direction nibbles resident per window (bytes)
constexpr int IS_ROWW = 128;        // bytes per table entry row: 32 lanes x 4 bytes
#ifndef BT_L2PF
#define BT_L2PF 6
#endif
constexpr int IS_PREFETCH = BT_L2PF;  // columns of L2 prefetch distance for the row buffer
#ifndef BT_ACC_SPLIT
#define BT_ACC_SPLIT 2
#endif
#ifndef BT_PF_TRACED
#define BT_PF_TRACED 3
#endif
#ifndef BT_PF_SCORE
#define BT_PF_SCORE 2
#endif
constexpr int IS_WARPS_PER_SM = 9;  // affine kernels: one warp per block, 24 KiB of profile each
constexpr int IS_SMEM_BYTES = 226 * 1024;   // per SM: 227 KiB less the block's reserved KiB
constexpr int IS_SMALL_LETTERS = 20;        // the standard amino acids: 11 warps per SM instead of 9

// warps per SM whose stripe profiles of n_eff letters fit (one signed byte per row: 1 KiB per letter)
inline int interseq_profile_warps(int n_eff) { return std::max(1, std::min(16, IS_SMEM_BYTES / (2 * n_eff * 512))); }
constexpr int IS_SLACK = 32;        // columns of slack behind the row buffer / column codes: the
                                    // column loop reads ahead without bounds checks

// One group of 64 pairs handled by one warp.
struct WarpDesc {
    int64_t dir_off;      // first 32-bit word of the group's direction nibbles
    int64_t rowbuf_off;   // first entry of the group's row buffer
    int64_t rowcode_off;  // first uint16 of the group's lane-interleaved row codes
    int64_t colcode_off;  // first uint16 of the group's lane-interleaved column codes
    int32_t nmax, mmax;   // longest first / second sequence of the group
    int32_t stripes, pad;
};

struct InterseqConst {
    int32_t gap_open, gap_ext, neg;  // neg: the 16-bit "minus infinity"
    int32_t n_rows, n_cols;
    int32_t score_only;
    int32_t mode;                    // IS_GLOBAL / IS_SEMI / IS_LOCAL
    int32_t affine;                  // 0: linear gap penalty (gap_open)
    int32_t stats;                   // traceback emits GapStats instead of step bytes
    int32_t terminal;                // terminal gap columns count in the statistics
    int32_t dir_layout;              // 0: predicate nibbles (interseq_fill_kernel), 1: tag nibbles
                                     // (interseq_tag_kernel.cuh; scores scaled by 4, neg likewise)
};

// A window is a run of consecutive pairs (in input order) that is sorted, packed, filled and
// traced as one unit; its inputs and outputs are contiguous ranges of the batch arrays, so
// windows can be streamed (H2D, kernels, D2H) independently of each other.
struct Window {
    int64_t pair_begin, pair_end;
    int64_t group_begin, group_end;
    int64_t dir_words, rowbuf_entries, rowcodes, colcodes;  // scratch the window needs
};

// Scratch buffers of one in-flight window.
struct SlotBuffers {
    DevBuf rowcodes, colcodes, dirs, rowbuf, lastrow, lastcol;
    DevBuf ctl;   // 4 ints: largest column letter of the window, group counters of the persistent kernels
};

// Host-side plan of one batch.
struct InterseqPlan {
    int64_t n_pairs = 0, n_groups = 0;
    int64_t max_dir_words = 0, max_rowbuf = 0, max_rowcodes = 0, max_colcodes = 0;
    int warps_per_sm = IS_WARPS_PER_SM;   // resident warps of the fill kernel that will run (window sizing)
    InterseqConst C{};
    std::vector<int32_t> order;        // group slot -> pair id (-1 = padding), window by window
    std::vector<WarpDesc> desc;        // offsets are relative to the window's scratch
    std::vector<Window> windows;
    DevBuf d_order, d_desc;
    SlotBuffers slots[2];
};

struct InterseqIO {
    const void *codes1, *codes2;
    PairMap map;
    const int32_t *matrix;
    int32_t *score;
    int64_t *trace_len;
    int32_t *first;
    uint8_t *ops;
    int32_t *start;  // local: first optimum cell per pair (2 ints), scratch of the batch
    int *flag = nullptr;  // raised by the packer when a code lies outside the matrix
};

// ---------------------------------------------------------------------------------------
// eligibility: proven-exact 16-bit range (see DESIGN.md "16-bit safety")
// ---------------------------------------------------------------------------------------
// Host threads this process may use for planning: the machine's hardware threads divided by the
// processes sharing it (one per GPU under torchrun: LOCAL_WORLD_SIZE), B200ALIGN_THREADS overrides.
inline unsigned interseq_host_threads() {
    if (const char *t = getenv("B200ALIGN_THREADS")) return (unsigned)std::max(1, atoi(t));
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char *w = getenv("LOCAL_WORLD_SIZE")) hw = std::max(1u, hw / (unsigned)std::max(1, atoi(w)));
    return hw;
}

// One threaded pass over the pair lengths: everything routing and allocation need to know.
struct PairScan {
    bool in_range = true;      // 1 <= len <= IS_MAX_LEN for every sequence
    int64_t nmax = 0, mmax = 0;
    double total = 0.0, max_pair = 0.0, max_diag = 0.0;
};

inline PairScan interseq_scan(const PairMap &hm, int64_t n_pairs) {
    const unsigned hw = interseq_host_threads();
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(hw, 8u), n_pairs / 65536));
    std::vector<PairScan> part((size_t)n_threads);
    auto work = [&](int t) {
        PairScan r;
        const int64_t k0 = n_pairs * t / n_threads, k1 = n_pairs * (t + 1) / n_threads;
        for (int64_t k = k0; k < k1; k++) {
            const int64_t n = hm.len1(k), m = hm.len2(k);
            if (n < 1 || m < 1 || n > IS_MAX_LEN || m > IS_MAX_LEN) r.in_range = false;
            r.nmax = std::max(r.nmax, n); r.mmax = std::max(r.mmax, m);
            const double c = (double)n * (double)m;
            r.total += c; r.max_pair = std::max(r.max_pair, c); r.max_diag = std::max(r.max_diag, (double)(n + m));
        }
        part[(size_t)t] = r;
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < n_threads; t++) threads.emplace_back(work, t);
    work(0);
    for (std::thread &t : threads) t.join();
    PairScan r;
    for (const PairScan &q : part) {
        r.in_range = r.in_range && q.in_range;
        r.nmax = std::max(r.nmax, q.nmax); r.mmax = std::max(r.mmax, q.mmax);
        r.total += q.total; r.max_pair = std::max(r.max_pair, q.max_pair); r.max_diag = std::max(r.max_diag, q.max_diag);
    }
    return r;
}

inline bool interseq_eligible(const BtParams &P, const PairMap &hm, int64_t n_pairs, int32_t min_score,
                              int32_t max_score, const PairScan *scan_in = nullptr) {
    if (P.banded || !P.use_matrix) return false;
    if (!P.affine && !P.local && !P.terminal_penalty) return false;  // linear semi-global: general kernel
    if (P.code_bytes != 1 || n_pairs < 64) return false;
    if (P.n_rows > 255 || P.n_cols > 255) return false;
    if (P.affine) {
        // stripe profile of signed bytes: 1 KiB per column letter and warp
        if (min_score < -128 || max_score > 127 || P.n_cols > 128) return false;
    } else if ((int64_t)P.n_rows * P.n_cols * IS_ROWW > 96 * 1024) return false;  // per-lane score table
    const PairScan scan = scan_in ? *scan_in : interseq_scan(hm, n_pairs);
    if (!scan.in_range) return false;
    const int64_t nmax = scan.nmax, mmax = scan.mmax;
    // Routing by a simple cost model: the packed kernel runs one group of 64 pairs per warp,
    // so a batch with few groups leaves most of the chip idle; the general kernel runs one
    // block per pair and is bound by ~1.2 us per anti-diagonal.  Measured rates on B200:
    // ~1.0e9 cells/s per resident warp for the packed kernel, ~9e10 cells/s for a full grid of
    // the general kernel.
    // B200ALIGN_ROUTE=packed|general overrides the cost model (tests exercise both routes on the
    // same inputs); the exactness conditions below always apply.
    const char *forced = getenv("B200ALIGN_ROUTE");
    if (forced && strcmp(forced, "general") == 0) return false;
    if (!(forced && strcmp(forced, "packed") == 0)) {
        const double t_packed = std::max(64.0 * scan.max_pair / 1.0e9, scan.total / 1.8e12);
        const double waves = std::ceil((double)n_pairs / 888.0);
        const double t_general = std::max(scan.total / 9.0e10, scan.max_diag * 1.2e-6 * waves);
        if (t_general < t_packed) return false;
    }
    const int64_t open = P.gap_open, ext = P.affine ? P.gap_ext : P.gap_open;
    const int64_t mn = std::min<int64_t>(min_score, 0), mx = std::max<int64_t>(max_score, 0);
    const int64_t neg = -32768 - open - ext - mn;                         // one open/ext and one sim may be added
    const int64_t lowest = 3 * open + (nmax + mmax + 2 + IS_R) * ext + 2 * mn;  // lowest finite value
    const int64_t highest = std::min(nmax, mmax) * mx - open + mx;        // highest finite value (F = G - open)
    return lowest > neg + 1 && highest < 32767;
}

// ---------------------------------------------------------------------------------------
// batch packer: lane-interleaved codes of each group
// ---------------------------------------------------------------------------------------
// Codes outside the substitution matrix (the reference would panic, scoring.rs:63) raise `flag`
// and are clamped to the last letter, so that the fill kernels' profile / table lookups stay
// inside their shared-memory windows whatever the caller sent; the host turns the flag into
// BT_ERR_INVALID_CODE after the batch.
__global__ void interseq_pack_kernel(const uint8_t *codes1, const uint8_t *codes2, const PairMap map,
                                     const int32_t *order, const WarpDesc *desc,
                                     uint16_t *rowcodes, uint16_t *colcodes, int64_t n_groups,
                                     uint32_t n_rows, uint32_t n_cols, int *flag, int *maxcol) {
    const int64_t g = blockIdx.x;
    if (g >= n_groups) return;
    const WarpDesc d = desc[g];
    const int rows = d.stripes * IS_R;
    // thread t serves lane t & 31 and walks positions with stride blockDim/32
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5, nsub = blockDim.x >> 5;
    const int32_t p_lo = order[g * 64 + 2 * lane], p_hi = order[g * 64 + 2 * lane + 1];
    int64_t a_lo = 0, a_hi = 0, b_lo = 0, b_hi = 0;
    int n_lo = 0, n_hi = 0, m_lo = 0, m_hi = 0;
    if (p_lo >= 0) { a_lo = map.beg1(p_lo); n_lo = (int)map.len1(p_lo); b_lo = map.beg2(p_lo); m_lo = (int)map.len2(p_lo); }
    if (p_hi >= 0) { a_hi = map.beg1(p_hi); n_hi = (int)map.len1(p_hi); b_hi = map.beg2(p_hi); m_hi = (int)map.len2(p_hi); }
    bool bad = false;
    for (int i = sub; i < rows; i += nsub) {
        uint32_t lo = i < n_lo ? codes1[a_lo + i] : 0, hi = i < n_hi ? codes1[a_hi + i] : 0;
        if (lo >= n_rows) { bad = true; lo = n_rows - 1; }
        if (hi >= n_rows) { bad = true; hi = n_rows - 1; }
        rowcodes[d.rowcode_off + (int64_t)i * 32 + lane] = (uint16_t)(lo | (hi << 8));
    }
    uint32_t top = 0;
    for (int j = sub; j < d.mmax; j += nsub) {
        uint32_t lo = j < m_lo ? codes2[b_lo + j] : 0, hi = j < m_hi ? codes2[b_hi + j] : 0;
        if (lo >= n_cols) { bad = true; lo = n_cols - 1; }
        if (hi >= n_cols) { bad = true; hi = n_cols - 1; }
        top = max(top, max(lo, hi));
        colcodes[d.colcode_off + (int64_t)j * 32 + lane] = (uint16_t)(lo | (hi << 8));
    }
    if (bad && flag) *flag = 1;
    if (maxcol) {
        // largest column letter of the window: decides how many letters the stripe profiles hold
        top = __reduce_max_sync(0xffffffffu, top);
        if (lane == 0 && top > 0) atomicMax(maxcol, (int)top);
    }
}

// ---------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------

// max of two packed s16 pairs; adds `c_lo` / `c_hi` to `acc` where the FIRST operand attains
// the maximum of the low / high half (a >= b).  ptxas fuses the compare pattern into one
// VIMNMX.S16x2 with two predicate outputs.
#define BT_MAX_PRED(out, a, b, acc_lo, acc_hi, c)                                        \
    asm("{\n\t.reg .pred pl, ph;\n\t.reg .s16 x0, x1, y0, y1;\n\t"                      \
        "max.s16x2 %0, %3, %4;\n\t"                                                      \
        "mov.b32 {x0, x1}, %0;\n\tmov.b32 {y0, y1}, %3;\n\t"                             \
        "setp.eq.s16 pl, x0, y0;\n\tsetp.eq.s16 ph, x1, y1;\n\t"                         \
        "@pl add.u32 %1, %1, %5;\n\t@ph add.u32 %2, %2, %5;\n\t}"                        \
        : "=&r"(out), "+r"(acc_lo), "+r"(acc_hi)                                          \
        : "r"(a), "r"(b), "r"(c))

enum : int { IS_GLOBAL = 0, IS_SEMI = 1, IS_LOCAL = 2 };

// v[idx] for a run-time idx in 0..15 as a select tree over the bits of idx (the obvious chain
// of 16 compares makes ptxas keep 32 loop-invariant predicates alive across the column loop,
// which starves the predicate-producing VIMNMX instructions)
__device__ __forceinline__ uint32_t bt_pick16(const uint32_t (&v)[16], int idx) {
    uint32_t a[8], b[4], c[2];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = (idx & 1) ? v[2 * k + 1] : v[2 * k];
#pragma unroll
    for (int k = 0; k < 4; k++) b[k] = (idx & 2) ? a[2 * k + 1] : a[2 * k];
#pragma unroll
    for (int k = 0; k < 2; k++) c[k] = (idx & 4) ? b[2 * k + 1] : b[2 * k];
    return (idx & 8) ? c[1] : c[0];
}

// Per-window scratch and per-batch result pointers handed to the kernels.
struct InterseqArgs {
    const WarpDesc *desc;
    const int32_t *order;
    PairMap map;
    const int32_t *matrix;
    const uint16_t *rowcodes, *colcodes;
    uint32_t *rowbuf;     // 3 planes [j][lane]: M, F2, H of the stripe's last row
    uint32_t *lastrow;    // semi-global: M of the pair's last row, 2 planes [j][lane] (low / high pair)
    uint32_t *lastcol;    // semi-global: M of the pair's last column, 2 planes [row][lane] (low / high pair)
    uint32_t *dirs;
    int32_t *score;
    int32_t *start;       // local: table cell (i, j) of the first optimum, 2 per pair
    int64_t n_groups;
};

// MODE selects the boundary conditions of pairwise.rs: IS_GLOBAL = terminal gaps penalised
// (:707-780), IS_SEMI = free terminal gaps (the waived last row / column is resolved by the
// traceback kernel from the captured border M values, see there), IS_LOCAL = Smith-Waterman
// (M clamped at 0, :659-666; clamping the gap states is not needed for scores or for the
// first-priority trace because M >= 0 always wins a tie against a non-positive gap state).
// One group of 64 pairs per warp, `blockDim.x / 32` warps per block (one block per SM slot: the
// 1 KiB of shared memory the system reserves per block is paid once, which is what lets 11 stripe
// profiles of 20 letters fit an SM instead of 9 of 24).  The profile holds the first `n_eff` column
// letters only; `ctl[0]` carries the largest column letter the packer saw in this window, so a
// variant built for fewer letters than the matrix has steps aside when a rarer letter turns up
// and the variant for all letters, launched right after it, does the work (as in
// interseq_tag_kernel.cuh).
template <int MODE, bool TRACED, int PF, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
interseq_fill_kernel(const InterseqConst C, const InterseqArgs A, const int n_eff, const int small_letters,
                     int *const ctl, const int counter_slot) {
    // Shared memory per warp: the stripe profile prof[half][b][lane][16 rows] (one signed byte per
    // row: the score of the stripe's row letter against column letter b), rebuilt for every
    // stripe from the matrix in global memory (L1-resident).
    extern __shared__ uint4 smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const int maxcol = ctl[0];
        if (n_eff < C.n_cols ? maxcol >= n_eff : (small_letters > 0 && maxcol < small_letters)) return;
    }
    uint4 *const smem_prof = smem_all + (size_t)warp * 2 * n_eff * 32;
  // persistent: one block per SM, every warp takes the next group of 64 pairs from a counter until
  // none is left (no partial last wave, no block waiting for its slowest warp)
  for (;;) {
    int64_t g = 0;
    if (lane == 0) g = atomicAdd(ctl + counter_slot, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= A.n_groups) break;
    const WarpDesc d = A.desc[g];
    const int32_t p_lo = A.order[g * 64 + 2 * lane], p_hi = A.order[g * 64 + 2 * lane + 1];
    int n_lo = 0, n_hi = 0, m_lo = 0, m_hi = 0;
    if (p_lo >= 0) { n_lo = (int)A.map.len1(p_lo); m_lo = (int)A.map.len2(p_lo); }
    if (p_hi >= 0) { n_hi = (int)A.map.len1(p_hi); m_hi = (int)A.map.len2(p_hi); }

    const uint32_t ext2 = ((uint32_t)C.gap_ext & 0xffffu) * 0x10001u;
    const uint32_t open2 = ((uint32_t)C.gap_open & 0xffffu) * 0x10001u;
    const uint32_t neg2 = ((uint32_t)C.neg & 0xffffu) * 0x10001u;
    // packed zero the compiler cannot see through: with a literal 0 ptxas rebuilds the operand
    // (PRMT RZ) for every VIADDMNMX / VIMNMX3 that clamps at zero, one ALU instruction per row
    const uint32_t zero2 = (uint32_t)C.n_cols >> 31;
    // this lane's 16 profile bytes of (half, letter b): pbase + (half * n_eff + b) * 512
    const uint32_t pbase = (uint32_t)__cvta_generic_to_shared(smem_prof) + lane * 16;
    const uint32_t pbase_hi = pbase + (uint32_t)n_eff * 512u;

    const uint16_t *rc = A.rowcodes + d.rowcode_off + lane;
    const uint16_t *cc = A.colcodes + d.colcode_off + lane;
    // row buffer: [j][plane][lane].  Traced: planes M, F2, H of the row above the current stripe
    // (the first row's o2 bit needs M and F2).  Score only: two planes - the F2 of the stripe's
    // first row, already formed by the row above (max(M, F2 + ext)), and H.
    constexpr int RBS = TRACED ? 96 : 64;          // words per column
    constexpr int RB_H = TRACED ? 64 : 32;         // plane of H
    uint32_t *const rb = A.rowbuf + d.rowbuf_off * (TRACED ? 3 : 2) + lane;
    uint4 *dir = TRACED ? (uint4 *)(A.dirs + d.dir_off) + lane : nullptr;
    // two planes each (low / high pair): their last rows / columns are captured at different times
    uint32_t *lastrow_lo = MODE == IS_SEMI ? A.lastrow + d.rowbuf_off * 2 + lane : nullptr;
    uint32_t *lastrow_hi = MODE == IS_SEMI ? lastrow_lo + (int64_t)d.mmax * 32 : nullptr;
    uint32_t *lastcol_lo = MODE == IS_SEMI ? A.lastcol + d.rowcode_off * 2 + lane : nullptr;
    uint32_t *lastcol_hi = MODE == IS_SEMI ? lastcol_lo + (int64_t)d.stripes * IS_R * 32 : nullptr;

    int32_t best_lo = 0, best_hi = 0;          // optimum (local: running maximum of M)
    int32_t bidx_lo = 0, bidx_hi = 0;          // local: row-major index i*(m+1)+j of the first optimum
    int32_t brow_lo = 0, brow_hi = 0;          // ... and its row
    uint32_t best2 = 0;                        // local score-only: packed running maximum
    // Row 0 of the table (pairwise.rs:747-780: M = G2 = -inf, H = G1 = open + (j-1)*ext, or 0
    // without terminal penalty / local) goes through the row buffer like every other stripe
    // boundary, so that the column loop below is the same for all stripes.
    {
        uint32_t h_row0 = MODE == IS_GLOBAL ? open2 : 0u;
        uint32_t *w = rb;
        for (int j = 0; j < d.mmax; j++, w += RBS) {
            w[0] = neg2; if (TRACED) w[32] = neg2;   // score only: F2 of row 1 = max(-inf, -inf + ext) = -inf
            w[RB_H] = h_row0;
            if (MODE == IS_GLOBAL) h_row0 = __vadd2(h_row0, ext2);
        }
    }
    for (int s = 0; s < d.stripes; s++) {
        const int r0 = s * IS_R;  // rows r0+1 .. r0+R of the DP table
        // column step at which this stripe holds the end cell (n, m) of the low / high pair
        const int j_lo = (n_lo - 1) / IS_R == s ? m_lo - 1 : -1;
        const int j_hi = (n_hi - 1) / IS_R == s ? m_hi - 1 : -1;
        // row of this stripe that is the last row of the low / high pair (or -1)
        const int rl_lo = (n_lo - 1) / IS_R == s ? (n_lo - 1) % IS_R : -1;
        const int rl_hi = (n_hi - 1) / IS_R == s ? (n_hi - 1) % IS_R : -1;
        const bool rows_valid = r0 + IS_R <= n_lo && r0 + IS_R <= n_hi;  // no padding rows in this stripe
        uint32_t Ml[IS_R], F1l[IS_R], Hl[IS_R];
        {
            // profile of this stripe: byte r of prof[half][b][lane] = matrix[row letter r][b]
            uint32_t a_lo[IS_R], a_hi[IS_R];
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                const uint32_t v = rc[(r0 + r) * 32];
                a_lo[r] = (v & 0xffu) * (uint32_t)C.n_cols;   // row offset in the matrix
                a_hi[r] = (v >> 8) * (uint32_t)C.n_cols;
            }
            __syncwarp();
            for (int b = 0; b < n_eff; b++) {
                const int32_t *col = A.matrix + b;
                uint32_t w_lo[4], w_hi[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    w_lo[q] = w_hi[q] = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        w_lo[q] |= ((uint32_t)__ldg(col + a_lo[4 * q + k]) & 0xffu) << (8 * k);
                        w_hi[q] |= ((uint32_t)__ldg(col + a_hi[4 * q + k]) & 0xffu) << (8 * k);
                    }
                }
                smem_prof[b * 32 + lane] = make_uint4(w_lo[0], w_lo[1], w_lo[2], w_lo[3]);
                smem_prof[(n_eff + b) * 32 + lane] = make_uint4(w_hi[0], w_hi[1], w_hi[2], w_hi[3]);
            }
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < IS_R; r++) {
            // column 0 (pairwise.rs:707-745): M = G1 = -inf, H = G2 = open + (i-1)*ext
            // (semi-global / local: H = 0)
            Ml[r] = neg2;
            F1l[r] = neg2;
            Hl[r] = MODE == IS_GLOBAL ? ((uint32_t)(C.gap_open + (r0 + r) * C.gap_ext) & 0xffffu) * 0x10001u : 0u;
        }
        // H(r0, 0): the origin for the first stripe, else column 0 of the row above
        uint32_t hu_prev = (s == 0 || MODE != IS_GLOBAL)
                               ? 0u : ((uint32_t)(C.gap_open + (r0 - 1) * C.gap_ext) & 0xffffu) * 0x10001u;
        // The row above the stripe and the column letters are fetched PF columns ahead (their
        // L2 / HBM latency hides behind PF x 16 rows of work); reads past the group's last
        // column land in the slack behind the buffers and are never used.
        const uint16_t *ccp = cc;
        uint32_t *rbp = rb;
        uint4 *dirp = TRACED ? dir + (int64_t)s * d.mmax * 32 : nullptr;
        uint32_t q_c[PF], q_m[PF], q_f[PF], q_h[PF];
#pragma unroll
        for (int k = 0; k < PF; k++) {
            q_c[k] = ccp[k * 32]; q_m[k] = rbp[k * RBS]; q_f[k] = TRACED ? rbp[k * RBS + 32] : 0u; q_h[k] = rbp[k * RBS + RB_H];
        }
        // the column loop is unrolled PF times so that the look-ahead registers need no rotation
        // (a register move would wait for the load it forwards)
        for (int j0 = 0; j0 < d.mmax; j0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) {
            const int j = j0 + u;
            if (u > 0 && j >= d.mmax) break;
            const uint32_t v = q_c[u];
            // traced: M and F2 of the row above; score only: plane 0 is this stripe's first-row F2 itself
            uint32_t up_m = TRACED ? q_m[u] : 0u, up_f2 = TRACED ? q_f[u] : q_m[u];
            const uint32_t hu = q_h[u];
            q_c[u] = ccp[PF * 32];
            q_m[u] = rbp[PF * RBS]; if (TRACED) q_f[u] = rbp[PF * RBS + 32]; q_h[u] = rbp[PF * RBS + RB_H];
            if (IS_PREFETCH > 0) {
                // pull the row buffer towards L2 well ahead of the register prefetch
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + (PF + IS_PREFETCH) * RBS));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + (PF + IS_PREFETCH) * RBS + 32));
                if (TRACED) asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + (PF + IS_PREFETCH) * RBS + 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ccp + (PF + IS_PREFETCH) * 32));
            }
            // the 16 + 16 profile bytes of this column's letters
            uint32_t pl[4], ph[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(pl[0]), "=r"(pl[1]), "=r"(pl[2]), "=r"(pl[3]) : "r"(pbase + (v & 0xffu) * 512u));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(ph[0]), "=r"(ph[1]), "=r"(ph[2]), "=r"(ph[3]) : "r"(pbase_hi + (v >> 8) * 512u));
            // Rows are software-pipelined so that every state array is updated in place:
            // F1 and M of row r+1 are formed (from the previous column's M and H) before
            // H of row r overwrites Hl[r].
            // direction nibbles: lo rows 0-7, lo rows 8-15, hi rows 0-7, hi rows 8-15; one set of
            // accumulators per predicate (BT_ACC_SPLIT) shortens the chains of dependent adds
            uint32_t acc[4 * BT_ACC_SPLIT] = {};
            uint32_t sim[IS_R];
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                // sign-extend byte r of the low / high pair's profile into the two 16-bit halves
                constexpr uint32_t k = 0;
                const uint32_t sel = (uint32_t)(r & 3) * 0x1111u + 0xC480u;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sim[r]) : "r"(pl[r >> 2]), "r"(ph[r >> 2]), "r"(sel));
                (void)k;
            }
#define BT_ACC_LO(r, b) acc[((b) % BT_ACC_SPLIT) * 4 + ((r) >> 3)]
#define BT_ACC_HI(r, b) acc[((b) % BT_ACC_SPLIT) * 4 + 2 + ((r) >> 3)]
#define BT_BIT(r, b) ((1u << (b)) << (4 * ((r) & 7)))
#define BT_ROW_F1_M(r, diag_h)                                                               \
    {                                                                                        \
        const uint32_t fe = __vadd2(F1l[r], ext2);                                           \
        if (TRACED) { BT_MAX_PRED(F1l[r], Ml[r], fe, BT_ACC_LO(r, 2), BT_ACC_HI(r, 2), BT_BIT(r, 2)); } \
        else F1l[r] = __vmaxs2(Ml[r], fe);                                                   \
        Ml[r] = MODE == IS_LOCAL ? __viaddmax_s16x2(diag_h, sim[r], zero2) : __vadd2(diag_h, sim[r]); \
    }
            BT_ROW_F1_M(0, hu_prev);
            hu_prev = hu;
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                if (r + 1 < IS_R) BT_ROW_F1_M(r + 1, Hl[r]);
                const uint32_t fv = __vadd2(up_f2, ext2);
                if (TRACED) {
                    uint32_t t;
                    BT_MAX_PRED(up_f2, up_m, fv, BT_ACC_LO(r, 3), BT_ACC_HI(r, 3), BT_BIT(r, 3));     // o2
                    BT_MAX_PRED(t, F1l[r], up_f2, BT_ACC_LO(r, 1), BT_ACC_HI(r, 1), BT_BIT(r, 1));   // hG
                    const uint32_t to = __vadd2(t, open2);
                    BT_MAX_PRED(Hl[r], Ml[r], to, BT_ACC_LO(r, 0), BT_ACC_HI(r, 0), BT_BIT(r, 0));   // hM
                } else {
                    if (r > 0) up_f2 = __vmaxs2(up_m, fv);
                    Hl[r] = __viaddmax_s16x2(__vmaxs2(F1l[r], up_f2), open2, Ml[r]);
                }
                up_m = Ml[r];
            }
            if (TRACED) { rbp[0] = up_m; rbp[32] = up_f2; rbp[64] = Hl[IS_R - 1]; }
            else { rbp[0] = __viaddmax_s16x2(up_f2, ext2, up_m); rbp[32] = Hl[IS_R - 1]; }   // F2 of the row below
            if (TRACED) {
#pragma unroll
                for (int k = 1; k < BT_ACC_SPLIT; k++) {
                    acc[0] |= acc[4 * k]; acc[1] |= acc[4 * k + 1]; acc[2] |= acc[4 * k + 2]; acc[3] |= acc[4 * k + 3];
                }
                *dirp = make_uint4(acc[0], acc[1], acc[2], acc[3]);
                dirp += 32;
            }

            if (MODE == IS_GLOBAL) {
                // optimum: H(n, m) (pairwise.rs:866-879)
                if (j == j_lo || j == j_hi) {
                    if (j == j_lo) best_lo = (int16_t)(bt_pick16(Hl, rl_lo) & 0xffffu);
                    if (j == j_hi) best_hi = (int16_t)(bt_pick16(Hl, rl_hi) >> 16);
                }
            } else if (MODE == IS_SEMI) {
                // border capture: M along the pair's last row and last column
                if (rl_lo >= 0 || rl_hi >= 0) {
                    uint32_t w_lo = 0, w_hi = 0;
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        if (r == rl_lo) w_lo = Ml[r];
                        if (r == rl_hi) w_hi = Ml[r];
                    }
                    if (rl_lo >= 0) lastrow_lo[(int64_t)j * 32] = w_lo;
                    if (rl_hi >= 0) lastrow_hi[(int64_t)j * 32] = w_hi;
                }
                if (j == m_lo - 1) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) lastcol_lo[(int64_t)(r0 + r) * 32] = Ml[r];
                }
                if (j == m_hi - 1) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) lastcol_hi[(int64_t)(r0 + r) * 32] = Ml[r];
                }
            } else {
                // local: running maximum of M over the cells inside the table, first in row-major
                // order (pairwise.rs:849-865).  Fast path: one packed max per row; the scalar
                // path runs where padding must be masked or (traced) a new optimum may appear.
                bool slow = false;
                if (!TRACED) {
                    // score only: packed running maximum; padding rows (last stripe) and padding
                    // columns are masked to 0, which never beats the running maximum (>= 0)
                    uint32_t cm = zero2;
                    if (rows_valid) {
#pragma unroll
                        for (int r = 0; r < IS_R; r += 2) cm = __vimax3_s16x2(cm, Ml[r], Ml[r + 1]);
                    } else {
#pragma unroll
                        for (int r = 0; r < IS_R; r++) {
                            const uint32_t rmask = (r0 + r < n_lo ? 0xffffu : 0u) | (r0 + r < n_hi ? 0xffff0000u : 0u);
                            cm = __vmaxs2(cm, Ml[r] & rmask);
                        }
                    }
                    const uint32_t cmask = (j < m_lo ? 0xffffu : 0u) | (j < m_hi ? 0xffff0000u : 0u);
                    best2 = __vmaxs2(best2, cm & cmask);
                } else {
                    uint32_t cm = Ml[0];
#pragma unroll
                    for (int r = 1; r < IS_R; r++) cm = __vmaxs2(cm, Ml[r]);
                    const bool inside = rows_valid && j < m_lo && j < m_hi;
                    slow = !inside;
                    if (inside) {
                        const int c_lo = (int16_t)(cm & 0xffffu), c_hi = (int16_t)(cm >> 16);
                        // a new optimum, or a tie that could precede the current first optimum in
                        // row-major order (only possible while that optimum lies in this stripe)
                        slow = c_lo > best_lo || c_hi > best_hi ||
                               (c_lo == best_lo && c_lo > 0 && brow_lo > r0 + 1) ||
                               (c_hi == best_hi && c_hi > 0 && brow_hi > r0 + 1);
                    }
                }
                if (slow) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        const int row = r0 + r + 1, col = j + 1;
                        const int v_lo = (int16_t)(Ml[r] & 0xffffu), v_hi = (int16_t)(Ml[r] >> 16);
                        if (row <= n_lo && col <= m_lo) {
                            const int idx = row * (m_lo + 1) + col;
                            if (v_lo > best_lo || (v_lo == best_lo && v_lo > 0 && idx < bidx_lo)) { best_lo = v_lo; bidx_lo = idx; brow_lo = row; }
                        }
                        if (row <= n_hi && col <= m_hi) {
                            const int idx = row * (m_hi + 1) + col;
                            if (v_hi > best_hi || (v_hi == best_hi && v_hi > 0 && idx < bidx_hi)) { best_hi = v_hi; bidx_hi = idx; brow_hi = row; }
                        }
                    }
                }
            }
            ccp += 32; rbp += RBS;
        }
        }
    }
    if (MODE == IS_LOCAL && !TRACED) {
        best_lo = max(best_lo, (int)(int16_t)(best2 & 0xffffu));
        best_hi = max(best_hi, (int)(int16_t)(best2 >> 16));
    }
    if (MODE != IS_SEMI) {  // semi-global scores come from the border pass
        if (p_lo >= 0) A.score[p_lo] = best_lo;
        if (p_hi >= 0) A.score[p_hi] = best_hi;
    }
    if (MODE == IS_LOCAL && TRACED) {
        if (p_lo >= 0) { A.start[2 * p_lo] = bidx_lo / (m_lo + 1); A.start[2 * p_lo + 1] = bidx_lo % (m_lo + 1); }
        if (p_hi >= 0) { A.start[2 * p_hi] = bidx_hi / (m_hi + 1); A.start[2 * p_hi + 1] = bidx_hi % (m_hi + 1); }
    }
    __syncwarp();   // the next group's profile overwrites this one's
  }
}

// ---------------------------------------------------------------------------------------
// traceback over the nibbles: one thread per pair, first-priority steps only
// (trace.rs:59-90 with max_number = 1; step order cell.rs:219-254)
//
// Semi-global (terminal_penalty = false): the reference drops the gap penalties in the last
// row and column (pairwise.rs:794-843), values that only flow along that border to the end
// cell.  With M(n, .) and M(., m) captured by the fill kernel the border is resolved here:
//   G1(n, m) = max_{j<m} M(n, j),   G2(n, m) = max_{i<n} M(i, m),
// the optimum is the first of M(n,m), G1(n,m), G2(n,m) attaining the maximum, and a start in
// G1 (G2) walks the last row (column) back to the right-most (bottom-most) maximum, because
// the reference prefers the M -> gap transition on ties (cell.rs:370-391).
// Local: the walk starts at the first optimum in state M and carries the running score; it
// ends at the first M state whose value is 0 (direction bits cleared, pairwise.rs:661-666).
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void interseq_traceback_kernel(const InterseqConst C, const InterseqArgs A,
                                          const uint8_t *__restrict__ codes1,
                                          const uint8_t *__restrict__ codes2, int64_t *__restrict__ trace_len,
                                          int32_t *__restrict__ first, uint8_t *__restrict__ ops,
                                          int64_t n_slots) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int32_t p = A.order[slot];
    if (p < 0) return;
    const int64_t g = slot >> 6;
    const int lane = (int)(slot & 63) >> 1, half = (int)(slot & 1);
    const WarpDesc d = A.desc[g];
    const int64_t o1 = A.map.beg1(p), o2 = A.map.beg2(p);
    const int n = (int)A.map.len1(p), m = (int)A.map.len2(p);
    const bool traced = !C.score_only;
    const bool tags = C.dir_layout != 0;
    const uint32_t *base = A.dirs + d.dir_off + lane * 4 + (tags ? 0 : half * 2);
    // score-only and statistics batches carry no step bytes
    uint8_t *out = (traced && !C.stats) ? ops + A.map.ops(p) : nullptr;
    GapStats gs(C.terminal != 0);
    // step bytes are written four at a time once the output position is word-aligned (one
    // store transaction instead of four; the unaligned head and the tail go out as bytes)
    uint32_t ew = 0;
    int enb = 0;
    const int ealign = (int)((uintptr_t)out & 3u);
    auto emit = [&](int op, int &len) {
        if (C.stats) gs.push(op);
        else if (enb == 0 && ((ealign + len) & 3) != 0) out[len] = (uint8_t)op;
        else {
            ew |= (uint32_t)op << (8 * enb);
            if (++enb == 4) { *(uint32_t *)(out + len - 3) = ew; ew = 0; enb = 0; }
        }
        len++;
    };
    auto emit_flush = [&](int len) {
        for (int k = 0; k < enb; k++) out[len - enb + k] = (uint8_t)(ew >> (8 * k));
    };
    auto nibble = [&](int i, int j) -> uint32_t {  // 1-based table cell (i, j)
        const int rg = i - 1, s = rg / IS_R, r = rg % IS_R;
        if (tags) {
            // tag layout: word r / 4, low pair in bits 0-15, high pair in bits 16-31; bits 0-1 of
            // a nibble name the state holding H (2 Match, 1 Gap1, 0 Gap2) -> hM / hG of this walk
            const uint32_t w = base[((int64_t)s * d.mmax + (j - 1)) * 128 + (r >> 2)];
            const uint32_t nb = (w >> (4 * (r & 3) + 16 * half)) & 0xfu;
            const uint32_t t = nb & 3u;
            return (nb & 0xcu) | (t == 2u ? 1u : (t == 1u ? 2u : 0u));
        }
        const uint32_t w = base[((int64_t)s * d.mmax + (j - 1)) * 128 + (r >> 3)];
        return (w >> (4 * (r & 7))) & 0xfu;
    };
    auto h_state = [&](int i, int j) -> int {  // state with the best score at cell (i, j), priority M, G1, G2
        if (i == 0) return j == 0 ? ST_MATCH : ST_GAP1;
        if (j == 0) return ST_GAP2;
        const uint32_t nb = nibble(i, j);
        return (nb & 1u) ? ST_MATCH : ((nb & 2u) ? ST_GAP1 : ST_GAP2);
    };
    int i = n, j = m, len = 0, st = ST_MATCH;
    int fi = 0, fj = 0;
    int32_t value = 0;
    if (MODE == IS_GLOBAL) {
        st = h_state(i, j);
    } else if (MODE == IS_SEMI) {
        const uint32_t *lr = A.lastrow + d.rowbuf_off * 2 + lane + (half ? (int64_t)d.mmax * 32 : 0);
        const uint32_t *lc = A.lastcol + d.rowcode_off * 2 + lane + (half ? (int64_t)d.stripes * IS_R * 32 : 0);
        // the tag kernel captures 4 M + 1
        const int sh = tags ? 2 : 0;
        auto m_row = [&](int jj) -> int { const uint32_t w = lr[(int64_t)(jj - 1) * 32]; return (int)(int16_t)(half ? w >> 16 : w & 0xffffu) >> sh; };
        auto m_col = [&](int ii) -> int { const uint32_t w = lc[(int64_t)(ii - 1) * 32]; return (int)(int16_t)(half ? w >> 16 : w & 0xffffu) >> sh; };
        const int mc = m_row(m);
        int g1 = C.neg, p1 = 0, g2 = C.neg, p2 = 0;
        for (int jj = 1; jj < m; jj++) { const int v = m_row(jj); if (v >= g1) { g1 = v; p1 = jj; } }
        for (int ii = 1; ii < n; ii++) { const int v = m_col(ii); if (v >= g2) { g2 = v; p2 = ii; } }
        const int best = max(mc, max(g1, g2));
        A.score[p] = best;
        if (traced) {
            fi = n; fj = m;
            if (mc == best) { st = ST_MATCH; }
            else if (g1 == best) { while (j > p1) { fi = i; fj = j; emit(OP_TO1, len); j--; } st = ST_MATCH; }
            else { while (i > p2) { fi = i; fj = j; emit(OP_TO2, len); i--; } st = ST_MATCH; }
        }
    } else {
        i = A.start[2 * p]; j = A.start[2 * p + 1];
        value = A.score[p];
    }
    if (!traced) return;
    const int32_t *mat = A.matrix;
    while (i > 0 || j > 0) {
        if (MODE == IS_LOCAL && (i == 0 || j == 0)) break;         // boundary cells have no direction
        if (i == 0) { fi = i; fj = j; emit(OP_TO1, len); j--; continue; }   // first row: leading gap in sequence 1
        if (j == 0) { fi = i; fj = j; emit(OP_TO2, len); i--; continue; }   // first column
        if (st == ST_MATCH) {
            if (MODE == IS_LOCAL) {
                if (value <= 0) break;                              // M clamped to 0: trace ends before this cell
                value -= mat[min((int)codes1[o1 + i - 1], C.n_rows - 1) * C.n_cols + min((int)codes2[o2 + j - 1], C.n_cols - 1)];
            }
            fi = i; fj = j;
            emit(OP_DIAG, len);
            i--; j--;
            st = h_state(i, j);
        } else if (st == ST_GAP1) {
            const uint32_t nb = nibble(i, j);
            fi = i; fj = j;
            emit(OP_TO1, len);
            j--;
            const bool opened = (nb & 4u) != 0;
            st = opened ? ST_MATCH : ST_GAP1;
            if (MODE == IS_LOCAL) value -= opened ? C.gap_open : C.gap_ext;
        } else {
            const uint32_t nb = nibble(i, j);
            fi = i; fj = j;
            emit(OP_TO2, len);
            i--;
            const bool opened = (nb & 8u) != 0;
            st = opened ? ST_MATCH : ST_GAP2;
            if (MODE == IS_LOCAL) value -= opened ? C.gap_open : C.gap_ext;
        }
    }
    if (!C.stats && out) emit_flush(len);
    trace_len[p] = len;
    if (C.stats) { gs.finish(); fi = gs.opens; fj = gs.exts; }   // statistics travel in `first`
    first[2 * p] = fi;
    first[2 * p + 1] = fj;
}

// ---------------------------------------------------------------------------------------
// linear gap penalty (reference: pairwise.rs:378-415): one state per cell,
//     S = max(S(i-1,j-1) + sim, S(i,j-1) + g, S(i-1,j) + g),
// first-priority step Diag, then To1 (left), then To2 (up) (cell.rs:125-138, :291-323).
// Two predicates per cell: d = [diag >= max(left, up) + g], l = [left >= up].
// IS_GLOBAL (terminal gaps penalised) and IS_LOCAL; semi-global linear batches use the
// general kernel.  Same mapping, buffers and row software-pipelining as the affine kernel;
// the row buffer has one plane (S of the stripe's last row) and there is no -inf state.
// ---------------------------------------------------------------------------------------
template <int MODE, bool TRACED>
__global__ void __launch_bounds__(IS_THREADS, 3)
interseq_linear_fill_kernel(const InterseqConst C, const InterseqArgs A) {
    extern __shared__ uint32_t table[];
    const int lane = threadIdx.x & 31;
    {
        const int entries = C.n_rows * C.n_cols;
        for (int e = threadIdx.x >> 5; e < entries; e += IS_THREADS / 32) {
            const int a = e / C.n_cols, b = e % C.n_cols;
            table[(b * C.n_rows + a) * 32 + lane] = (uint32_t)A.matrix[e] & 0xffffu;
        }
    }
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * (IS_THREADS / 32) + (threadIdx.x >> 5);
    if (g >= A.n_groups) return;
    const WarpDesc d = A.desc[g];
    const int32_t p_lo = A.order[g * 64 + 2 * lane], p_hi = A.order[g * 64 + 2 * lane + 1];
    int n_lo = 0, n_hi = 0, m_lo = 0, m_hi = 0;
    if (p_lo >= 0) { n_lo = (int)A.map.len1(p_lo); m_lo = (int)A.map.len2(p_lo); }
    if (p_hi >= 0) { n_hi = (int)A.map.len1(p_hi); m_hi = (int)A.map.len2(p_hi); }

    const uint32_t gap2 = ((uint32_t)C.gap_open & 0xffffu) * 0x10001u;
    const uint32_t zero2 = (uint32_t)C.n_cols >> 31;  // packed zero in a register (see the affine kernel)
    const uint32_t col_stride = (uint32_t)C.n_rows * IS_ROWW;
    const char *tbase = (const char *)table + lane * 4;
    const uint16_t *rc = A.rowcodes + d.rowcode_off + lane;
    const uint16_t *cc = A.colcodes + d.colcode_off + lane;
    uint32_t *rb_s = A.rowbuf + d.rowbuf_off * 3 + lane;
    uint2 *dir = TRACED ? (uint2 *)(A.dirs + d.dir_off) + lane : nullptr;

    int32_t best_lo = 0, best_hi = 0, bidx_lo = 0, bidx_hi = 0, brow_lo = 0, brow_hi = 0;
    uint32_t best2 = 0;
    for (int s = 0; s < d.stripes; s++) {
        const int r0 = s * IS_R;
        const int j_lo = (n_lo - 1) / IS_R == s ? m_lo - 1 : -1;
        const int j_hi = (n_hi - 1) / IS_R == s ? m_hi - 1 : -1;
        const bool rows_valid = r0 + IS_R <= n_lo && r0 + IS_R <= n_hi;
        uint32_t rp_lo[IS_R], rp_hi[IS_R], Sl[IS_R];
#pragma unroll
        for (int r = 0; r < IS_R; r++) {
            const uint32_t v = rc[(int64_t)(r0 + r) * 32];
            rp_lo[r] = (v & 0xffu) * IS_ROWW;
            rp_hi[r] = (v >> 8) * IS_ROWW;
            // column 0 (pairwise.rs:434-450): S(i, 0) = i * g, local 0
            Sl[r] = MODE == IS_GLOBAL ? ((uint32_t)((r0 + r + 1) * C.gap_open) & 0xffffu) * 0x10001u : 0u;
        }
        uint32_t hu_prev = MODE == IS_GLOBAL ? ((uint32_t)(r0 * C.gap_open) & 0xffffu) * 0x10001u : 0u;  // S(r0, 0)
        uint32_t s_row0 = MODE == IS_GLOBAL ? gap2 : 0u;  // S(0, j) = j * g (pairwise.rs:452-467)
        // look-ahead registers for column code and row above, LPF columns ahead; the loop is
        // unrolled LPF times so that they need no rotation (reads past the last column land in
        // the slack behind the buffers); both streams are pulled into L2 further ahead
        constexpr int LPF = 2;
        uint32_t q_c[LPF], q_s[LPF];
#pragma unroll
        for (int k = 0; k < LPF; k++) { q_c[k] = cc[(int64_t)k * 32]; q_s[k] = s > 0 ? rb_s[(int64_t)k * 32] : 0u; }
        for (int j0 = 0; j0 < d.mmax; j0 += LPF) {
#pragma unroll
        for (int u = 0; u < LPF; u++) {
            const int j = j0 + u;
            if (u > 0 && j >= d.mmax) break;
            const uint32_t v = q_c[u];
            uint32_t up = q_s[u];
            q_c[u] = cc[(int64_t)(j + LPF) * 32];
            if (s > 0) q_s[u] = rb_s[(int64_t)(j + LPF) * 32];
            asm volatile("prefetch.global.L2 [%0];" ::"l"(cc + (int64_t)(j + LPF + IS_PREFETCH) * 32));
            if (s > 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(rb_s + (int64_t)(j + LPF + IS_PREFETCH) * 32));
            const char *t_lo = tbase + (v & 0xffu) * col_stride;
            const char *t_hi = tbase + (v >> 8) * col_stride;
            if (s == 0) {
                up = s_row0;
                if (MODE == IS_GLOBAL) s_row0 = __vadd2(s_row0, gap2);
            }
            uint32_t acc[2] = {0, 0};  // 2 bits per row: low pair, high pair
            uint32_t a[IS_R];
            // diagonal candidates first: they read the previous column's S before it is replaced
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                const uint32_t s_lo = *(const uint32_t *)(t_lo + rp_lo[r]);
                const uint32_t s_hi = *(const uint32_t *)(t_hi + rp_hi[r]);
                a[r] = __vadd2(r == 0 ? hu_prev : Sl[r - 1], s_hi * 65536u + s_lo);
            }
            hu_prev = up;
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                uint32_t t, sc;
                if (TRACED) {
                    BT_MAX_PRED(t, Sl[r], up, acc[0], acc[1], 2u << (2 * r));       // left >= up
                    const uint32_t tg = __vadd2(t, gap2);
                    BT_MAX_PRED(sc, a[r], tg, acc[0], acc[1], 1u << (2 * r));       // diag >= gap
                } else {
                    sc = __viaddmax_s16x2(__vmaxs2(Sl[r], up), gap2, a[r]);
                }
                if (MODE == IS_LOCAL) sc = __vmaxs2(sc, zero2);  // pairwise.rs:410-413
                Sl[r] = sc;
                up = sc;
            }
            if (s + 1 < d.stripes) rb_s[(int64_t)j * 32] = up;
            if (TRACED) dir[((int64_t)s * d.mmax + j) * 32] = make_uint2(acc[0], acc[1]);
            if (MODE == IS_GLOBAL) {
                if (j == j_lo || j == j_hi) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        if (j == j_lo && r0 + r + 1 == n_lo) best_lo = (int16_t)(Sl[r] & 0xffffu);
                        if (j == j_hi && r0 + r + 1 == n_hi) best_hi = (int16_t)(Sl[r] >> 16);
                    }
                }
            } else {
                bool slow = false;
                if (!TRACED) {
                    // score only: packed running maximum, padding rows / columns masked to 0
                    uint32_t cm = zero2;
                    if (rows_valid) {
#pragma unroll
                        for (int r = 0; r < IS_R; r += 2) cm = __vimax3_s16x2(cm, Sl[r], Sl[r + 1]);
                    } else {
#pragma unroll
                        for (int r = 0; r < IS_R; r++) {
                            const uint32_t rmask = (r0 + r < n_lo ? 0xffffu : 0u) | (r0 + r < n_hi ? 0xffff0000u : 0u);
                            cm = __vmaxs2(cm, Sl[r] & rmask);
                        }
                    }
                    const uint32_t cmask = (j < m_lo ? 0xffffu : 0u) | (j < m_hi ? 0xffff0000u : 0u);
                    best2 = __vmaxs2(best2, cm & cmask);
                } else {
                    uint32_t cm = Sl[0];
#pragma unroll
                    for (int r = 1; r < IS_R; r++) cm = __vmaxs2(cm, Sl[r]);
                    const bool inside = rows_valid && j < m_lo && j < m_hi;
                    slow = !inside;
                    if (inside) {
                        const int c_lo = (int16_t)(cm & 0xffffu), c_hi = (int16_t)(cm >> 16);
                        // a new optimum, or a tie that could precede the current first optimum in
                        // row-major order (only possible while that optimum lies in this stripe)
                        slow = c_lo > best_lo || c_hi > best_hi ||
                               (c_lo == best_lo && c_lo > 0 && brow_lo > r0 + 1) ||
                               (c_hi == best_hi && c_hi > 0 && brow_hi > r0 + 1);
                    }
                }
                if (slow) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        const int row = r0 + r + 1, col = j + 1;
                        const int v_lo = (int16_t)(Sl[r] & 0xffffu), v_hi = (int16_t)(Sl[r] >> 16);
                        if (row <= n_lo && col <= m_lo) {
                            const int idx = row * (m_lo + 1) + col;
                            if (v_lo > best_lo || (v_lo == best_lo && v_lo > 0 && idx < bidx_lo)) { best_lo = v_lo; bidx_lo = idx; brow_lo = row; }
                        }
                        if (row <= n_hi && col <= m_hi) {
                            const int idx = row * (m_hi + 1) + col;
                            if (v_hi > best_hi || (v_hi == best_hi && v_hi > 0 && idx < bidx_hi)) { best_hi = v_hi; bidx_hi = idx; brow_hi = row; }
                        }
                    }
                }
            }
        }
        }
    }
    if (MODE == IS_LOCAL && !TRACED) {
        best_lo = max(best_lo, (int)(int16_t)(best2 & 0xffffu));
        best_hi = max(best_hi, (int)(int16_t)(best2 >> 16));
    }
    if (p_lo >= 0) A.score[p_lo] = best_lo;
    if (p_hi >= 0) A.score[p_hi] = best_hi;
    if (MODE == IS_LOCAL && TRACED) {
        if (p_lo >= 0) { A.start[2 * p_lo] = bidx_lo / (m_lo + 1); A.start[2 * p_lo + 1] = bidx_lo % (m_lo + 1); }
        if (p_hi >= 0) { A.start[2 * p_hi] = bidx_hi / (m_hi + 1); A.start[2 * p_hi + 1] = bidx_hi % (m_hi + 1); }
    }
}

template <int MODE>
__global__ void interseq_linear_traceback_kernel(const InterseqConst C, const InterseqArgs A,
                                                 const uint8_t *__restrict__ codes1,
                                                 const uint8_t *__restrict__ codes2,
                                                 int64_t *__restrict__ trace_len, int32_t *__restrict__ first,
                                                 uint8_t *__restrict__ ops, int64_t n_slots) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int32_t p = A.order[slot];
    if (p < 0) return;
    const int64_t g = slot >> 6;
    const int lane = (int)(slot & 63) >> 1, half = (int)(slot & 1);
    const WarpDesc d = A.desc[g];
    const int64_t o1 = A.map.beg1(p), o2 = A.map.beg2(p);
    const int n = (int)A.map.len1(p), m = (int)A.map.len2(p);
    const uint32_t *base = A.dirs + d.dir_off + lane * 2 + half;
    uint8_t *out = C.stats ? nullptr : ops + A.map.ops(p);
    GapStats gs(C.terminal != 0);
    // step bytes are written four at a time once the output position is word-aligned (one
    // store transaction instead of four; the unaligned head and the tail go out as bytes)
    uint32_t ew = 0;
    int enb = 0;
    const int ealign = (int)((uintptr_t)out & 3u);
    auto emit = [&](int op, int &len) {
        if (C.stats) gs.push(op);
        else if (enb == 0 && ((ealign + len) & 3) != 0) out[len] = (uint8_t)op;
        else {
            ew |= (uint32_t)op << (8 * enb);
            if (++enb == 4) { *(uint32_t *)(out + len - 3) = ew; ew = 0; enb = 0; }
        }
        len++;
    };
    auto emit_flush = [&](int len) {
        for (int k = 0; k < enb; k++) out[len - enb + k] = (uint8_t)(ew >> (8 * k));
    };
    int i = n, j = m, len = 0, fi = 0, fj = 0;
    int32_t value = 0;
    if (MODE == IS_LOCAL) { i = A.start[2 * p]; j = A.start[2 * p + 1]; value = A.score[p]; }
    while (i > 0 || j > 0) {
        if (MODE == IS_LOCAL && (i == 0 || j == 0 || value <= 0)) break;  // score <= 0: direction cleared
        fi = i; fj = j;
        if (i == 0) { emit(OP_TO1, len); j--; continue; }
        if (j == 0) { emit(OP_TO2, len); i--; continue; }
        const int rg = i - 1, s = rg / IS_R, r = rg % IS_R;
        const uint32_t w = base[((int64_t)s * d.mmax + (j - 1)) * 64];
        const uint32_t bits = (w >> (2 * r)) & 3u;
        if (bits & 1u) {
            if (MODE == IS_LOCAL) value -= A.matrix[min((int)codes1[o1 + i - 1], C.n_rows - 1) * C.n_cols + min((int)codes2[o2 + j - 1], C.n_cols - 1)];
            emit(OP_DIAG, len); i--; j--;
        } else {
            if (MODE == IS_LOCAL) value -= C.gap_open;
            if (bits & 2u) { emit(OP_TO1, len); j--; }
            else { emit(OP_TO2, len); i--; }
        }
    }
    if (!C.stats && out) emit_flush(len);
    trace_len[p] = len;
    if (C.stats) { gs.finish(); fi = gs.opens; fj = gs.exts; }
    first[2 * p] = fi;
    first[2 * p + 1] = fj;
}

}  // namespace bt
#include "interseq_tag_kernel.cuh"
namespace bt {

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

// Groups per wave of the affine fill kernel (one warp per block, IS_WARPS_PER_SM blocks per SM):
// windows are whole numbers of waves, so that they lose little to partial waves.
inline int64_t interseq_wave_groups(int warps_per_sm = IS_WARPS_PER_SM) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;   // B200; only reached without a device (bt_debug_plan in the CPU tests)
    }
    return (int64_t)sms * warps_per_sm;
}

// Pairs per window for a batch of n_pairs when a window should span `waves` full waves.
inline int64_t interseq_pairs_per_window(int64_t n_pairs, int waves, int warps_per_sm = IS_WARPS_PER_SM) {
    int64_t win_pairs = interseq_wave_groups(warps_per_sm) * waves * 64;
    if (const char *g = getenv("B200ALIGN_WIN_GROUPS")) win_pairs = std::max<int64_t>(64, atoll(g) * 64);
    const int64_t n_windows = std::max<int64_t>(1, (n_pairs + win_pairs - 1) / win_pairs);
    return ((n_pairs + n_windows - 1) / n_windows + 63) / 64 * 64;  // balanced windows
}

struct WindowPlan {
    Window win{};
    std::vector<int32_t> order;
    std::vector<WarpDesc> desc;
    int64_t cells = 0;
};

// Runs f(t, begin, end) on n_threads chunks of [0, cnt).
template <typename F>
inline void interseq_parallel_chunks(int64_t cnt, int n_threads, F f) {
    std::vector<std::thread> threads;
    for (int t = 1; t < n_threads; t++) threads.emplace_back(f, t, cnt * t / n_threads, cnt * (t + 1) / n_threads);
    f(0, (int64_t)0, cnt / n_threads);
    for (std::thread &t : threads) t.join();
}

// One stable counting pass: out = in (or the identity id0 + q) ordered by key(id) ascending.
template <typename K>
inline void interseq_counting_pass(const int32_t *in, int64_t id0, int64_t cnt, K key, int32_t *out, int n_threads) {
    const size_t buckets = IS_MAX_LEN + 2;
    std::vector<int64_t> hist((size_t)n_threads * buckets, 0);
    auto id = [&](int64_t q) -> int32_t { return in ? in[q] : (int32_t)(id0 + q); };
    interseq_parallel_chunks(cnt, n_threads, [&](int t, int64_t b, int64_t e) {
        int64_t *h = hist.data() + (size_t)t * buckets;
        for (int64_t q = b; q < e; q++) h[(size_t)key(id(q))]++;
    });
    int64_t run = 0;   // bucket-major, thread-minor prefix: keeps the pass stable
    for (size_t c = 0; c < buckets; c++)
        for (int t = 0; t < n_threads; t++) { int64_t &h = hist[(size_t)t * buckets + c]; const int64_t v = h; h = run; run += v; }
    interseq_parallel_chunks(cnt, n_threads, [&](int t, int64_t b, int64_t e) {
        int64_t *h = hist.data() + (size_t)t * buckets;
        for (int64_t q = b; q < e; q++) { const int32_t k = id(q); out[h[(size_t)key(k)]++] = k; }
    });
}

// Pair ids `ids[0..cnt)` sorted by (len1, len2) ascending: two stable counting passes
// (threaded for large batches).
inline void interseq_sort_pairs(const PairMap &hm, const int32_t *ids, int64_t id0, int64_t cnt,
                                std::vector<int32_t> &sorted) {
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), cnt / 262144));
    std::vector<int32_t> tmp((size_t)cnt);
    sorted.resize((size_t)cnt);
    interseq_counting_pass(ids, id0, cnt, [&](int32_t k) { return hm.len2(k); }, tmp.data(), n_threads);
    interseq_counting_pass(tmp.data(), 0, cnt, [&](int32_t k) { return hm.len1(k); }, sorted.data(), n_threads);
}

// Cost (row steps of the fill kernel) of the group of 64 slots starting at `slots`.
inline int64_t interseq_group_cost(const PairMap &hm, const int32_t *slots) {
    int64_t nmax = 1, mmax = 1;
    for (int q = 0; q < 64; q++) {
        if (slots[q] < 0) continue;
        nmax = std::max<int64_t>(nmax, hm.len1(slots[q])); mmax = std::max<int64_t>(mmax, hm.len2(slots[q]));
    }
    return (nmax + IS_R - 1) / IS_R * mmax;
}

// Reorder whole groups of `order` (64 slots each) by descending cost (stable).
inline void interseq_order_groups_by_cost(const PairMap &hm, std::vector<int32_t> &order) {
    const int64_t groups = (int64_t)order.size() / 64;
    std::vector<int64_t> cost((size_t)groups);
    std::vector<int32_t> perm((size_t)groups);
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), groups / 8192));
    interseq_parallel_chunks(groups, n_threads, [&](int, int64_t b, int64_t e) {
        for (int64_t g = b; g < e; g++) { cost[(size_t)g] = interseq_group_cost(hm, order.data() + g * 64); perm[(size_t)g] = (int32_t)g; }
    });
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    std::vector<int32_t> out(order.size());
    interseq_parallel_chunks(groups, n_threads, [&](int, int64_t b, int64_t e) {
        for (int64_t g = b; g < e; g++) memcpy(out.data() + g * 64, order.data() + (int64_t)perm[(size_t)g] * 64, 64 * 4);
    });
    order.swap(out);
}

// Host array without value initialisation (std::vector would memset hundreds of MB on one thread
// before the planner's threads overwrite every element).
template <typename T>
struct RawBuf {
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void alloc(size_t count) { p.reset(new T[count]); n = count; }
    T *data() { return p.get(); }
    const T *data() const { return p.get(); }
    bool empty() const { return n == 0; }
    size_t size() const { return n; }
    T &operator[](size_t i) { return p[i]; }
    const T &operator[](size_t i) const { return p[i]; }
};

// Batch-wide order of a ragged resident batch: the result of interseq_sort_pairs (ascending by
// (len1, len2), stable), reversed into groups of 64 slots and ordered by descending group cost
// (interseq_order_groups_by_cost) - but the lengths travel with the pair ids as 64-bit records
// (len1 << 48 | len2 << 32 | id), so the counting passes and everything after them stream
// through memory instead of looking the lengths of every pair up again (10^8 pairs: the
// look-ups were most of the planning time).  Also returns the lengths per slot
// (len1 << 16 | len2, 0 for padding) and the longest sequences per group.
inline void interseq_order_batch(const PairMap &hm, int64_t n_pairs, RawBuf<int32_t> &order,
                                 RawBuf<uint32_t> &order_nm, std::vector<int32_t> &group_n,
                                 std::vector<int32_t> &group_m, double *t_sorted_ms) {
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), n_pairs / 262144));
    const size_t buckets = IS_MAX_LEN + 2;
    // uninitialised on purpose: the threads below touch (and so place) the pages themselves
    std::unique_ptr<uint64_t[]> rec_a(new uint64_t[(size_t)n_pairs]), rec_b(new uint64_t[(size_t)n_pairs]);
    interseq_parallel_chunks(n_pairs, n_threads, [&](int, int64_t b, int64_t e) {
        for (int64_t q = b; q < e; q++)
            rec_a[(size_t)q] = (uint64_t)hm.len1(q) << 48 | (uint64_t)hm.len2(q) << 32 | (uint64_t)(uint32_t)q;
    });
    // one stable counting pass over the records by the 16-bit field at `shift`
    auto pass = [&](const uint64_t *in, uint64_t *out, int shift) {
        std::vector<int64_t> hist((size_t)n_threads * buckets, 0);
        interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t b, int64_t e) {
            int64_t *h = hist.data() + (size_t)t * buckets;
            for (int64_t q = b; q < e; q++) h[(size_t)(in[q] >> shift & 0xffffu)]++;
        });
        int64_t run = 0;   // bucket-major, thread-minor prefix: keeps the pass stable
        for (size_t c = 0; c < buckets; c++)
            for (int t = 0; t < n_threads; t++) { int64_t &h = hist[(size_t)t * buckets + c]; const int64_t v = h; h = run; run += v; }
        interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t b, int64_t e) {
            int64_t *h = hist.data() + (size_t)t * buckets;
            for (int64_t q = b; q < e; q++) out[h[(size_t)(in[q] >> shift & 0xffffu)]++] = in[q];
        });
    };
    pass(rec_a.get(), rec_b.get(), 32);   // by len2
    pass(rec_b.get(), rec_a.get(), 48);   // then by len1
    rec_b.reset();
    if (t_sorted_ms) *t_sorted_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();

    // groups of 64 slots over the descending order; cost and longest sequences per group
    const int64_t groups = (n_pairs + 63) / 64;
    std::vector<int64_t> cost((size_t)groups);
    std::vector<int32_t> perm((size_t)groups), gn((size_t)groups), gm((size_t)groups);
    const int g_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), groups / 8192));
    interseq_parallel_chunks(groups, g_threads, [&](int, int64_t b, int64_t e) {
        for (int64_t g = b; g < e; g++) {
            int32_t nmax = 1, mmax = 1;
            const int64_t s1 = std::min<int64_t>(n_pairs, g * 64 + 64);
            for (int64_t slot = g * 64; slot < s1; slot++) {
                const uint64_t r = rec_a[(size_t)(n_pairs - 1 - slot)];
                nmax = std::max<int32_t>(nmax, (int32_t)(r >> 48)); mmax = std::max<int32_t>(mmax, (int32_t)(r >> 32 & 0xffffu));
            }
            gn[(size_t)g] = nmax; gm[(size_t)g] = mmax;
            cost[(size_t)g] = (int64_t)(nmax + IS_R - 1) / IS_R * mmax;
            perm[(size_t)g] = (int32_t)g;
        }
    });
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    order.alloc((size_t)groups * 64);
    order_nm.alloc((size_t)groups * 64);
    group_n.resize((size_t)groups); group_m.resize((size_t)groups);
    interseq_parallel_chunks(groups, g_threads, [&](int, int64_t b, int64_t e) {
        for (int64_t g = b; g < e; g++) {
            const int64_t src = perm[(size_t)g];
            group_n[(size_t)g] = gn[(size_t)src]; group_m[(size_t)g] = gm[(size_t)src];
            const int64_t s1 = std::min<int64_t>(n_pairs, src * 64 + 64);
            int64_t q = g * 64;
            for (int64_t slot = src * 64; slot < s1; slot++, q++) {
                const uint64_t r = rec_a[(size_t)(n_pairs - 1 - slot)];
                order[(size_t)q] = (int32_t)(uint32_t)r;
                order_nm[(size_t)q] = (uint32_t)(r >> 48) << 16 | (uint32_t)(r >> 32 & 0xffffu);
            }
            for (; q < g * 64 + 64; q++) { order[(size_t)q] = -1; order_nm[(size_t)q] = 0u; }   // padding slots
        }
    });
}

inline void interseq_plan_window(WindowPlan &wp, const PairMap &hm, int64_t w0, int64_t w1, int score_only,
                                 int dir_words_per_lane, const int32_t *slots = nullptr, int64_t n_slot_ids = 0,
                                 const uint32_t *slot_nm = nullptr) {
    Window &win = wp.win;
    win.pair_begin = w0; win.pair_end = w1;
    int64_t groups;
    if (slots) {
        // resident batch: the groups were formed and ordered globally (interseq_plan_host), which
        // also copies them into the plan; wp.order stays empty
        groups = n_slot_ids / 64;
    } else {
        const int64_t cnt = w1 - w0;
        std::vector<int32_t> sorted;
        interseq_sort_pairs(hm, nullptr, w0, cnt, sorted);
        groups = (cnt + 63) / 64;
        wp.order.assign((size_t)groups * 64, -1);
        std::copy(sorted.rbegin(), sorted.rend(), wp.order.begin());
        interseq_order_groups_by_cost(hm, wp.order);
    }
    wp.desc.resize((size_t)groups);
    const int32_t *ord = slots ? slots : wp.order.data();
    for (int64_t g = 0; g < groups; g++) {
        int nmax = 1, mmax = 1;
        for (int q = 0; q < 64; q++) {
            const int32_t k = ord[g * 64 + q];
            if (k < 0) continue;
            // lengths per slot when the caller carries them (interseq_order_batch)
            const int64_t n = slot_nm ? (int64_t)(slot_nm[g * 64 + q] >> 16) : hm.len1(k);
            const int64_t m = slot_nm ? (int64_t)(slot_nm[g * 64 + q] & 0xffffu) : hm.len2(k);
            nmax = std::max<int>(nmax, (int)n); mmax = std::max<int>(mmax, (int)m);
            wp.cells += n * m;
        }
        WarpDesc &d = wp.desc[(size_t)g];
        d.nmax = nmax; d.mmax = mmax; d.stripes = (nmax + IS_R - 1) / IS_R; d.pad = 0;
        d.dir_off = win.dir_words; d.rowbuf_off = win.rowbuf_entries;
        d.rowcode_off = win.rowcodes; d.colcode_off = win.colcodes;
        // affine: uint4 (4 bits x 16 rows x 2 pairs), linear: uint2, per lane, column and stripe
        if (!score_only) win.dir_words += (int64_t)d.stripes * mmax * 32 * dir_words_per_lane;
        win.rowbuf_entries += (int64_t)mmax * 32;
        win.rowcodes += (int64_t)d.stripes * IS_R * 32;
        win.colcodes += (int64_t)mmax * 32;
    }
}

// Plan of one window of consecutive pairs [w0, w1) for the streaming path: same result as
// interseq_plan_window (sort by (len1, len2), reversed into groups of 64, groups by descending
// cost), but the lengths travel with the pair ids as 64-bit records (len1 << 48 | len2 << 32 |
// id) through two counting passes, so that nothing looks a length up twice: the per-pair
// look-ups of the offsets were most of the 25 - 65 ns per pair the first version took, and with
// one process per GPU the host has four cores per rank for it.
inline void interseq_plan_window_records(WindowPlan &wp, const PairMap &hm, int64_t w0, int64_t w1, int score_only,
                                         int dir_words_per_lane) {
    Window &win = wp.win;
    win.pair_begin = w0; win.pair_end = w1;
    const int64_t cnt = w1 - w0;
    const size_t buckets = IS_MAX_LEN + 2;
    std::unique_ptr<uint64_t[]> rec_a(new uint64_t[(size_t)cnt]), rec_b(new uint64_t[(size_t)cnt]);
    std::vector<int32_t> hist1(buckets, 0), hist2(buckets, 0);
    for (int64_t q = 0; q < cnt; q++) {
        const int64_t n = hm.len1(w0 + q), m = hm.len2(w0 + q);
        rec_a[(size_t)q] = (uint64_t)n << 48 | (uint64_t)m << 32 | (uint64_t)(uint32_t)(w0 + q);
        hist1[(size_t)n]++; hist2[(size_t)m]++;
    }
    auto exclusive = [&](std::vector<int32_t> &h) { int32_t run = 0; for (int32_t &v : h) { const int32_t c = v; v = run; run += c; } };
    exclusive(hist1); exclusive(hist2);
    for (int64_t q = 0; q < cnt; q++) rec_b[(size_t)hist2[(size_t)(rec_a[(size_t)q] >> 32 & 0xffffu)]++] = rec_a[(size_t)q];   // by len2
    for (int64_t q = 0; q < cnt; q++) rec_a[(size_t)hist1[(size_t)(rec_b[(size_t)q] >> 48)]++] = rec_b[(size_t)q];             // then by len1 (stable)
    rec_b.reset();
    // groups of 64 over the descending order, their longest sequences and cost
    const int64_t groups = (cnt + 63) / 64;
    std::vector<int64_t> cost((size_t)groups);
    std::vector<int32_t> perm((size_t)groups), gn((size_t)groups), gm((size_t)groups);
    for (int64_t g = 0; g < groups; g++) {
        int32_t nmax = 1, mmax = 1;
        const int64_t s1 = std::min<int64_t>(cnt, g * 64 + 64);
        for (int64_t slot = g * 64; slot < s1; slot++) {
            const uint64_t r = rec_a[(size_t)(cnt - 1 - slot)];
            nmax = std::max<int32_t>(nmax, (int32_t)(r >> 48)); mmax = std::max<int32_t>(mmax, (int32_t)(r >> 32 & 0xffffu));
        }
        gn[(size_t)g] = nmax; gm[(size_t)g] = mmax;
        cost[(size_t)g] = (int64_t)(nmax + IS_R - 1) / IS_R * mmax;
        perm[(size_t)g] = (int32_t)g;
    }
    std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    wp.order.assign((size_t)groups * 64, -1);
    wp.desc.resize((size_t)groups);
    for (int64_t g = 0; g < groups; g++) {
        const int64_t src = perm[(size_t)g];
        const int64_t s1 = std::min<int64_t>(cnt, src * 64 + 64);
        int64_t q = g * 64;
        for (int64_t slot = src * 64; slot < s1; slot++, q++) {
            const uint64_t r = rec_a[(size_t)(cnt - 1 - slot)];
            wp.order[(size_t)q] = (int32_t)(uint32_t)r;
            wp.cells += (int64_t)(r >> 48) * (int64_t)(r >> 32 & 0xffffu);
        }
        const int nmax = gn[(size_t)src], mmax = gm[(size_t)src];
        WarpDesc &d = wp.desc[(size_t)g];
        d.nmax = nmax; d.mmax = mmax; d.stripes = (nmax + IS_R - 1) / IS_R; d.pad = 0;
        d.dir_off = win.dir_words; d.rowbuf_off = win.rowbuf_entries;
        d.rowcode_off = win.rowcodes; d.colcode_off = win.colcodes;
        if (!score_only) win.dir_words += (int64_t)d.stripes * mmax * 32 * dir_words_per_lane;
        win.rowbuf_entries += (int64_t)mmax * 32;
        win.rowcodes += (int64_t)d.stripes * IS_R * 32;
        win.colcodes += (int64_t)mmax * 32;
    }
}

// Kernel constants of a plan; returns the "no direction nibbles" flag of the trace mode
// (0 = traces, 1 = score only, 2 = statistics).
inline int interseq_setup(InterseqPlan &plan, const BtParams &P, int64_t n_pairs, int trace_mode, int32_t min_score) {
    plan.n_pairs = n_pairs;
    plan.C.gap_open = P.gap_open; plan.C.gap_ext = P.affine ? P.gap_ext : P.gap_open;
    plan.C.affine = P.affine;
    plan.C.neg = -32768 - plan.C.gap_open - plan.C.gap_ext - std::min(min_score, 0);
    plan.C.n_rows = P.n_rows; plan.C.n_cols = P.n_cols;
    plan.C.stats = trace_mode == 2;
    plan.C.terminal = P.terminal_penalty;
    plan.C.score_only = trace_mode == 1;
    plan.C.mode = P.local ? IS_LOCAL : (P.terminal_penalty ? IS_GLOBAL : IS_SEMI);
    plan.C.dir_layout = 0;
    // windows are whole waves of the kernel that will run: size them for profiles of the 20
    // standard letters (windows with rarer letters run the 9-warp variant on slightly odd waves)
    if (P.affine) plan.warps_per_sm = interseq_profile_warps(std::min<int>(P.n_cols, IS_SMALL_LETTERS));
    return plan.C.score_only;
}

// Traced affine batches whose scores provably fit 13 bits run the tag kernel: scores scaled by 4,
// so the 16-bit "minus infinity" moves to -8192 - open - ext - min(0, min sim).
inline void interseq_enable_tags(InterseqPlan &plan, const BtParams &P, const PairScan &scan, int32_t min_score,
                                 int32_t max_score) {
    if (plan.C.score_only || !interseq_tag_eligible(P, scan, min_score, max_score)) return;
    plan.C.dir_layout = 1;
    // persistent tag kernel: windows without the rare letters (B, Z, X, * of the protein alphabet)
    // are the rule, size the waves for them
    plan.warps_per_sm = interseq_tag_warps(std::min<int>(P.n_cols, TG_SMALL_LETTERS));
    plan.C.neg = -8192 - plan.C.gap_open - plan.C.gap_ext - std::min(min_score, 0);
}

// Host part of the plan of a resident batch (no device work: also reachable through
// bt_debug_plan for the CPU tests): group order, descriptors and windows.
inline void interseq_plan_host(InterseqPlan &plan, const BtParams &P, const PairMap &hm, int64_t n_pairs,
                               int score_only, int32_t min_score, int32_t max_score, int64_t *cells_out,
                               int waves_per_window) {
    score_only = interseq_setup(plan, P, n_pairs, score_only, min_score);
    const PairScan batch_scan = interseq_scan(hm, n_pairs);
    interseq_enable_tags(plan, P, batch_scan, min_score, max_score);

    const int64_t per_window = interseq_pairs_per_window(n_pairs, waves_per_window, plan.warps_per_sm);
    const int64_t groups_per_window = per_window / 64;
    const int dwpl = P.affine ? 4 : 2;
    // The whole batch is resident: sort it globally by (len1, len2), form the groups of 64, order
    // the groups by descending cost and cut the windows from that list, so that the groups of a
    // window cost about the same and no window waits for a straggler.  A window also ends when
    // its direction nibbles would exceed IS_DIR_BUDGET (long pairs: 5000 x 5000 residues need
    // 0.8 GB per group).
    // a batch is "uniform" when no pair costs more than a few times the average one: then
    // stragglers cannot dominate a window and the batch-wide sort buys nothing
    bool uniform = false;
    {
        const PairScan &scan = batch_scan;
        uniform = scan.max_pair * (double)n_pairs <= 4.0 * scan.total;
        if (const char *f = getenv("B200ALIGN_SORT")) uniform = strcmp(f, "window") == 0;
    }
    const bool trace_plan = getenv("B200ALIGN_TRACE") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tp0 = now_ms();
    double tp1 = tp0, tp2 = tp0, tp3 = tp0;
    RawBuf<int32_t> global;
    RawBuf<uint32_t> global_nm;                // lengths per slot of `global` (batch-wide sort only)
    std::vector<int32_t> group_n, group_m;     // longest sequences per group (batch-wide sort only)
    std::vector<int64_t> wstart{0};            // first group of every window (+ end)
    const int64_t n_groups_total = (n_pairs + 63) / 64;
    {
        if (!uniform) {
            // batch-wide sort, groups ordered by cost (also inside a single window: longest first)
            interseq_order_batch(hm, n_pairs, global, global_nm, group_n, group_m, &tp1);
        } else {
            global.alloc((size_t)(n_groups_total * 64));
            std::fill(global.data(), global.data() + global.size(), -1);
            // pairs of about the same size: windows of consecutive pairs, sorted one by one -
            // their results stay contiguous, which the traceback's step-byte stores like
            for (int64_t w0 = 0; w0 < n_pairs; w0 += per_window) {
                const int64_t cnt = std::min(per_window, n_pairs - w0);
                std::vector<int32_t> sorted, seg((size_t)((cnt + 63) / 64 * 64), -1);
                interseq_sort_pairs(hm, nullptr, w0, cnt, sorted);
                std::copy(sorted.rbegin(), sorted.rend(), seg.begin());
                interseq_order_groups_by_cost(hm, seg);
                // per_window is a multiple of 64, so only the last segment carries padding
                std::copy(seg.begin(), seg.end(), global.data() + w0);
            }
            tp1 = now_ms();
        }
        tp2 = now_ms();
        int64_t in_window = 0, dir_bytes = 0;
        for (int64_t g = 0; g < n_groups_total; g++) {
            int64_t bytes = 0;
            if (!score_only && !group_n.empty()) {
                bytes = (int64_t)(group_n[(size_t)g] + IS_R - 1) / IS_R * group_m[(size_t)g] * 32 * dwpl * 4;
            } else if (!score_only) {
                int64_t nmax = 1, mmax = 1;
                for (int q = 0; q < 64; q++) {
                    const int32_t k = global[(size_t)(g * 64 + q)];
                    if (k < 0) continue;
                    nmax = std::max(nmax, hm.len1(k)); mmax = std::max(mmax, hm.len2(k));
                }
                bytes = (nmax + IS_R - 1) / IS_R * mmax * 32 * dwpl * 4;
            }
            if (in_window > 0 && (in_window >= groups_per_window || dir_bytes + bytes > IS_DIR_BUDGET)) {
                wstart.push_back(g);
                in_window = 0; dir_bytes = 0;
            }
            in_window++; dir_bytes += bytes;
        }
        wstart.push_back(n_groups_total);
    }
    const int64_t n_windows = (int64_t)wstart.size() - 1;
    std::vector<WindowPlan> wps((size_t)n_windows);
    {
        // windows are independent: lay them out on a few host threads
        const unsigned hw = interseq_host_threads();
        const int n_threads = (int)std::min<int64_t>(std::min<unsigned>(hw, 8u), n_windows);
        std::atomic<int64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const int64_t w = next.fetch_add(1);
                if (w >= n_windows) break;
                const int64_t g0 = wstart[(size_t)w] * 64, g1 = wstart[(size_t)w + 1] * 64;
                interseq_plan_window(wps[(size_t)w], hm, 0, 0, score_only, dwpl, global.data() + g0, g1 - g0,
                                     global_nm.empty() ? nullptr : global_nm.data() + g0);
            }
        };
        std::vector<std::thread> threads;
        for (int t = 1; t < n_threads; t++) threads.emplace_back(work);
        work();
        for (std::thread &t : threads) t.join();
    }
    tp3 = now_ms();
    plan.windows.clear();
    plan.order.clear();
    plan.desc.clear();
    int64_t cells = 0;
    {
        // concatenate the windows' group orders / descriptors (copied on a few threads)
        int64_t g = 0;
        for (WindowPlan &wp : wps) {
            wp.win.group_begin = g;
            g += (int64_t)wp.desc.size();
            wp.win.group_end = g;
        }
        plan.order.resize((size_t)g * 64);
        plan.desc.resize((size_t)g);
        const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 8u), n_windows));
        std::atomic<int64_t> next{0};
        auto copy = [&]() {
            for (;;) {
                const int64_t w = next.fetch_add(1);
                if (w >= n_windows) break;
                WindowPlan &wp = wps[(size_t)w];
                // window w holds groups wstart[w] .. wstart[w + 1] of the batch-wide list
                memcpy(plan.order.data() + wp.win.group_begin * 64, global.data() + wstart[(size_t)w] * 64,
                       (size_t)(wstart[(size_t)w + 1] - wstart[(size_t)w]) * 64 * 4);
                if (!wp.desc.empty()) memcpy(plan.desc.data() + wp.win.group_begin, wp.desc.data(), wp.desc.size() * sizeof(WarpDesc));
            }
        };
        std::vector<std::thread> threads;
        for (int t = 1; t < n_threads; t++) threads.emplace_back(copy);
        copy();
        for (std::thread &t : threads) t.join();
    }
    for (WindowPlan &wp : wps) {
        plan.max_dir_words = std::max(plan.max_dir_words, wp.win.dir_words);
        plan.max_rowbuf = std::max(plan.max_rowbuf, wp.win.rowbuf_entries);
        plan.max_rowcodes = std::max(plan.max_rowcodes, wp.win.rowcodes);
        plan.max_colcodes = std::max(plan.max_colcodes, wp.win.colcodes);
        plan.windows.push_back(wp.win);
        cells += wp.cells;
    }
    plan.n_groups = (int64_t)plan.desc.size();
    if (cells_out) *cells_out = cells;
    if (trace_plan)
        fprintf(stderr, "[b200align] plan: sort %.1f ms, group order %.1f ms, window cut + layout %.1f ms, merge %.1f ms\n",
                tp1 - tp0, tp2 - tp1, tp3 - tp2, now_ms() - tp3);
}

inline int interseq_prepare(InterseqPlan &plan, const BtParams &P, const PairMap &hm, int64_t n_pairs, int score_only,
                            int32_t min_score, int32_t max_score, int64_t *cells_out, int n_slots,
                            int waves_per_window, cudaStream_t stream) {
    interseq_plan_host(plan, P, hm, n_pairs, score_only, min_score, max_score, cells_out, waves_per_window);
    cudaError_t e;
    if ((e = plan.d_order.alloc(plan.order.size() * 4)) != cudaSuccess) goto fail;
    if ((e = plan.d_desc.alloc(plan.desc.size() * sizeof(WarpDesc))) != cudaSuccess) goto fail;
    for (int sl = 0; sl < n_slots; sl++) {
        SlotBuffers &b = plan.slots[sl];
        if ((e = b.rowcodes.alloc((size_t)plan.max_rowcodes * 2)) != cudaSuccess) goto fail;
        if ((e = b.colcodes.alloc((size_t)(plan.max_colcodes + IS_SLACK * 32) * 2)) != cudaSuccess) goto fail;
        if ((e = b.rowbuf.alloc((size_t)(plan.max_rowbuf + IS_SLACK * 32) * 12)) != cudaSuccess) goto fail;
        if ((e = b.dirs.alloc((size_t)plan.max_dir_words * 4)) != cudaSuccess) goto fail;
        if ((e = b.ctl.alloc(16)) != cudaSuccess) goto fail;
        if (plan.C.mode == IS_SEMI) {
            if ((e = b.lastrow.alloc((size_t)plan.max_rowbuf * 8)) != cudaSuccess) goto fail;
            if ((e = b.lastcol.alloc((size_t)plan.max_rowcodes * 8)) != cudaSuccess) goto fail;
        }
    }
    if ((e = cudaMemcpyAsync(plan.d_order.ptr, plan.order.data(), plan.order.size() * 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto fail;
    if ((e = cudaMemcpyAsync(plan.d_desc.ptr, plan.desc.data(), plan.desc.size() * sizeof(WarpDesc), cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto fail;
    return BT_OK;
fail:
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BT_ERR_OOM : BT_ERR_CUDA;
}

inline InterseqArgs interseq_args(InterseqPlan &plan, const Window &win, int slot, const InterseqIO &io) {
    SlotBuffers &b = plan.slots[slot];
    InterseqArgs A{};
    A.desc = plan.d_desc.as<WarpDesc>() + win.group_begin;
    A.order = plan.d_order.as<int32_t>() + win.group_begin * 64;
    A.map = io.map; A.matrix = io.matrix;
    A.rowcodes = b.rowcodes.as<uint16_t>(); A.colcodes = b.colcodes.as<uint16_t>();
    A.rowbuf = b.rowbuf.as<uint32_t>(); A.lastrow = b.lastrow.as<uint32_t>(); A.lastcol = b.lastcol.as<uint32_t>();
    A.dirs = b.dirs.as<uint32_t>();
    A.score = io.score; A.start = io.start;
    A.n_groups = win.group_end - win.group_begin;
    return A;
}

template <int MODE, bool TRACED>
inline cudaError_t interseq_launch_linear_fill(const InterseqConst &C, const InterseqArgs &A, unsigned grid,
                                               size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(interseq_linear_fill_kernel<MODE, TRACED>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    interseq_linear_fill_kernel<MODE, TRACED><<<grid, IS_THREADS, smem, stream>>>(C, A);
    return cudaGetLastError();
}

// register prefetch distance (columns) of the row above the stripe: the score-only kernels spend
// less time per column and have registers to spare, so they look further ahead
template <bool TRACED> constexpr int interseq_pf() { return TRACED ? BT_PF_TRACED : BT_PF_SCORE; }

template <int MODE, bool TRACED>
inline cudaError_t interseq_launch_fill_variant(const InterseqConst &C, const InterseqArgs &A, int n_eff,
                                                int small_letters, int *ctl, int counter_slot, cudaStream_t stream) {
    // profile: 2 halves x n_eff letters x 32 lanes x 16 rows, per warp
    const int warps = interseq_profile_warps(n_eff);
    const size_t smem = (size_t)warps * 2 * n_eff * 512;
    const int64_t sms = interseq_wave_groups() / IS_WARPS_PER_SM;
    const unsigned grid = (unsigned)std::min<int64_t>(sms, (A.n_groups + warps - 1) / warps);
    cudaError_t e;
    if (warps > 11) {   // small alphabets: up to 16 warps (128 registers each)
        auto kernel = interseq_fill_kernel<MODE, TRACED, interseq_pf<TRACED>(), 16>;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kernel<<<grid, 32 * warps, smem, stream>>>(C, A, n_eff, small_letters, ctl, counter_slot);
    } else {
        auto kernel = interseq_fill_kernel<MODE, TRACED, interseq_pf<TRACED>(), 11>;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kernel<<<grid, 32 * warps, smem, stream>>>(C, A, n_eff, small_letters, ctl, counter_slot);
    }
    return cudaGetLastError();
}

// Two launches when the matrix has more letters than the 20 standard amino acids (B, Z, X, * of the
// protein alphabet): the variant whose profiles hold 20 letters (more warps per SM) runs for windows
// that use only those, else it returns at once and the variant for all letters does the work.
template <int MODE, bool TRACED>
inline cudaError_t interseq_launch_fill(const InterseqConst &C, const InterseqArgs &A, int *ctl,
                                        cudaStream_t stream, int64_t *launches) {
    int small_letters = 0;
    cudaError_t e = cudaSuccess;
    if (C.n_cols > IS_SMALL_LETTERS && interseq_profile_warps(IS_SMALL_LETTERS) > interseq_profile_warps(C.n_cols)) {
        small_letters = IS_SMALL_LETTERS;
        e = interseq_launch_fill_variant<MODE, TRACED>(C, A, small_letters, small_letters, ctl, 1, stream);
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    e = interseq_launch_fill_variant<MODE, TRACED>(C, A, C.n_cols, small_letters, ctl, 2, stream);
    if (e == cudaSuccess) (*launches)++;
    return e;
}

inline cudaError_t interseq_fill_window(InterseqPlan &plan, const Window &win, int slot, const InterseqIO &io,
                                        cudaStream_t stream, int64_t *launches) {
    const InterseqArgs A = interseq_args(plan, win, slot, io);
    const int64_t groups = A.n_groups;
    int *const ctl = plan.slots[slot].ctl.as<int>();
    cudaError_t e = cudaMemsetAsync(ctl, 0, 16, stream);
    if (e != cudaSuccess) return e;
    interseq_pack_kernel<<<(unsigned)groups, 256, 0, stream>>>(
        (const uint8_t *)io.codes1, (const uint8_t *)io.codes2, io.map, A.order, A.desc,
        plan.slots[slot].rowcodes.as<uint16_t>(), plan.slots[slot].colcodes.as<uint16_t>(), groups,
        (uint32_t)plan.C.n_rows, (uint32_t)plan.C.n_cols, io.flag, ctl);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    (*launches)++;
    const size_t smem = (size_t)plan.C.n_rows * plan.C.n_cols * IS_ROWW;
    const unsigned grid = (unsigned)((groups + IS_THREADS / 32 - 1) / (IS_THREADS / 32));
    const bool traced = !plan.C.score_only;
    if (!plan.C.affine) {
        if (plan.C.mode == IS_GLOBAL)
            e = traced ? interseq_launch_linear_fill<IS_GLOBAL, true>(plan.C, A, grid, smem, stream)
                       : interseq_launch_linear_fill<IS_GLOBAL, false>(plan.C, A, grid, smem, stream);
        else
            e = traced ? interseq_launch_linear_fill<IS_LOCAL, true>(plan.C, A, grid, smem, stream)
                       : interseq_launch_linear_fill<IS_LOCAL, false>(plan.C, A, grid, smem, stream);
        if (e == cudaSuccess) (*launches)++;
        return e;
    }
    if (traced && plan.C.dir_layout == 1) {
        const int sms = (int)(interseq_wave_groups() / IS_WARPS_PER_SM);
        switch (plan.C.mode) {
        case IS_GLOBAL: e = interseq_launch_tag_fill<IS_GLOBAL>(plan.C, A, ctl, sms, stream, launches); break;
        case IS_SEMI: e = interseq_launch_tag_fill<IS_SEMI>(plan.C, A, ctl, sms, stream, launches); break;
        default: e = interseq_launch_tag_fill<IS_LOCAL>(plan.C, A, ctl, sms, stream, launches); break;
        }
        return e;
    }
    switch (plan.C.mode) {
    case IS_GLOBAL:
        e = traced ? interseq_launch_fill<IS_GLOBAL, true>(plan.C, A, ctl, stream, launches)
                   : interseq_launch_fill<IS_GLOBAL, false>(plan.C, A, ctl, stream, launches);
        break;
    case IS_SEMI:
        e = traced ? interseq_launch_fill<IS_SEMI, true>(plan.C, A, ctl, stream, launches)
                   : interseq_launch_fill<IS_SEMI, false>(plan.C, A, ctl, stream, launches);
        break;
    default:
        e = traced ? interseq_launch_fill<IS_LOCAL, true>(plan.C, A, ctl, stream, launches)
                   : interseq_launch_fill<IS_LOCAL, false>(plan.C, A, ctl, stream, launches);
        break;
    }
    return e;
}

inline cudaError_t interseq_traceback_window(InterseqPlan &plan, const Window &win, int slot,
                                             const InterseqIO &io, cudaStream_t stream, int64_t *launches) {
    // semi-global scores are produced by the border pass even when no trace is wanted
    if (plan.C.score_only && plan.C.mode != IS_SEMI) return cudaSuccess;
    const InterseqArgs A = interseq_args(plan, win, slot, io);
    const int64_t slots = A.n_groups * 64;
    const unsigned grid = (unsigned)((slots + 127) / 128);
    const uint8_t *c1 = (const uint8_t *)io.codes1, *c2 = (const uint8_t *)io.codes2;
    if (!plan.C.affine) {
        if (plan.C.mode == IS_GLOBAL)
            interseq_linear_traceback_kernel<IS_GLOBAL><<<grid, 128, 0, stream>>>(plan.C, A, c1, c2, io.trace_len, io.first, io.ops, slots);
        else
            interseq_linear_traceback_kernel<IS_LOCAL><<<grid, 128, 0, stream>>>(plan.C, A, c1, c2, io.trace_len, io.first, io.ops, slots);
    } else if (plan.C.mode == IS_GLOBAL)
        interseq_traceback_kernel<IS_GLOBAL><<<grid, 128, 0, stream>>>(plan.C, A, c1, c2, io.trace_len, io.first, io.ops, slots);
    else if (plan.C.mode == IS_SEMI)
        interseq_traceback_kernel<IS_SEMI><<<grid, 128, 0, stream>>>(plan.C, A, c1, c2, io.trace_len, io.first, io.ops, slots);
    else
        interseq_traceback_kernel<IS_LOCAL><<<grid, 128, 0, stream>>>(plan.C, A, c1, c2, io.trace_len, io.first, io.ops, slots);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) (*launches)++;
    return e;
}

// ---------------------------------------------------------------------------------------
// Streaming plan for the pipelined one-shot path: worker threads plan the windows in order
// while the caller already copies and launches the windows that are ready.  Scratch is sized
// from the batch-wide longest sequences, because the exact per-window sizes are not known
// when the first window starts.
// ---------------------------------------------------------------------------------------
struct StreamingPlanner {
    InterseqPlan &plan;
    PairMap hm;
    int64_t n_pairs, per_window, n_windows;
    std::vector<int64_t> bounds;                // pair range of window w: [bounds[w], bounds[w + 1])
    int no_dirs, dir_words_per_lane;
    std::vector<WindowPlan> wps;
    std::vector<int64_t> group_begin;           // n_windows + 1
    std::unique_ptr<std::atomic<int>[]> ready;
    PinnedBuf staging;                          // page-locked copies of order / descriptors (async H2D)
    std::vector<std::thread> threads;
    std::atomic<int64_t> next{0};
    bool on_device;                             // windows are planned by device_plan.cuh on the compute stream

    StreamingPlanner(InterseqPlan &p, const PairMap &map, int64_t n, int waves, bool affine, bool device_plan)
        : plan(p), hm(map), n_pairs(n), no_dirs(p.C.score_only), dir_words_per_lane(affine ? 4 : 2),
          on_device(device_plan) {
        per_window = interseq_pairs_per_window(n_pairs, waves, p.warps_per_sm);
        // Windows are whole waves of the fill kernel (a partial wave leaves warps idle for the
        // 3 - 4 ms a group takes).  The first window is a single wave, so that the kernels start
        // while most of the input is still on its way; the rest is cut into equally many waves.
        bounds.push_back(0);
        const int64_t wave = interseq_wave_groups(p.warps_per_sm) * 64;
        if (n_pairs > 2 * wave && !getenv("B200ALIGN_WIN_GROUPS")) {
            bounds.push_back(wave);
            const int64_t rest = n_pairs - wave;
            const int64_t k = std::max<int64_t>(1, (rest + waves * wave - 1) / (waves * wave));
            const int64_t per = ((rest + k - 1) / k + wave - 1) / wave * wave;
            for (int64_t q = 1; q < k && wave + q * per < n_pairs; q++) bounds.push_back(wave + q * per);
        } else {
            for (int64_t q = per_window; q < n_pairs; q += per_window) bounds.push_back(q);
        }
        bounds.push_back(n_pairs);
        n_windows = (int64_t)bounds.size() - 1;
        wps.resize((size_t)n_windows);
        ready.reset(new std::atomic<int>[(size_t)n_windows]);
        group_begin.assign((size_t)n_windows + 1, 0);
        for (int64_t w = 0; w < n_windows; w++) {
            ready[(size_t)w].store(0);
            const int64_t cnt = bounds[(size_t)w + 1] - bounds[(size_t)w];
            group_begin[(size_t)w + 1] = group_begin[(size_t)w] + (cnt + 63) / 64;
        }
        plan.windows.assign((size_t)n_windows, Window{});
        plan.n_groups = group_begin[(size_t)n_windows];
    }
    ~StreamingPlanner() {
        for (std::thread &t : threads) t.join();
    }
    // device allocations: order / descriptors of all windows, `n_slots` scratch sets sized by bounds
    cudaError_t allocate(int n_slots, int64_t nmax, int64_t mmax) {
        int64_t gmax = 0;
        for (int64_t w = 0; w < n_windows; w++) gmax = std::max(gmax, group_begin[(size_t)w + 1] - group_begin[(size_t)w]);
        const int64_t stripes = (nmax + IS_R - 1) / IS_R;
        plan.max_dir_words = no_dirs ? 0 : gmax * stripes * mmax * 32 * dir_words_per_lane;
        plan.max_rowbuf = gmax * mmax * 32;
        plan.max_rowcodes = gmax * stripes * IS_R * 32;
        plan.max_colcodes = gmax * mmax * 32;
        cudaError_t e;
        if ((e = plan.d_order.alloc((size_t)plan.n_groups * 64 * 4)) != cudaSuccess) return e;
        if ((e = plan.d_desc.alloc((size_t)plan.n_groups * sizeof(WarpDesc))) != cudaSuccess) return e;
        if (!on_device && (e = staging.alloc((size_t)plan.n_groups * (64 * 4 + sizeof(WarpDesc)))) != cudaSuccess) return e;
        for (int sl = 0; sl < n_slots; sl++) {
            SlotBuffers &b = plan.slots[sl];
            if ((e = b.rowcodes.alloc((size_t)plan.max_rowcodes * 2)) != cudaSuccess) return e;
            if ((e = b.colcodes.alloc((size_t)(plan.max_colcodes + IS_SLACK * 32) * 2)) != cudaSuccess) return e;
            if ((e = b.rowbuf.alloc((size_t)(plan.max_rowbuf + IS_SLACK * 32) * 12)) != cudaSuccess) return e;
            if ((e = b.dirs.alloc((size_t)plan.max_dir_words * 4)) != cudaSuccess) return e;
            if ((e = b.ctl.alloc(16)) != cudaSuccess) return e;
            if (plan.C.mode == IS_SEMI) {
                if ((e = b.lastrow.alloc((size_t)plan.max_rowbuf * 8)) != cudaSuccess) return e;
                if ((e = b.lastcol.alloc((size_t)plan.max_rowcodes * 8)) != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    }
    void start() {
        if (on_device) return;
        const unsigned hw = interseq_host_threads();
        const int n_threads = (int)std::min<int64_t>(std::min<unsigned>(hw, 8u), n_windows);
        for (int t = 0; t < n_threads; t++) {
            threads.emplace_back([this]() {
                for (;;) {
                    const int64_t w = next.fetch_add(1);
                    if (w >= n_windows) break;
                    const int64_t w0 = bounds[(size_t)w], w1 = bounds[(size_t)w + 1];
                    WindowPlan &wp = wps[(size_t)w];
                    interseq_plan_window_records(wp, hm, w0, w1, no_dirs, dir_words_per_lane);
                    memcpy(stage_order() + group_begin[(size_t)w] * 64, wp.order.data(), wp.order.size() * 4);
                    memcpy(stage_desc() + group_begin[(size_t)w], wp.desc.data(), wp.desc.size() * sizeof(WarpDesc));
                    ready[(size_t)w].store(1, std::memory_order_release);
                }
            });
        }
    }
    // blocks until window w is planned; uploads its order / descriptors on `stream`
    cudaError_t take(int64_t w, cudaStream_t stream) {
        WindowPlan &wp = wps[(size_t)w];
        if (on_device) {
            // pair and group ranges follow from the counts alone; the rest is built on the device
            wp.win.pair_begin = bounds[(size_t)w]; wp.win.pair_end = bounds[(size_t)w + 1];
            wp.win.group_begin = group_begin[(size_t)w];
            wp.win.group_end = group_begin[(size_t)w + 1];
            plan.windows[(size_t)w] = wp.win;
            return cudaSuccess;
        }
        while (!ready[(size_t)w].load(std::memory_order_acquire)) std::this_thread::yield();
        wp.win.group_begin = group_begin[(size_t)w];
        wp.win.group_end = group_begin[(size_t)w + 1];
        plan.windows[(size_t)w] = wp.win;
        cudaError_t e = cudaMemcpyAsync(plan.d_order.as<int32_t>() + wp.win.group_begin * 64,
                                        stage_order() + wp.win.group_begin * 64, wp.order.size() * 4,
                                        cudaMemcpyHostToDevice, stream);
        if (e != cudaSuccess) return e;
        return cudaMemcpyAsync(plan.d_desc.as<WarpDesc>() + wp.win.group_begin, stage_desc() + wp.win.group_begin,
                               wp.desc.size() * sizeof(WarpDesc), cudaMemcpyHostToDevice, stream);
    }
    int32_t *stage_order() const { return staging.as<int32_t>(); }
    WarpDesc *stage_desc() const { return (WarpDesc *)(staging.as<char>() + (size_t)plan.n_groups * 64 * 4); }
};

}  // namespace bt
