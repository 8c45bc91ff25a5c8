// Intra-pair wavefront kernel: the int32, reference-complete fill for long pairs, few pairs and
// the single-pair API (align_optimal / align_banded called on one pair).
//
// general_fill_kernel sweeps anti-diagonals with one __syncthreads per wavefront (about 1.2 us
// each): correct for every mode, but a 1368 x 1345 pair needs 2713 barriers.  Here a pair is
// cut into stripes of 32 R rows, one warp per stripe in flight.  Lane t of a warp owns R
// consecutive rows and walks the columns with a skew of one column per lane (at step k it is at
// column k - t), so the cell above its first row was finished by lane t - 1 in the previous
// step and arrives by __shfl_up_sync together with the diagonal one; the left neighbours are the
// lane's own registers.  No block barrier in the loop: warps of consecutive stripes run
// concurrently, the lower one a few dozen columns behind, and hand over the stripe's last row
// through a buffer (shared memory when it fits) guarded by a progress counter per stripe.
//
// The recurrence, the boundary values and the direction sets are those of the reference
// (pairwise.rs:378-415, :608-681; banded.rs:348-375, :486-548) in wrapping int32 arithmetic
// with the reference's -inf, so every mode stays bit-identical: linear / affine, global /
// semi-global / local, matrix or match / mismatch, any code width, full table or band (the
// band's straightened table of banded.rs:164-255 is addressed as c = seq_j + 1 - (i - 1) -
// lower).  Direction bytes are full tie sets (cell.rs:44-80) so that the co-optimal enumeration
// of trace.rs:21-92 works on the result; a lane collects four columns per row in a register and
// stores them as one aligned word (row stride padded to 16).  The waived gap penalties of the
// semi-global mode (pairwise.rs:482-530, :795-843) are per-row / per-step operands of the same
// cell code, not branches in it.
//
// Table row 0 and column 0 are closed forms of pairwise.rs:430-467 / :695-780 (valid while no
// sum can wrap, which the router checks: BtBatch::wavefront).
#pragma once
#include "common.cuh"
#include <type_traits>

#include "general_kernel.cuh"

namespace bt {

constexpr int WF_MAX_WARPS = 16;
constexpr int WF_CHUNK = 8;          // steps between progress updates

struct WfArgs {
    int32_t *bnd;        // global scratch for the stripe boundaries (nullptr: shared memory)
    int64_t bnd_block;   // ints per block in `bnd`
    int32_t bnd_cols;    // columns per boundary slot
    int32_t mat_smem;    // matrix staged in shared memory
    int32_t lean_rows;   // rows per lane of the lean kernel (strip_kernel.cuh): 4 or 8
};

// closed forms of the first row (pairwise.rs:452-467, :747-780) and first column (:434-450, :707-745)
template <bool AFFINE>
__device__ __forceinline__ Cell3 wf_row0(const DevParams &P, int j, uint32_t &dir) {
    Cell3 c;
    dir = 0;
    const int32_t NI = P.neg_inf;
    if (AFFINE) {
        if (j == 0) { c.m = 0; c.g1 = NI; c.g2 = NI; return c; }
        c.m = NI; c.g2 = NI;
        if (P.local) { c.g1 = 0; return c; }
        c.g1 = P.terminal ? wadd(P.gap_open, (int32_t)((uint32_t)(j - 1) * (uint32_t)P.gap_ext)) : 0;
        dir = j == 1 ? A_MG1 : A_G1G1;
    } else {
        c.g1 = c.g2 = 0;
        if (j == 0 || P.local) { c.m = 0; return c; }
        c.m = P.terminal ? (int32_t)((uint32_t)j * (uint32_t)P.gap_open) : 0;
        dir = L_GAP1;
    }
    return c;
}
template <bool AFFINE>
__device__ __forceinline__ Cell3 wf_col0(const DevParams &P, int i, uint32_t &dir) {
    Cell3 c;
    dir = 0;
    const int32_t NI = P.neg_inf;
    if (AFFINE) {
        if (i == 0) { c.m = 0; c.g1 = NI; c.g2 = NI; return c; }
        c.m = NI; c.g1 = NI;
        if (P.local) { c.g2 = 0; return c; }
        c.g2 = P.terminal ? wadd(P.gap_open, (int32_t)((uint32_t)(i - 1) * (uint32_t)P.gap_ext)) : 0;
        dir = i == 1 ? A_MG2 : A_G2G2;
    } else {
        c.g1 = c.g2 = 0;
        if (i == 0 || P.local) { c.m = 0; return c; }
        c.m = P.terminal ? (int32_t)((uint32_t)i * (uint32_t)P.gap_open) : 0;
        dir = L_GAP2;
    }
    return c;
}

template <bool AFFINE>
__device__ __forceinline__ Cell3 wf_shfl_up(const Cell3 &c) {
    Cell3 r;
    r.m = __shfl_up_sync(0xffffffffu, c.m, 1);
    r.g1 = AFFINE ? __shfl_up_sync(0xffffffffu, c.g1, 1) : 0;
    r.g2 = AFFINE ? __shfl_up_sync(0xffffffffu, c.g2, 1) : 0;
    return r;
}
template <bool AFFINE>
__device__ __forceinline__ Cell3 wf_shfl(const Cell3 &c, int src) {
    Cell3 r;
    r.m = __shfl_sync(0xffffffffu, c.m, src);
    r.g1 = AFFINE ? __shfl_sync(0xffffffffu, c.g1, src) : 0;
    r.g2 = AFFINE ? __shfl_sync(0xffffffffu, c.g2, src) : 0;
    return r;
}

enum : int { WF_GLOBAL = 0, WF_SEMI = 1, WF_LOCAL = 2 };

// MODE: WF_GLOBAL (terminal gaps penalised; every band), WF_SEMI (full table, gap penalties waived
// in the last row / column, pairwise.rs:482-530, :795-843), WF_LOCAL.
template <bool AFFINE, bool BANDED, int MODE, bool TRACED, int R>
__global__ void __launch_bounds__(32 * WF_MAX_WARPS)
wavefront_fill_kernel(const DevParams P, const BatchView V, const WfArgs W_) {
    constexpr bool LOCAL = MODE == WF_LOCAL;
    constexpr bool SEMI = MODE == WF_SEMI;
    extern __shared__ int32_t wf_smem[];
    __shared__ unsigned long long prog[WF_MAX_WARPS];
    __shared__ int32_t red_score[WF_MAX_WARPS];
    __shared__ int64_t red_idx[WF_MAX_WARPS];

    const int64_t p = V.pair_begin + blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    const int64_t o1 = V.map.beg1(p), o2 = V.map.beg2(p);
    const int n = (int)V.map.len1(p), m = (int)V.map.len2(p);
    constexpr int NS = AFFINE ? 3 : 1;
    constexpr int ROWS = 32 * R;
    const int32_t NI = P.neg_inf;

    int lower = 0, upper = 0, W = 0;
    if (BANDED) {
        lower = (int)V.bands[2 * p]; upper = (int)V.bands[2 * p + 1];
        W = upper - lower + 1;
    }
    const int64_t rs = table_stride(1, BANDED ? (int64_t)W + 2 : (int64_t)m + 1);
    uint8_t *const dir = TRACED ? V.dirs + V.dir_off[p] : nullptr;
    int32_t *const cand = (BANDED && !LOCAL) ? V.cand + V.cand_off[p] : nullptr;
    const uint32_t mark_bit = P.mark ? (uint32_t)DIR_MARK : 0u;

    // shared memory: [matrix] [boundary slots]
    const int32_t *mat = V.matrix;
    int32_t *smem_next = wf_smem;
    if (W_.mat_smem) {
        for (int k = tid; k < P.n_rows * P.n_cols; k += nthr) smem_next[k] = V.matrix[k];
        mat = smem_next;
        smem_next += P.n_rows * P.n_cols;
    }
    volatile int32_t *const bnd = W_.bnd ? W_.bnd + (int64_t)blockIdx.x * W_.bnd_block : smem_next;
    const int bnd_cols = W_.bnd_cols;
    if (tid < WF_MAX_WARPS) prog[tid] = 0ull;

    // table row 0 (full table; the band's row 0 is never stored) and the candidates' defaults
    if (!BANDED && TRACED) {
        for (int j = tid; j <= m; j += nthr) {
            uint32_t dd;
            const Cell3 c = wf_row0<AFFINE>(P, j, dd);
            dir[j] = (uint8_t)(dd | (c.m == P.mark_value ? mark_bit : 0u));
        }
    }
    if (cand) {
        for (int k = tid; k < W + 2; k += nthr) {
            cand[k] = 0;
            if (AFFINE) { cand[(W + 2) + k] = NI; cand[2 * (W + 2) + k] = NI; }
        }
    }
    __syncthreads();

    // local: first maximum in row-major order (pairwise.rs:536-551, :849-865); the scan starts
    // at the origin (band: at the first edge cell (0, 1)), whose M is 0
    int32_t best = LOCAL ? 0 : INT32_MIN;
    int best_i = 0, best_c = BANDED ? 1 : 0;
    const Cell3 EDGE = {0, AFFINE ? NI : 0, AFFINE ? NI : 0};  // banded.rs:334-337, :462-472
    const Cell3 SENT = {NI, AFFINE ? NI : 0, AFFINE ? NI : 0}; // banded.rs:339-345, :474-484
    const int32_t open = P.gap_open, ext = AFFINE ? P.gap_ext : P.gap_open;
    const int n_stripes = (n + ROWS - 1) / ROWS;
    // columns walked by a lane: full table 0 .. (m | 3) (the padding completes the last word of
    // direction bytes); band: W + R - 1 columns of its R rows' windows, plus 3 for the same reason
    const int lane_cols = BANDED ? W + R - 1 + 3 : (m | 3) + 1;
    const int lane_skew = BANDED ? R + 1 : 1;
    const int n_steps = 31 * lane_skew + lane_cols;
    // producer step that must be complete before lane 0 of the next stripe runs its step k
    const int lead = (BANDED ? ROWS : 0) + 32;
    const bool byte_codes = P.code_bytes == 1;

    for (int q = warp; q < n_stripes; q += nw) {
        const int base = q * ROWS;                 // rows base + 1 .. base + ROWS
        const int i0 = base + lane * R + 1;        // first row of this lane
        volatile int32_t *const bnd_out = bnd + (int64_t)(q % nw) * bnd_cols * NS;
        volatile int32_t *const bnd_in = bnd + (int64_t)((q + nw - 1) % nw) * bnd_cols * NS;
        // seq column (0-based) of lane 0 at step 0; lane t at step k: sj = sj0 + k - t
        const int sj0 = BANDED ? base + lower : -1;
        const uint8_t *const col_bytes = (const uint8_t *)V.codes2 + o2 + (sj0 - lane);

        // per-row constants and state
        const int32_t *rowmat[R];
        uint64_t rowcode[R];
        Cell3 cur[R];
        uint32_t dw[R];
        uint8_t *dir_row[R];     // direction bytes of the lane's rows
        int32_t o1r[R], e1r[R];  // semi-global: penalties of a horizontal step, waived in the last row
        int32_t simn[R];         // substitution scores of the next step (fetched one step ahead)
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = i0 + r;
            rowcode[r] = i <= n ? load_code(V.codes1, o1 + i - 1, P.code_bytes) : 0;
            rowmat[r] = mat + rowcode[r] * (uint64_t)P.n_cols;
            cur[r] = BANDED ? SENT : EDGE;
            dw[r] = 0;
            dir_row[r] = TRACED ? dir + (int64_t)i * rs : nullptr;
            o1r[r] = (SEMI && i == n) ? 0 : open;
            e1r[r] = (SEMI && i == n) ? 0 : ext;
            simn[r] = 0;
        }
        // substitution scores of the lane's rows against the column of step k (0 outside sequence 2)
        auto fetch_sims = [&](int k) {
            const int sj = sj0 + k - lane;
            if (sj >= 0 && sj < m) {
                const uint64_t s2 = byte_codes ? (uint64_t)col_bytes[k] : load_code(V.codes2, o2 + sj, P.code_bytes);
                if (P.use_matrix) {
#pragma unroll
                    for (int r = 0; r < R; r++) simn[r] = rowmat[r][(uint32_t)s2];
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) simn[r] = rowcode[r] == s2 ? P.match : P.mismatch;
                }
            }
        };
        fetch_sims(0);
        const bool rows_in = i0 + R - 1 <= n;      // all rows of the lane are table rows
        Cell3 recv_prev = SENT;   // cell above the lane's first row, one column back
        // state of the row above the stripe as this lane 0 sees it
        auto above = [&](int sj) -> Cell3 {
            const int i = base;    // table row of the cell
            if (BANDED) {
                if (sj < i - 1 + lower || sj > i - 1 + upper) return SENT;
                if (i == 0 || sj < 0) return EDGE;
                if (sj >= m) return SENT;
            } else {
                if (sj + 1 < 0 || sj + 1 > m) return EDGE;   // beside the table (the value is never used)
                if (i == 0) {
                    uint32_t dd;
                    return wf_row0<AFFINE>(P, sj + 1, dd);
                }
            }
            Cell3 c;
            const int64_t at = (int64_t)(sj + 1) * NS;
            c.m = bnd_in[at];
            c.g1 = AFFINE ? bnd_in[at + 1] : 0;
            c.g2 = AFFINE ? bnd_in[at + 2] : 0;
            return c;
        };
        const unsigned long long want_tag = (unsigned long long)(q > 0 ? q - 1 : 0) << 32;

        for (int k0 = 0; k0 < n_steps; k0 += WF_CHUNK) {
            if (q > 0) {
                // wait until the stripe above has produced what this chunk reads
                if (lane == 0) {
                    const unsigned long long want = want_tag | (unsigned long long)(k0 + WF_CHUNK - 1 + lead);
                    const volatile unsigned long long *flag = &prog[(q - 1) % nw];
                    while (*flag < want) __nanosleep(100);
                    __threadfence_block();
                }
                __syncwarp();
            }
            const int k1 = min(k0 + WF_CHUNK, n_steps);
            // the cells above lane 0 for the whole chunk: lane l fetches the one of step k0 + l (one
            // memory latency per chunk instead of one per step)
            Cell3 pre = SENT;
            if (lane < WF_CHUNK) pre = above(sj0 + k0 + lane);
            if (k0 == 0) recv_prev = lane == 0 ? above(sj0 - 1) : (BANDED ? SENT : EDGE);
            for (int k = k0; k < k1; k++) {
                const int u = k - lane * lane_skew;          // column index inside the lane's window
                const int sj = sj0 + k - lane;               // 0-based column of sequence 2 (-1: table column 0)
                Cell3 recv = wf_shfl_up<AFFINE>(cur[R - 1]);
                const Cell3 first_above = wf_shfl<AFFINE>(pre, k - k0);
                if (lane == 0) recv = first_above;
                const bool active = u >= 0 && u < lane_cols;
                const bool in_seq = sj >= 0 && sj < m;
                int32_t sim[R];
#pragma unroll
                for (int r = 0; r < R; r++) sim[r] = simn[r];
                if (k + 1 < n_steps) fetch_sims(k + 1);      // in flight during this step
                if (active && (BANDED || sj >= 0)) {
                    // penalties of a vertical step (waived in the last column, semi-global)
                    const int32_t o2v = (SEMI && sj == m - 1) ? 0 : open, e2v = (SEMI && sj == m - 1) ? 0 : ext;
                    uint32_t dd_row[R];
                    // The R cells of this step (pairwise.rs:378-415 linear, :608-681 affine), in two
                    // phases: M and G1 of every row depend only on the previous column and are
                    // independent of each other; only G2 (linear: the vertical candidate) runs down
                    // the rows.  Cells outside the band / the sequence take their fixed values by
                    // selects, so the step is one basic block.
                    Cell3 nw_[R];
                    bool filled[R];
                    uint32_t fixed_m[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const Cell3 left = cur[r];
                        const Cell3 diag = r == 0 ? recv_prev : cur[r - 1];
                        const int32_t ho = SEMI ? o1r[r] : open, he = SEMI ? e1r[r] : ext;
                        bool sent = false, edge = false, keep = false;
                        if (BANDED) {
                            const int c = sj - (i0 + r - 1) - lower + 1;
                            sent = !((unsigned)(c - 1) < (unsigned)W) || sj >= m;   // outside the band: sentinel
                            edge = !sent && sj < 0;                                  // in the band, left of the sequence
                        } else {
                            keep = !in_seq;                                          // padding columns behind the table
                        }
                        filled[r] = !(sent || edge || keep);
                        uint32_t d = 0;
                        Cell3 out;
                        if (AFFINE) {
                            const int32_t mm = wadd(diag.m, sim[r]), g1m = wadd(diag.g1, sim[r]), g2m = wadd(diag.g2, sim[r]);
                            const int32_t mg1 = wadd(left.m, ho), g1g1 = wadd(left.g1, he);
                            out.m = __vimax3_s32(mm, g1m, g2m);
                            out.g1 = max(mg1, g1g1);
                            d = (mm == out.m ? (uint32_t)A_MM : 0u) | (g1m == out.m ? (uint32_t)A_G1M : 0u) |
                                (g2m == out.m ? (uint32_t)A_G2M : 0u);
                            d |= (mg1 >= g1g1 ? (uint32_t)A_MG1 : 0u) | (g1g1 >= mg1 ? (uint32_t)A_G1G1 : 0u);
                            if (LOCAL) {
                                if (out.m <= 0) { d &= ~(uint32_t)(A_MM | A_G1M | A_G2M); out.m = 0; }
                                if (out.g1 <= 0) { d &= ~(uint32_t)(A_MG1 | A_G1G1); out.g1 = NI; }
                            }
                            out.g2 = 0;
                            if (BANDED) {
                                if (sent) { out.m = NI; out.g1 = NI; d = 0; }
                                if (edge) { out.m = 0; out.g1 = NI; d = 0; }
                            } else if (keep) { out.m = left.m; out.g1 = left.g1; d = 0; }
                        } else {
                            // linear: diagonal and horizontal candidates now, the vertical one in the chain
                            out.m = max(wadd(diag.m, sim[r]), wadd(left.m, ho));
                            out.g1 = out.g2 = 0;
                        }
                        nw_[r] = out;
                        fixed_m[r] = d;
                    }
                    {
                        int32_t up_m = recv.m, up_g2 = recv.g2;
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const Cell3 left = cur[r];
                            Cell3 out = nw_[r];
                            uint32_t d = fixed_m[r];
                            bool sent = false, edge = false, keep = false;
                            if (BANDED) {
                                const int c = sj - (i0 + r - 1) - lower + 1;
                                sent = !((unsigned)(c - 1) < (unsigned)W) || sj >= m;
                                edge = !sent && sj < 0;
                            } else {
                                keep = !in_seq;
                            }
                            if (AFFINE) {
                                const int32_t mg2 = wadd(up_m, o2v), g2g2 = wadd(up_g2, e2v);
                                int32_t g2 = max(mg2, g2g2);
                                uint32_t dg = (mg2 >= g2g2 ? (uint32_t)A_MG2 : 0u) | (g2g2 >= mg2 ? (uint32_t)A_G2G2 : 0u);
                                if (LOCAL && g2 <= 0) { dg = 0; g2 = NI; }
                                if (BANDED) { if (sent || edge) { g2 = NI; dg = 0; } }
                                else if (keep) { g2 = left.g2; dg = 0; }
                                out.g2 = g2;
                                d |= dg;
                            } else {
                                const int32_t a = wadd((r == 0 ? recv_prev : cur[r - 1]).m, sim[r]);
                                const int32_t b = wadd(left.m, SEMI ? o1r[r] : open), c2 = wadd(up_m, o2v);
                                int32_t mx = max(out.m, c2);
                                d = (a == mx ? (uint32_t)L_MATCH : 0u) | (b == mx ? (uint32_t)L_GAP1 : 0u) |
                                    (c2 == mx ? (uint32_t)L_GAP2 : 0u);
                                if (LOCAL && mx <= 0) { mx = 0; d = 0; }
                                if (BANDED) {
                                    if (sent) { mx = NI; d = 0; }
                                    if (edge) { mx = 0; d = 0; }
                                } else if (keep) { mx = left.m; d = 0; }
                                out.m = mx;
                            }
                            if (LOCAL && filled[r] && out.m == P.mark_value) d |= mark_bit;
                            dd_row[r] = d;
                            nw_[r] = out;
                            up_m = out.m; up_g2 = out.g2;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < R; r++) cur[r] = nw_[r];
                    // ---- everything that depends on where the cells are (per step, not per cell)
                    if (LOCAL && in_seq) {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const int i = i0 + r;
                            const int c = BANDED ? sj - (i - 1) - lower + 1 : sj + 1;
                            const bool ok = i <= n && (!BANDED || (unsigned)(c - 1) < (unsigned)W);
                            // a later cell with the same score comes first in row-major order only in a higher row
                            if (ok && (cur[r].m > best || (cur[r].m == best && i < best_i))) { best = cur[r].m; best_i = i; best_c = c; }
                        }
                    }
                    if (TRACED) {
#pragma unroll
                        for (int r = 0; r < R; r++) {
                            const int i = i0 + r;
                            dw[r] = __byte_perm(dw[r], dd_row[r], 0x4321);
                            if (BANDED) {
                                const int c = sj - (i - 1) - lower + 1;
                                if ((c & 3) == 3 && c >= 3 && c < rs && i <= n) *(uint32_t *)(dir_row[r] + (c - 3)) = dw[r];
                            } else {
                                const int j = sj + 1;
                                if ((j & 3) == 3 && (rows_in || i <= n)) *(uint32_t *)(dir_row[r] + (j - 3)) = dw[r];
                            }
                        }
                    }
                    if (!LOCAL && in_seq) {
                        if (BANDED) {
                            // end candidates: last row, or last column of sequence 2 (banded.rs:598-638)
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                const int i = i0 + r;
                                const int c = sj - (i - 1) - lower + 1;
                                if (i <= n && (unsigned)(c - 1) < (unsigned)W && (i == n || sj == m - 1)) {
                                    cand[c] = cur[r].m;
                                    if (AFFINE) { cand[(W + 2) + c] = cur[r].g1; cand[2 * (W + 2) + c] = cur[r].g2; }
                                }
                            }
                        } else if (sj == m - 1) {
                            // optimum and traceback start of the global modes (pairwise.rs:532-559, :845-881)
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                if (i0 + r != n) continue;
                                const Cell3 out = cur[r];
                                int32_t mx = out.m;
                                int st = ST_MATCH;
                                if (AFFINE) {
                                    mx = max(out.m, max(out.g1, out.g2));
                                    st = (out.m == mx) ? ST_MATCH : ((out.g1 == mx) ? ST_GAP1 : ST_GAP2);
                                }
                                V.score[p] = mx;
                                V.start_idx[p] = (int64_t)n * rs + m;
                                V.start_state[p] = st;
                                V.end_vals[3 * p] = out.m; V.end_vals[3 * p + 1] = out.g1; V.end_vals[3 * p + 2] = out.g2;
                            }
                        }
                    }
                } else if (active) {
                    // table column 0 of the full table (pairwise.rs:434-450, :707-745)
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        uint32_t dd;
                        cur[r] = wf_col0<AFFINE>(P, i0 + r, dd);
                        if (LOCAL && cur[r].m == P.mark_value) dd |= mark_bit;
                        if (TRACED) dw[r] = __byte_perm(dw[r], dd, 0x4321);
                    }
                } else if (BANDED) {
                    // outside the lane's window every cell is outside the band
#pragma unroll
                    for (int r = 0; r < R; r++) cur[r] = SENT;
                }
                recv_prev = recv;
                // the stripe's last row goes to the next stripe
                if (lane == 31 && active && sj + 1 >= 0 && sj + 1 < bnd_cols) {
                    const int64_t at = (int64_t)(sj + 1) * NS;
                    bnd_out[at] = cur[R - 1].m;
                    if (AFFINE) { bnd_out[at + 1] = cur[R - 1].g1; bnd_out[at + 2] = cur[R - 1].g2; }
                }
            }
            if (q + 1 < n_stripes) {
                __syncwarp();
                if (lane == 31) {
                    __threadfence_block();
                    *(volatile unsigned long long *)&prog[q % nw] =
                        ((unsigned long long)q << 32) | (unsigned long long)(k1 >= n_steps ? 0xffffffffu : (unsigned)k1);
                }
            }
        }
    }
    __syncthreads();

    // ---- optimum and traceback start (pairwise.rs:532-559, :845-881; banded.rs:377-414, :550-593)
    if (LOCAL) {
        Best bb;
        bb.score = best; bb.idx = (int64_t)best_i * rs + best_c;
        for (int off = 16; off > 0; off >>= 1) {
            int32_t os = __shfl_down_sync(0xffffffffu, bb.score, off);
            int64_t oi = __shfl_down_sync(0xffffffffu, bb.idx, off);
            bb.offer(os, oi);
        }
        if (lane == 0) { red_score[warp] = bb.score; red_idx[warp] = bb.idx; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < nw; k++) bb.offer(red_score[k], red_idx[k]);
            V.score[p] = bb.score;
            V.start_idx[p] = bb.idx;
            V.start_state[p] = ST_MATCH;
            V.end_vals[3 * p] = bb.score; V.end_vals[3 * p + 1] = NI; V.end_vals[3 * p + 2] = NI;
        }
    } else if (!BANDED) {
        if (tid == 0 && (n == 0 || m == 0)) {
            // the end cell lies on the table's border
            uint32_t dd;
            const Cell3 c = n == 0 ? wf_row0<AFFINE>(P, m, dd) : wf_col0<AFFINE>(P, n, dd);
            int32_t mx = c.m;
            int st = ST_MATCH;
            if (AFFINE) {
                mx = max(c.m, max(c.g1, c.g2));
                st = (c.m == mx) ? ST_MATCH : ((c.g1 == mx) ? ST_GAP1 : ST_GAP2);
            }
            V.score[p] = mx;
            V.start_idx[p] = (int64_t)n * rs + m;
            V.start_state[p] = st;
            V.end_vals[3 * p] = c.m; V.end_vals[3 * p + 1] = c.g1; V.end_vals[3 * p + 2] = c.g2;
        }
    } else {
        if (tid == 0) {
            // banded.rs:598-638: candidate of band column j
            int32_t mx = INT32_MIN;
            for (int j = 1; j <= W; j++) {
                mx = max(mx, cand[j]);
                if (AFFINE) mx = max(mx, max(cand[(W + 2) + j], cand[2 * (W + 2) + j]));
            }
            int st_found = -1; int64_t idx_found = 0;
            for (int st = 0; st < NS && st_found < 0; st++) {
                for (int j = 1; j <= W; j++) {
                    if (cand[st * (W + 2) + j] == mx) {
                        const int sj = j + (n - 1) + lower - 1;
                        const int i = sj < m ? n : m - j - lower + 1;
                        st_found = st; idx_found = (int64_t)i * rs + j;
                        break;
                    }
                }
            }
            V.score[p] = mx;
            V.start_idx[p] = idx_found;
            V.start_state[p] = st_found < 0 ? 0 : st_found;
            V.end_vals[3 * p] = mx; V.end_vals[3 * p + 1] = NI; V.end_vals[3 * p + 2] = NI;
        }
    }
}

// ---------------------------------------------------------------------------------------
// First-optimal traceback of one pair by one warp (trace.rs:59-90 with max_number = 1; step order
// cell.rs:125-138, :219-254).  general_traceback_kernel walks a table with one thread and pays a
// global-memory latency per step (0.4 ms for the 2700 steps of a Cas9-sized pair); here the warp
// keeps a 32 x 128 byte tile of the direction table in shared memory (one coalesced load per
// tile, loaded again only when the walk leaves it).  Full tables: all lanes advance the walk in
// lockstep and take runs of Match-to-Match diagonal steps 32 cells at a time (see below); bands:
// lane 0 walks inside the tile.  Same outputs as general_traceback_kernel.
// ---------------------------------------------------------------------------------------
constexpr int WT_ROWS = 32, WT_COLS = 128;

// LEAN: the table was written by strip_fill_kernel (strip_kernel.cuh): bit 0 = Gap1 opened from
// Match, bit 1 = Gap2 opened from Match, bits 2-4 = state the Match value came from (7 Match,
// 6 Gap1, 4 Gap2, 0: the local alignment starts here), no bytes for table row 0 / column 0, which
// are walked by rule (pairwise.rs:707-780: a leading gap down to the origin; local: the end).
template <bool AFFINE, bool BANDED, bool LEAN = false>
__global__ void __launch_bounds__(128)
wavefront_traceback_kernel(const DevParams P, const BatchView V) {
    __shared__ uint4 tiles[4][WT_ROWS * WT_COLS / 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t p = V.pair_begin + (int64_t)blockIdx.x * 4 + warp;
    if (p >= V.pair_end) return;
    const int64_t n = V.map.len1(p), m = V.map.len2(p);
    int64_t rs = table_stride(P.dir_pad, m + 1);
    if (BANDED) rs = table_stride(P.dir_pad, V.bands[2 * p + 1] - V.bands[2 * p] + 3);
    const int64_t n_rows = n + 1;
    const uint8_t *dir = V.dirs + V.dir_off[p];
    uint8_t *ops = P.stats ? nullptr : V.ops + V.map.ops(p);
    uint8_t *tile = (uint8_t *)tiles[warp];
    GapStats gs(P.terminal != 0);
    const int64_t cap = n + m;
    const int64_t start = V.start_idx[p];
    int64_t i = start / rs, j = start % rs, last_i = 0, last_j = 0, len = 0;
    int st = V.start_state[p];
    int64_t ti = -1, tj = -1;      // origin of the cached tile
    if (!BANDED) {
        // Full tables: every lane holds the walk's state (i, j, st, len) and they advance it
        // together.  The tile is anchored with the current cell in its bottom-right corner (the walk
        // only moves up and left).  In state Match, lane t looks at the cell t steps up the diagonal:
        // the leading lanes whose cells say "Match came from Match" are one run of diagonal steps,
        // taken in one go (ballot + find-first-set) - most of an alignment of similar sequences.
        // Everything else is one generic step, decided identically by all lanes from a broadcast
        // shared-memory read.
        for (;;) {
            if (i < 0 || j < 0 || len >= cap) break;
            if (ti < 0 || i < ti || i >= ti + WT_ROWS || j < tj || j >= tj + WT_COLS) {
                ti = max((int64_t)0, i - (WT_ROWS - 1));
                tj = max((int64_t)0, ((j + 16) & ~(int64_t)15) - WT_COLS);
                __syncwarp();
                for (int q = lane; q < WT_ROWS * (WT_COLS / 16); q += 32) {
                    const int r = q >> 3, ch = q & 7;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (ti + r < n_rows && tj + ch * 16 < rs)
                        v = *(const uint4 *)(dir + (ti + r) * rs + tj + ch * 16);
                    tiles[warp][q] = v;
                }
                __syncwarp();
            }
            if (st == ST_MATCH && i >= 1 && j >= 1) {
                // cells (i - t, j - t), t = lane: inside the table's interior and inside the tile?
                const int64_t ii = i - lane, jj = j - lane;
                bool go = ii >= 1 && jj >= 1 && ii >= ti && jj >= tj;
                if (go) {
                    const uint32_t dd = tile[(ii - ti) * WT_COLS + (jj - tj)];
                    go = LEAN ? ((dd >> 2) & 7u) == 7u : (AFFINE ? (dd & A_MM) != 0 : (dd & L_MATCH) != 0);
                }
                const unsigned stop = ~__ballot_sync(0xffffffffu, go);
                int run = stop ? __ffs((int)stop) - 1 : 32;          // leading lanes that continue the diagonal in Match
                run = (int)min((int64_t)run, cap - len);
                if (run > 0) {
                    if (P.stats) { gs.push(OP_DIAG); gs.len += run - 1; }
                    else if (lane < run) ops[len + lane] = (uint8_t)OP_DIAG;
                    len += run;
                    last_i = i - (run - 1); last_j = j - (run - 1);
                    i -= run; j -= run;
                    continue;
                }
            }
            const uint32_t d = tile[(i - ti) * WT_COLS + (j - tj)];
            int op, ns = 0;
            bool done = false;
            if (LEAN) {
                if (i == 0 || j == 0) {
                    if (P.local || (i == 0 && j == 0)) done = true;
                    if (i == 0) { op = OP_TO1; ns = ST_GAP1; } else { op = OP_TO2; ns = ST_GAP2; }
                } else if (st == ST_MATCH) {
                    const uint32_t tag = (d >> 2) & 7u;
                    if (tag == 0) done = true;
                    op = OP_DIAG; ns = tag == 7u ? ST_MATCH : (tag == 6u ? ST_GAP1 : ST_GAP2);
                } else if (st == ST_GAP1) {
                    op = OP_TO1; ns = (d & 1u) ? ST_MATCH : ST_GAP1;
                } else {
                    op = OP_TO2; ns = (d & 2u) ? ST_MATCH : ST_GAP2;
                }
            } else if (AFFINE) {
                op = OP_DIAG;
                if (st == ST_MATCH) {
                    if (d & A_MM) { op = OP_DIAG; ns = ST_MATCH; }
                    else if (d & A_G1M) { op = OP_DIAG; ns = ST_GAP1; }
                    else if (d & A_G2M) { op = OP_DIAG; ns = ST_GAP2; }
                    else done = true;
                } else if (st == ST_GAP1) {
                    if (d & A_MG1) { op = OP_TO1; ns = ST_MATCH; }
                    else if (d & A_G1G1) { op = OP_TO1; ns = ST_GAP1; }
                    else done = true;
                } else {
                    if (d & A_MG2) { op = OP_TO2; ns = ST_MATCH; }
                    else if (d & A_G2G2) { op = OP_TO2; ns = ST_GAP2; }
                    else done = true;
                }
            } else {
                op = OP_DIAG;
                if (d & L_MATCH) op = OP_DIAG;
                else if (d & L_GAP1) op = OP_TO1;
                else if (d & L_GAP2) op = OP_TO2;
                else done = true;
            }
            if (done) break;
            if (P.stats) gs.push(op); else if (lane == 0) ops[len] = (uint8_t)op;
            len++;
            last_i = i; last_j = j;
            if (op == OP_DIAG) { i--; j--; }
            else if (op == OP_TO1) { j--; }
            else { i--; }
            st = ns;
        }
    } else
    for (;;) {
        // (re)load the tile around (i, j): rows ti .. ti + 31, byte columns tj .. tj + 127
        const int64_t wi = i & ~(int64_t)(WT_ROWS - 1), wj = j & ~(int64_t)(WT_COLS - 1);
        if (wi != ti || wj != tj) {
            ti = wi; tj = wj;
            __syncwarp();
            // 32 rows x 8 chunks of 16 bytes; lane l takes chunk l & 7 of rows l / 8, l / 8 + 4, ...
            for (int q = lane; q < WT_ROWS * (WT_COLS / 16); q += 32) {
                const int r = q >> 3, ch = q & 7;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (ti + r < n_rows && tj + ch * 16 < rs)
                    v = *(const uint4 *)(dir + (ti + r) * rs + tj + ch * 16);
                tiles[warp][q] = v;
            }
            __syncwarp();
        }
        int leave = 0;     // lane 0: 1 = walked out of the tile, 2 = trace finished
        if (lane == 0) {
            while (len < cap) {
                if (i < ti || j < tj || j >= tj + WT_COLS) { leave = 1; break; }
                const uint32_t d = tile[(i - ti) * WT_COLS + (j - tj)];
                int op, ns = 0;
                if (LEAN) {
                    if (i == 0 || j == 0) {
                        if (P.local || (i == 0 && j == 0)) { leave = 2; break; }
                        if (i == 0) { op = OP_TO1; ns = ST_GAP1; } else { op = OP_TO2; ns = ST_GAP2; }
                    } else if (st == ST_MATCH) {
                        const uint32_t tag = (d >> 2) & 7u;
                        if (tag == 0) { leave = 2; break; }
                        op = OP_DIAG; ns = tag == 7u ? ST_MATCH : (tag == 6u ? ST_GAP1 : ST_GAP2);
                    } else if (st == ST_GAP1) {
                        op = OP_TO1; ns = (d & 1u) ? ST_MATCH : ST_GAP1;
                    } else {
                        op = OP_TO2; ns = (d & 2u) ? ST_MATCH : ST_GAP2;
                    }
                } else if (AFFINE) {
                    if (st == ST_MATCH) {
                        if (d & A_MM) { op = OP_DIAG; ns = ST_MATCH; }
                        else if (d & A_G1M) { op = OP_DIAG; ns = ST_GAP1; }
                        else if (d & A_G2M) { op = OP_DIAG; ns = ST_GAP2; }
                        else { leave = 2; break; }
                    } else if (st == ST_GAP1) {
                        if (d & A_MG1) { op = OP_TO1; ns = ST_MATCH; }
                        else if (d & A_G1G1) { op = OP_TO1; ns = ST_GAP1; }
                        else { leave = 2; break; }
                    } else {
                        if (d & A_MG2) { op = OP_TO2; ns = ST_MATCH; }
                        else if (d & A_G2G2) { op = OP_TO2; ns = ST_GAP2; }
                        else { leave = 2; break; }
                    }
                } else {
                    if (d & L_MATCH) op = OP_DIAG;
                    else if (d & L_GAP1) op = OP_TO1;
                    else if (d & L_GAP2) op = OP_TO2;
                    else { leave = 2; break; }
                }
                if (P.stats) gs.push(op); else ops[len] = (uint8_t)op;
                len++;
                last_i = i; last_j = j;
                // table moves of a step: full table (-1,-1), (0,-1), (-1,0); band (-1,0), (0,-1), (-1,+1)
                if (op == OP_DIAG) { i--; if (!BANDED) j--; }
                else if (op == OP_TO1) { j--; }
                else { i--; if (BANDED) j++; }
                st = ns;
            }
            if (len >= cap && leave == 0) leave = 2;
        }
        leave = __shfl_sync(0xffffffffu, leave, 0);
        i = __shfl_sync(0xffffffffu, i, 0);
        j = __shfl_sync(0xffffffffu, j, 0);
        if (leave == 2 || i < 0 || j < 0) break;
    }
    if (lane == 0) {
        V.trace_len[p] = len;
        if (P.stats) {
            gs.finish();
            V.first[2 * p] = gs.opens; V.first[2 * p + 1] = gs.exts;
        } else {
            V.first[2 * p] = (int32_t)last_i;
            V.first[2 * p + 1] = (int32_t)last_j;
        }
    }
}

}  // namespace bt
