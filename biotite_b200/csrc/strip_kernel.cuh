// Lean intra-pair fill: the fast path of the single-pair API and of batches with few or long
// pairs (BASELINE config 1, the reference's own Cas9-sized benchmark pair, 10^2 - 10^4 pairs that
// cannot fill the inter-sequence kernels' groups of 64).
//
// Same decomposition as wavefront_fill_kernel (stripes of 128 rows, one warp per stripe in
// flight, lane t owns 4 rows and runs one column behind lane t - 1, the row above arrives by
// __shfl_up_sync, consecutive stripes hand their last row over through shared memory guarded by a
// progress counter) and the same outputs (score, start cell / state, one direction byte per cell
// in the reference's bit layout, cell.rs:44-80), so the warp-cooperative traceback and the host
// code around it are shared.  What differs is the cell:
//
//   * affine 3-state recurrence (pairwise.rs:608-681) on full tables, 8-bit codes, matrix scoring;
//     every other mode stays on wavefront_fill_kernel;
//   * scores are carried scaled by 8 with the state in the low bits - M' = 8 M + 7,
//     G1' = 8 G1 + 6, G2' = 8 G2 + 4 - so that H' = max3(M', G1', G2') (one VIMNMX3) is the
//     best state's score AND, in its low three bits, exactly the direction bits
//     A_MM | A_G1M | A_G2M (7), A_G1M | A_G2M (6) or A_G2M (4) the next diagonal cell stores:
//     ties resolve M > G1 > G2 as the reference's first-optimal walk does (cell.rs:219-254);
//   * the gap states are two adds and a maximum each; the "opened from M" bit is one compare, and
//     the vertical one is computed by the row above (Strip2), so two values cross a lane boundary;
//   * the byte holds the first-priority choice plus the lower-priority bits below it, which the
//     first-optimal walk never reaches: the trace of index 0 (trace.rs:59-90) is identical to the
//     one the full tie sets give, which is all a batch (max_number = 1) returns.  Callers that
//     enumerate co-optimal traces (max_number > 1) keep the full-set kernel;
//   * local mode clamps M only: a gap state <= 0 can never win a maximum or be walked into when
//     gap penalties are <= 0, so leaving it unclamped changes no score and no visited byte.
//
// Exact while 8 |value| stays far from 2^29 (checked by the router: strip_eligible).
#pragma once
#include "common.cuh"
#include "general_kernel.cuh"
#include "wavefront_kernel.cuh"

namespace bt {

constexpr int SK_CHUNK = 16;                  // steps between progress updates / boundary prefetches
constexpr int32_t SK_NEG = -(1 << 29);        // "minus infinity" of the scaled scores

// What the row below needs of a cell (i, j): h = H'(i, j) (diagonal input of (i + 1, j + 1)) and
// t = G2'(i + 1, j) = max(M'(i, j) + 8 open - 3, G2'(i, j) + 8 ext) with bit 0 set when the gap
// was opened from M (ties included): the vertical recurrence of the row below, done by the
// producer, so that two values cross a lane / stripe boundary instead of three.
struct Strip2 { int32_t h, t; };

// row 0 of the table (pairwise.rs:747-780) as seen from row 1
__device__ __forceinline__ Strip2 sk_row0(const DevParams &P, int j, int32_t o2c, int32_t e2c) {
    Strip2 c;
    if (j == 0) { c.h = 7; c.t = (7 + o2c) | 1; return c; }       // M(0, 0) = 0: G2(1, 0) opens from it
    c.h = (P.local || !P.terminal) ? 6 : 8 * (P.gap_open + (j - 1) * P.gap_ext) + 6;
    c.t = (SK_NEG + o2c) | 1;                                     // M(0, j) = G2(0, j) = -inf
    return c;
}

template <int MODE, bool TRACED, int R>
__global__ void __launch_bounds__(32 * WF_MAX_WARPS)
strip_fill_kernel(const DevParams P, const BatchView V, const WfArgs W_) {
    constexpr bool LOCAL = MODE == WF_LOCAL;
    constexpr bool SEMI = MODE == WF_SEMI;
    constexpr int ROWS = 32 * R, NS = 2;
    extern __shared__ int32_t wf_smem[];
    __shared__ unsigned long long prog[WF_MAX_WARPS];
    __shared__ int32_t red_score[WF_MAX_WARPS];
    __shared__ int64_t red_idx[WF_MAX_WARPS];

    const int64_t p = V.pair_begin + blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
    const int64_t o1 = V.map.beg1(p), o2 = V.map.beg2(p);
    const int n = (int)V.map.len1(p), m = (int)V.map.len2(p);
    const int64_t rs = table_stride(1, (int64_t)m + 1);
    uint8_t *const dir = TRACED ? V.dirs + V.dir_off[p] : nullptr;

    // shared memory: [matrix, scaled by 8] [boundary slots]
    int32_t *const mat8 = wf_smem;
    const int mat_n = P.n_rows * P.n_cols;
    for (int k = tid; k < mat_n; k += nthr) mat8[k] = V.matrix[k] * 8;
    volatile int32_t *const bnd = W_.bnd ? W_.bnd + (int64_t)blockIdx.x * W_.bnd_block : wf_smem + mat_n;
    const int bnd_cols = W_.bnd_cols;
    if (tid < WF_MAX_WARPS) prog[tid] = 0ull;
    if (TRACED) {
        // table row 0: leading gap in sequence 1 (pairwise.rs:747-780); local: no directions
        for (int j = tid; j < (int)rs; j += nthr)
            dir[j] = (uint8_t)((j == 0 || j > m || LOCAL) ? 0 : (j == 1 ? A_MG1 : A_G1G1));
    }
    __syncthreads();

    const int32_t open8 = 8 * P.gap_open, ext8 = 8 * P.gap_ext;
    const int n_stripes = (n + ROWS - 1) / ROWS;
    const int lane_cols = (m | 3) + 1;            // table columns 0 .. (m | 3): whole direction words
    const int n_steps = 31 + lane_cols;
    const uint8_t *const codes1 = (const uint8_t *)V.codes1 + o1;
    const uint8_t *const codes2 = (const uint8_t *)V.codes2 + o2;
    int32_t best = 7;                              // local: first maximum in row-major order, M'(0, 0) = 7
    int best_i = 0, best_j = 0;

    for (int q = warp; q < n_stripes; q += nw) {
        const int base = q * ROWS;
        const int i0 = base + lane * R + 1;        // first table row of this lane
        volatile int32_t *const bnd_out = bnd + (int64_t)(q % nw) * bnd_cols * NS;
        volatile int32_t *const bnd_in = bnd + (int64_t)((q + nw - 1) % nw) * bnd_cols * NS;
        int32_t M[R], G1[R], H[R], simn[R];
        const int32_t *rowmat[R];
        uint32_t dw[R];
        int32_t rbest[R];
        int rcol[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int i = i0 + r;
            rowmat[r] = mat8 + (i <= n ? (int)codes1[i - 1] : 0) * P.n_cols;
            M[r] = G1[r] = H[r] = SK_NEG;
            simn[r] = 0; dw[r] = 0;
            rbest[r] = 7; rcol[r] = 0;
        }
        // the last row of the table inside this lane (semi-global: its horizontal steps are free)
        const int r_last = (SEMI && n >= i0 && n < i0 + R) ? n - i0 : -1;
        int32_t last_t = SK_NEG;                   // Strip2::t of the lane's last row
        int32_t diag_h = SK_NEG;                   // H' of the cell above the lane's first row, one column back
        uint8_t *const dir_lane = TRACED ? dir + (int64_t)i0 * rs : nullptr;
        const int rows_here = min(R, n - i0 + 1);  // table rows among the lane's R rows (<= 0: none)
        const bool rows_in = rows_here == R;
        auto fetch_sims = [&](int k) {
            const int j = k - lane;
            if (j >= 1 && j <= m) {
                const int c2 = (int)codes2[j - 1];
#pragma unroll
                for (int r = 0; r < R; r++) simn[r] = rowmat[r][c2];
            }
        };
        fetch_sims(0);
        auto above = [&](int j) -> Strip2 {        // cell (base, j): the row above the stripe
            if (j < 0 || j > m) return Strip2{SK_NEG, SK_NEG};
            const int32_t o2c = ((SEMI && j == m) ? 0 : open8) - 3, e2c = (SEMI && j == m) ? 0 : ext8;
            if (base == 0) return sk_row0(P, j, o2c, e2c);
            Strip2 c;
            c.h = bnd_in[(int64_t)j * NS]; c.t = bnd_in[(int64_t)j * NS + 1];
            return c;
        };
        const unsigned long long want_tag = (unsigned long long)(q > 0 ? q - 1 : 0) << 32;

        for (int k0 = 0; k0 < n_steps; k0 += SK_CHUNK) {
            if (q > 0) {
                if (lane == 0) {
                    const unsigned long long want = want_tag | (unsigned long long)(k0 + SK_CHUNK - 1 + 32);
                    const volatile unsigned long long *flag = &prog[(q - 1) % nw];
                    while (*flag < want) __nanosleep(100);
                    __threadfence_block();
                }
                __syncwarp();
            }
            const int k1 = min(k0 + SK_CHUNK, n_steps);
            Strip2 pre = Strip2{SK_NEG, SK_NEG};
            if (lane < SK_CHUNK) pre = above(k0 + lane);
            for (int k = k0; k < k1; k++) {
                const int j = k - lane;               // table column of this lane at this step
                // the cell above the lane's first row in this column: last row of the lane above
                int32_t up_h = __shfl_up_sync(0xffffffffu, H[R - 1], 1);
                int32_t up_t = __shfl_up_sync(0xffffffffu, last_t, 1);
                const int32_t ph = __shfl_sync(0xffffffffu, pre.h, k - k0), pt = __shfl_sync(0xffffffffu, pre.t, k - k0);
                if (lane == 0) { up_h = ph; up_t = pt; }
                int32_t sim[R];
#pragma unroll
                for (int r = 0; r < R; r++) sim[r] = simn[r];
                if (k + 1 < n_steps) fetch_sims(k + 1);
                if (j >= 1 && j <= m) {
                    // vertical step: M' + 8 open - 3 (state bits 7 -> 4), G2' + 8 ext; waived in the last column
                    const int32_t o2c = ((SEMI && j == m) ? 0 : open8) - 3, e2c = (SEMI && j == m) ? 0 : ext8;
                    int32_t dh = diag_h, t = up_t;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        int32_t mn = (dh | 7) + sim[r];
                        uint32_t byte = (uint32_t)dh & 7u;
                        if (LOCAL) {
                            if (mn <= 7) { mn = 7; byte = 0; }       // M <= 0: trace ends here (pairwise.rs:661-666)
                        }
                        // horizontal step: M' + 8 open - 1 (state bits 7 -> 6), G1' + 8 ext; free in the last row
                        const bool free1 = SEMI && r == r_last;
                        const int32_t a1 = M[r] + (free1 ? -1 : open8 - 1), b1 = G1[r] + (free1 ? 0 : ext8);
                        const int32_t g1 = max(a1, b1);
                        byte |= a1 >= b1 ? (uint32_t)(A_MG1 | A_G1G1) : (uint32_t)A_G1G1;
                        byte |= (t & 1) ? (uint32_t)(A_MG2 | A_G2G2) : (uint32_t)A_G2G2;
                        const int32_t g2 = t & ~1;
                        const int32_t h = __vimax3_s32(mn, g1, g2);
                        // G2' of the row below in this column
                        const int32_t a2 = mn + o2c, b2 = g2 + e2c;
                        t = max(a2, b2) | (a2 >= b2 ? 1 : 0);
                        dh = H[r];
                        M[r] = mn; G1[r] = g1; H[r] = h;
                        if (LOCAL) { if (mn > rbest[r]) { rbest[r] = mn; rcol[r] = j; } }
                        if (TRACED) dw[r] = __byte_perm(dw[r], byte, 0x4321);
                    }
                    last_t = t;
                } else if (j == 0) {
                    // table column 0 (pairwise.rs:707-745): leading gap in sequence 2
                    const int32_t e2c = (SEMI && m == 0) ? 0 : ext8;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int i = i0 + r;
                        M[r] = SK_NEG; G1[r] = SK_NEG;
                        H[r] = (LOCAL || !P.terminal) ? 4 : 8 * (P.gap_open + (i - 1) * P.gap_ext) + 4;
                        const uint32_t byte = LOCAL ? 0u : (i == 1 ? (uint32_t)A_MG2 : (uint32_t)A_G2G2);
                        if (TRACED) dw[r] = __byte_perm(dw[r], byte, 0x4321);
                    }
                    last_t = H[R - 1] + e2c;        // G2(i + 1, 0) extends G2(i, 0): never read (column 0 is closed-form)
                } else if (j > m && j < lane_cols) {
                    if (TRACED) {
#pragma unroll
                        for (int r = 0; r < R; r++) dw[r] = __byte_perm(dw[r], 0u, 0x4321);
                    }
                }
                if (TRACED && (j & 3) == 3 && (unsigned)j < (unsigned)lane_cols) {
                    // four columns of every row as one word; rows below the table have no storage
                    uint8_t *const at = dir_lane + (j - 3);
                    if (rows_in) {
#pragma unroll
                        for (int r = 0; r < R; r++) *(uint32_t *)(at + (int64_t)r * rs) = dw[r];
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++)
                            if (r < rows_here) *(uint32_t *)(at + (int64_t)r * rs) = dw[r];
                    }
                }
                diag_h = up_h;
                // the stripe's last row goes to the next stripe
                if (lane == 31 && j >= 0 && j <= m && j < bnd_cols) {
                    bnd_out[(int64_t)j * NS] = H[R - 1]; bnd_out[(int64_t)j * NS + 1] = last_t;
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
        if (!LOCAL && m >= 1) {
            // optimum and traceback start of the global modes (pairwise.rs:845-881): the end cell's
            // states are still in the registers of the lane that owns row n
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (i0 + r != n) continue;
                const int32_t h = H[r];
                V.score[p] = h >> 3;
                V.start_idx[p] = (int64_t)n * rs + m;
                V.start_state[p] = (h & 7) == 7 ? ST_MATCH : ((h & 7) == 6 ? ST_GAP1 : ST_GAP2);
                V.end_vals[3 * p] = M[r] >> 3; V.end_vals[3 * p + 1] = G1[r] >> 3; V.end_vals[3 * p + 2] = h >> 3;
            }
        }
        if (LOCAL) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = i0 + r;
                if (i <= n && rbest[r] > best) { best = rbest[r]; best_i = i; best_j = rcol[r]; }
            }
        }
    }
    __syncthreads();

    if (LOCAL) {
        Best bb;
        bb.score = best; bb.idx = (int64_t)best_i * rs + best_j;
        for (int off = 16; off > 0; off >>= 1) {
            const int32_t os = __shfl_down_sync(0xffffffffu, bb.score, off);
            const int64_t oi = __shfl_down_sync(0xffffffffu, bb.idx, off);
            bb.offer(os, oi);
        }
        if (lane == 0) { red_score[warp] = bb.score; red_idx[warp] = bb.idx; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < nw; k++) bb.offer(red_score[k], red_idx[k]);
            V.score[p] = bb.score >> 3;
            V.start_idx[p] = bb.idx;
            V.start_state[p] = ST_MATCH;
            V.end_vals[3 * p] = bb.score >> 3; V.end_vals[3 * p + 1] = P.neg_inf; V.end_vals[3 * p + 2] = P.neg_inf;
        }
    } else if (tid == 0 && (n == 0 || m == 0)) {
        // the end cell lies on the table's border (closed forms of wavefront_kernel.cuh)
        uint32_t dd;
        const Cell3 c = n == 0 ? wf_row0<true>(P, m, dd) : wf_col0<true>(P, n, dd);
        const int32_t mx = max(c.m, max(c.g1, c.g2));
        V.score[p] = mx;
        V.start_idx[p] = (int64_t)n * rs + m;
        V.start_state[p] = (c.m == mx) ? ST_MATCH : ((c.g1 == mx) ? ST_GAP1 : ST_GAP2);
        V.end_vals[3 * p] = c.m; V.end_vals[3 * p + 1] = c.g1; V.end_vals[3 * p + 2] = c.g2;
    }
}

// Modes the lean kernel serves, and the range in which its scaled arithmetic is exact.
inline bool strip_eligible(const BtParams &P, int64_t max_span, int32_t min_score, int32_t max_score) {
    if (!P.affine || P.banded || !P.use_matrix || P.code_bytes != 1) return false;
    if ((int64_t)P.n_rows * P.n_cols > 4096) return false;
    const int64_t step = std::max<int64_t>({std::abs((int64_t)min_score), std::abs((int64_t)max_score),
                                            std::abs((int64_t)P.gap_open), std::abs((int64_t)P.gap_ext)});
    return step * (max_span + 4) < (1ll << 24);
}

}  // namespace bt
