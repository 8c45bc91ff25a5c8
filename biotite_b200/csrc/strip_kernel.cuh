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
//   * scores are carried scaled by 32 with the state and the gap flags in the low five bits:
//         M'  = 32 M  + 0x1f            (state tag 7 in bits 2-4)
//         G1' = 32 G1 + 0x18 + f1       (tag 6; f1 = 1: opened from M, ties included)
//         G2' = 32 G2 + 0x10 + 2 f2     (tag 4; f2 likewise in bit 1)
//     so that every maximum says by itself which candidate won, with the reference's priorities
//     (cell.rs:219-254: Match from M before G1 before G2, a gap opened from M before it is
//     extended): H' = max3(M', G1', G2') is one VIMNMX3 whose bits 2-4 are the state the next
//     diagonal cell comes from, G1' = max(M' + 32 open - 6, (G1' & ~1) + 32 ext) carries f1, the
//     vertical state likewise.  No compare, no select: the direction byte of a cell is two LOP3 -
//     bits 2-4 of the diagonal H', bit 0 of G1', bit 1 of G2' (bits 5-7 are not defined) - a
//     layout of its own that wavefront_traceback_kernel<.., LEAN = true> decodes; the walk
//     (trace.rs:59-90 with max_number = 1) is the reference's.  Callers that enumerate co-optimal
//     traces (max_number > 1) keep the full-set kernel and the reference's byte layout;
//   * the vertical gap state of the row below is computed by the row above (Strip2), so two
//     values cross a lane / stripe boundary;
//   * local mode clamps M only: a gap state <= 0 can never win a maximum or be walked into when
//     gap penalties are <= 0, so leaving it unclamped changes no score and no visited byte.
//
// Exact while 32 |value| stays far from 2^29 (checked by the router: strip_eligible).
#pragma once
#include "common.cuh"
#include "general_kernel.cuh"
#include "wavefront_kernel.cuh"

namespace bt {

constexpr int SK_CHUNK = 16;                  // steps between progress updates / boundary prefetches
constexpr int32_t SK_NEG = -(1 << 29);        // "minus infinity" of the scaled scores

// What the row below needs of a cell (i, j): h = H'(i, j) (diagonal input of (i + 1, j + 1)) and
// t = G2'(i + 1, j) = max(M'(i, j) + 32 open - 13, (G2'(i, j) & ~2) + 32 ext), flag included: the
// vertical recurrence of the row below, done by the producer, so that two values cross a lane /
// stripe boundary instead of three.
struct Strip2 { int32_t h, t; };

// row 0 of the table (pairwise.rs:747-780) as seen from row 1
__device__ __forceinline__ Strip2 sk_row0(const DevParams &P, int j, int32_t o2c, int32_t e2c) {
    Strip2 c;
    if (j == 0) { c.h = 0x1f; c.t = 0x1f + o2c; return c; }        // M(0, 0) = 0: G2(1, 0) opens from it
    c.h = (P.local || !P.terminal) ? 0x18 : 32 * (P.gap_open + (j - 1) * P.gap_ext) + 0x18;
    c.t = SK_NEG + 0x1f + o2c;                                    // M(0, j) = G2(0, j) = -inf
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
    const uint32_t rs32 = (uint32_t)rs;           // m < 2^30 (checked by the router)
    uint8_t *const dir = TRACED ? V.dirs + V.dir_off[p] : nullptr;

    // shared memory: [matrix, scaled by 32] [boundary slots]
    int32_t *const mat8 = wf_smem;
    const int mat_n = P.n_rows * P.n_cols;
    for (int k = tid; k < mat_n; k += nthr) mat8[k] = V.matrix[k] * 32;
    volatile int32_t *const bnd = W_.bnd ? W_.bnd + (int64_t)blockIdx.x * W_.bnd_block : wf_smem + mat_n;
    const int bnd_cols = W_.bnd_cols;
    if (tid < WF_MAX_WARPS) prog[tid] = 0ull;
    // (table row 0 and column 0 need no direction bytes: the lean traceback walks the borders by rule)
    __syncthreads();

    const int32_t open8 = 32 * P.gap_open, ext8 = 32 * P.gap_ext;
    const int n_stripes = (n + ROWS - 1) / ROWS;
    const int lane_cols = (m | 3) + 1;            // table columns 0 .. (m | 3): whole direction words
    const int n_steps = 31 + lane_cols;
    const uint8_t *const codes1 = (const uint8_t *)V.codes1 + o1;
    const uint8_t *const codes2 = (const uint8_t *)V.codes2 + o2;
    int32_t best = 0x1f;                           // local: first maximum in row-major order, M'(0, 0) = 0x1f
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
            rbest[r] = 0x1f; rcol[r] = 0;
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
            const int32_t o2c = ((SEMI && j == m) ? 0 : open8) - 13, e2c = (SEMI && j == m) ? 0 : ext8;
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
                    // acquire: the producer's boundary entries are visible once its counter is
                    const uint32_t flag = (uint32_t)__cvta_generic_to_shared(&prog[(q - 1) % nw]);
                    for (;;) {
                        unsigned long long seen;
                        asm volatile("ld.acquire.cta.shared.u64 %0, [%1];" : "=l"(seen) : "r"(flag) : "memory");
                        if (seen >= want) break;
                        __nanosleep(100);
                    }
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
                    // vertical step: M' + 32 open - 13 (tag 7 -> 4, flag bit 1), G2' + 32 ext; waived in the last column
                    const int32_t o2c = ((SEMI && j == m) ? 0 : open8) - 13, e2c = (SEMI && j == m) ? 0 : ext8;
                    int32_t dh = diag_h, t = up_t;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        int32_t mn = (dh | 0x1f) + sim[r];
                        int32_t msrc = dh;                            // bits 2-4: where M comes from
                        if (LOCAL) {
                            if (mn <= 0x1f) { mn = 0x1f; msrc = 0; }  // M <= 0: trace ends here (pairwise.rs:661-666)
                        }
                        // horizontal step: M' + 32 open - 6 (tag 7 -> 6, flag bit 0), G1' + 32 ext; free in the last row
                        const bool free1 = SEMI && r == r_last;
                        const int32_t g1 = max(M[r] + (free1 ? -6 : open8 - 6), (G1[r] & ~1) + (free1 ? 0 : ext8));
                        const int32_t g2 = t & ~2;
                        const int32_t h = __vimax3_s32(mn, g1, g2);
                        if (TRACED) {
                            // bit 0: G1 opened, bit 1: G2 opened, bits 2-4: source of M; bits 5-7 undefined
                            const uint32_t byte = (((uint32_t)g1 & ~2u) | ((uint32_t)t & 2u)) & ~0x1cu | ((uint32_t)msrc & 0x1cu);
                            dw[r] = __byte_perm(dw[r], byte, 0x4321);
                        }
                        // G2' of the row below in this column
                        t = max(mn + o2c, g2 + e2c);
                        dh = H[r];
                        M[r] = mn; G1[r] = g1; H[r] = h;
                        if (LOCAL) { if (mn > rbest[r]) { rbest[r] = mn; rcol[r] = j; } }
                    }
                    last_t = t;
                } else if (j == 0) {
                    // table column 0 (pairwise.rs:707-745): leading gap in sequence 2
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int i = i0 + r;
                        M[r] = SK_NEG; G1[r] = SK_NEG;
                        H[r] = (LOCAL || !P.terminal) ? 0x10 : 32 * (P.gap_open + (i - 1) * P.gap_ext) + 0x10;
                        if (TRACED) dw[r] = __byte_perm(dw[r], 0u, 0x4321);
                    }
                    last_t = H[R - 1];              // never read: column 0 of the row below is closed-form too
                } else if (j > m && j < lane_cols) {
                    if (TRACED) {
#pragma unroll
                        for (int r = 0; r < R; r++) dw[r] = __byte_perm(dw[r], 0u, 0x4321);
                    }
                }
                if (TRACED && (j & 3) == 3 && (unsigned)j < (unsigned)lane_cols) {
                    // four columns of every row as one word; rows below the table have no storage.
                    // (Because of the lane skew a quarter of the lanes stores at every step: the
                    // address is one 32 x 32 + 64 bit multiply-add per row.)
                    uint8_t *const at = dir_lane + (j - 3);
                    if (rows_in) {
#pragma unroll
                        for (int r = 0; r < R; r++) *(uint32_t *)(at + (uint32_t)r * rs32) = dw[r];
                    } else {
#pragma unroll
                        for (int r = 0; r < R; r++)
                            if (r < rows_here) *(uint32_t *)(at + (uint32_t)r * rs32) = dw[r];
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
                    // release: the lane's boundary entries are visible before the counter says so
                    const unsigned long long done =
                        ((unsigned long long)q << 32) | (unsigned long long)(k1 >= n_steps ? 0xffffffffu : (unsigned)k1);
                    const uint32_t flag = (uint32_t)__cvta_generic_to_shared(&prog[q % nw]);
                    asm volatile("st.release.cta.shared.u64 [%0], %1;" ::"r"(flag), "l"(done) : "memory");
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
                const int tag = (h >> 2) & 7;
                V.score[p] = h >> 5;
                V.start_idx[p] = (int64_t)n * rs + m;
                V.start_state[p] = tag == 7 ? ST_MATCH : (tag == 6 ? ST_GAP1 : ST_GAP2);
                V.end_vals[3 * p] = M[r] >> 5; V.end_vals[3 * p + 1] = G1[r] >> 5; V.end_vals[3 * p + 2] = h >> 5;
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
            V.score[p] = bb.score >> 5;
            V.start_idx[p] = bb.idx;
            V.start_state[p] = ST_MATCH;
            V.end_vals[3 * p] = bb.score >> 5; V.end_vals[3 * p + 1] = P.neg_inf; V.end_vals[3 * p + 2] = P.neg_inf;
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
    return step * (max_span + 4) < (1ll << 22);
}

}  // namespace bt
