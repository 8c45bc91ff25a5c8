// General int32 wavefront kernel: one thread block per pair, every mode of the
// reference (linear/affine, global/semi-global/local, matrix or match/mismatch,
// full table or band, any code width), full direction sets (the reference's
// 3-/7-bit flags, one byte per cell) so that co-optimal enumeration is possible.
//
// It computes exactly what the reference's fill loops compute
//   full table: src/rust/sequence/align/pairwise.rs:188-227 (+ :378-415, :608-681,
//               first row/column :430-467, :695-780)
//   band      : src/rust/sequence/align/banded.rs:184-214 (+ :348-375, :486-548)
// but sweeps anti-diagonals (full: d = i + j; straightened band: w = 2i + c)
// with three rolling wavefront buffers instead of a row-major table of cells.
// Arithmetic is wrapping int32 with the reference's neg_inf sentinel, so results
// are bit-identical by construction.  This is the reference-complete path; the
// throughput path for short pairs is interseq_kernel.cuh.
#pragma once
#include "common.cuh"

namespace bt {

struct Cell3 {
    int32_t m, g1, g2;
};

// cell.rs:291-323 - the set of arg-maxima of three candidates
__device__ __forceinline__ int32_t max3_set(int32_t a, int32_t b, int32_t c, uint32_t &set) {
    int32_t mx = max(a, max(b, c));
    set = (a == mx ? 1u : 0u) | (b == mx ? 2u : 0u) | (c == mx ? 4u : 0u);
    return mx;
}

template <bool AFFINE>
__device__ __forceinline__ Cell3 cell_update(const DevParams &P, int32_t sim, const Cell3 &fm,
                                             const Cell3 &f1, const Cell3 &f2, bool waive1,
                                             bool waive2, uint32_t &dir) {
    Cell3 out;
    if (AFFINE) {
        // pairwise.rs:608-681 / banded.rs:486-548
        int32_t mm = wadd(fm.m, sim), g1m = wadd(fm.g1, sim), g2m = wadd(fm.g2, sim);
        int32_t mg1 = waive1 ? f1.m : wadd(f1.m, P.gap_open);
        int32_t g1g1 = waive1 ? f1.g1 : wadd(f1.g1, P.gap_ext);
        int32_t mg2 = waive2 ? f2.m : wadd(f2.m, P.gap_open);
        int32_t g2g2 = waive2 ? f2.g2 : wadd(f2.g2, P.gap_ext);
        uint32_t d;
        out.m = max3_set(mm, g1m, g2m, d);
        out.g1 = max(mg1, g1g1);
        d |= (mg1 >= g1g1 ? (uint32_t)A_MG1 : 0u) | (g1g1 >= mg1 ? (uint32_t)A_G1G1 : 0u);
        out.g2 = max(mg2, g2g2);
        d |= (mg2 >= g2g2 ? (uint32_t)A_MG2 : 0u) | (g2g2 >= mg2 ? (uint32_t)A_G2G2 : 0u);
        if (P.local) {
            if (out.m <= 0) { d &= ~(uint32_t)(A_MM | A_G1M | A_G2M); out.m = 0; }
            if (out.g1 <= 0) { d &= ~(uint32_t)(A_MG1 | A_G1G1); out.g1 = P.neg_inf; }
            if (out.g2 <= 0) { d &= ~(uint32_t)(A_MG2 | A_G2G2); out.g2 = P.neg_inf; }
        }
        dir = d;
    } else {
        // pairwise.rs:378-415 / banded.rs:348-375
        int32_t a = wadd(fm.m, sim);
        int32_t b = waive1 ? f1.m : wadd(f1.m, P.gap_open);
        int32_t c = waive2 ? f2.m : wadd(f2.m, P.gap_open);
        uint32_t d;
        out.m = max3_set(a, b, c, d);
        out.g1 = out.g2 = 0;
        if (P.local && out.m <= 0) { out.m = 0; d = 0; }
        dir = d;
    }
    return out;
}

template <bool AFFINE>
__device__ __forceinline__ Cell3 load_cell(const int32_t *buf, int idx) {
    Cell3 c;
    if (AFFINE) {
        c.m = buf[3 * idx]; c.g1 = buf[3 * idx + 1]; c.g2 = buf[3 * idx + 2];
    } else {
        c.m = buf[idx]; c.g1 = c.g2 = 0;
    }
    return c;
}
template <bool AFFINE>
__device__ __forceinline__ void store_cell(int32_t *buf, int idx, const Cell3 &c) {
    if (AFFINE) {
        buf[3 * idx] = c.m; buf[3 * idx + 1] = c.g1; buf[3 * idx + 2] = c.g2;
    } else {
        buf[idx] = c.m;
    }
}

__device__ __forceinline__ int32_t similarity(const DevParams &P, const int32_t *mat,
                                              uint64_t s1, uint64_t s2) {
    if (P.use_matrix) return mat[s1 * (uint64_t)P.n_cols + s2];  // scoring.rs:62-64
    return s1 == s2 ? P.match : P.mismatch;                       // scoring.rs:101-108
}

// Keeps the first (smallest table index) maximum, as the row-major scan of
// pairwise.rs:536-551 / :849-865 does.
struct Best {
    int32_t score;
    int64_t idx;
    __device__ __forceinline__ void offer(int32_t s, int64_t i) {
        if (s > score || (s == score && i < idx)) { score = s; idx = i; }
    }
};

template <bool AFFINE, bool BANDED>
__global__ void __launch_bounds__(1024)
general_fill_kernel(const DevParams P, const BatchView V) {
    extern __shared__ int32_t smem[];
    __shared__ int32_t red_score[32];
    __shared__ int64_t red_idx[32];

    const int64_t p = V.pair_begin + blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int64_t o1 = V.map.beg1(p), o2 = V.map.beg2(p);
    const int n = (int)V.map.len1(p), m = (int)V.map.len2(p);
    constexpr int NS = AFFINE ? 3 : 1;
    const int32_t NI = P.neg_inf;

    int lower = 0, upper = 0, W = 0;
    if (BANDED) {
        lower = (int)V.bands[2 * p]; upper = (int)V.bands[2 * p + 1];
        W = upper - lower + 1;
    }
    const int64_t rs = BANDED ? (int64_t)W + 2 : (int64_t)m + 1;  // table row stride
    const int slots = BANDED ? W + 2 : n + 1;                     // entries per wavefront

    int32_t *buf;
    if (V.gscratch) buf = V.gscratch + (int64_t)blockIdx.x * V.gscratch_stride;
    else buf = smem;
    const int32_t *mat = V.matrix;
    if (P.matrix_in_smem) {
        int32_t *smat = smem + (V.gscratch ? 0 : 3 * slots * NS);
        for (int k = tid; k < P.n_rows * P.n_cols; k += nthr) smat[k] = V.matrix[k];
        mat = smat;
        __syncthreads();  // the first wavefront reads entries staged by other threads
    }
    uint8_t *dir = P.score_only ? nullptr : V.dirs + V.dir_off[p];
    int32_t *cand = (BANDED && !P.local) ? V.cand + V.cand_off[p] : nullptr;

    Best best;
    best.score = INT32_MIN; best.idx = INT64_MAX;
    const Cell3 EDGE = {0, AFFINE ? NI : 0, AFFINE ? NI : 0};  // banded.rs:334-337, :462-472
    const Cell3 SENT = {NI, AFFINE ? NI : 0, AFFINE ? NI : 0}; // banded.rs:339-345, :474-484
    const bool semi = !P.terminal && !P.local;

    if (!BANDED) {
        // origin (pairwise.rs:430-432, :695-705)
        if (tid == 0) {
            Cell3 o = {0, NI, NI};
            store_cell<AFFINE>(buf, 0, o);
            if (dir) dir[0] = (P.mark && P.mark_value == 0) ? DIR_MARK : 0;
            if (P.local) best.offer(0, 0);
        }
        __syncthreads();
        for (int d = 1; d <= n + m; d++) {
            int32_t *cur = buf + (d % 3) * slots * NS;
            const int32_t *p1 = buf + ((d + 2) % 3) * slots * NS;
            const int32_t *p2 = buf + ((d + 1) % 3) * slots * NS;
            if (tid == 0) {
                if (d <= m) {  // first row cell (0, d): fill_start2
                    Cell3 pr = load_cell<AFFINE>(p1, 0), c;
                    uint32_t dd = 0;
                    if (AFFINE) {
                        if (P.local) { c.m = NI; c.g1 = 0; c.g2 = NI; }
                        else {
                            int32_t open = P.terminal ? wadd(pr.m, P.gap_open) : pr.m;
                            int32_t ext = P.terminal ? wadd(pr.g1, P.gap_ext) : pr.g1;
                            c.m = NI; c.g2 = NI;
                            if (open >= ext) { c.g1 = open; dd = A_MG1; }
                            else { c.g1 = ext; dd = A_G1G1; }
                        }
                    } else {
                        if (P.local) c.m = 0;
                        else { c.m = P.terminal ? wadd(pr.m, P.gap_open) : 0; dd = L_GAP1; }
                        c.g1 = c.g2 = 0;
                    }
                    store_cell<AFFINE>(cur, 0, c);
                    if (P.local) best.offer(c.m, d);
                    if (dir) dir[d] = (uint8_t)(dd | ((P.mark && c.m == P.mark_value) ? DIR_MARK : 0));
                }
                if (d <= n) {  // first column cell (d, 0): fill_start1
                    Cell3 pr = load_cell<AFFINE>(p1, d - 1), c;
                    uint32_t dd = 0;
                    if (AFFINE) {
                        if (P.local) { c.m = NI; c.g1 = NI; c.g2 = 0; }
                        else {
                            int32_t open = P.terminal ? wadd(pr.m, P.gap_open) : pr.m;
                            int32_t ext = P.terminal ? wadd(pr.g2, P.gap_ext) : pr.g2;
                            c.m = NI; c.g1 = NI;
                            if (open >= ext) { c.g2 = open; dd = A_MG2; }
                            else { c.g2 = ext; dd = A_G2G2; }
                        }
                    } else {
                        if (P.local) c.m = 0;
                        else { c.m = P.terminal ? wadd(pr.m, P.gap_open) : 0; dd = L_GAP2; }
                        c.g1 = c.g2 = 0;
                    }
                    store_cell<AFFINE>(cur, d, c);
                    if (P.local) best.offer(c.m, (int64_t)d * rs);
                    if (dir) dir[(int64_t)d * rs] = (uint8_t)(dd | ((P.mark && c.m == P.mark_value) ? DIR_MARK : 0));
                }
            }
            const int ilo = max(1, d - m), ihi = min(n, d - 1);
            for (int i = ilo + tid; i <= ihi; i += nthr) {
                const int j = d - i;
                const uint64_t s1 = load_code(V.codes1, o1 + i - 1, P.code_bytes);
                const uint64_t s2 = load_code(V.codes2, o2 + j - 1, P.code_bytes);
                const int32_t sim = similarity(P, mat, s1, s2);
                Cell3 fm = load_cell<AFFINE>(p2, i - 1);
                Cell3 f1 = load_cell<AFFINE>(p1, i);
                Cell3 f2 = load_cell<AFFINE>(p1, i - 1);
                uint32_t dd;
                Cell3 c = cell_update<AFFINE>(P, sim, fm, f1, f2, semi && i == n, semi && j == m, dd);
                store_cell<AFFINE>(cur, i, c);
                const int64_t idx = (int64_t)i * rs + j;
                if (P.local) best.offer(c.m, idx);
                if (dir) dir[idx] = (uint8_t)(dd | ((P.mark && c.m == P.mark_value) ? DIR_MARK : 0));
            }
            __syncthreads();
        }
    } else {
        if (P.local && tid == 0) best.offer(0, 1);  // first edge cell (0,1) of the row-major scan
        if (cand) {
            // end candidates that are never filled (n == 0) keep the edge value
            for (int k = tid; k < W + 2; k += nthr) {
                cand[k] = 0;
                if (AFFINE) { cand[(W + 2) + k] = NI; cand[2 * (W + 2) + k] = NI; }
            }
            __syncthreads();
        }
        const int wmax = 2 * n + W;
        for (int w = 3; w <= wmax; w++) {
            int32_t *cur = buf + (w % 3) * slots * NS;
            const int32_t *p1 = buf + ((w + 2) % 3) * slots * NS;
            const int32_t *p2 = buf + ((w + 1) % 3) * slots * NS;
            const int ilo = max(1, (w - W + 1) >> 1), ihi = min(n, (w - 1) >> 1);
            for (int i = ilo + tid; i <= ihi; i += nthr) {
                const int c = w - 2 * i;
                const int sj = c + i - 2 + lower;  // banded.rs:248-250
                if (sj < 0 || sj >= m) continue;
                const uint64_t s1 = load_code(V.codes1, o1 + i - 1, P.code_bytes);
                const uint64_t s2 = load_code(V.codes2, o2 + sj, P.code_bytes);
                const int32_t sim = similarity(P, mat, s1, s2);
                // diag (i-1, c), left (i, c-1), up (i-1, c+1)  (banded.rs:194-196)
                Cell3 fm = (i >= 2 && sj >= 1) ? load_cell<AFFINE>(p2, c) : EDGE;
                Cell3 f1 = (c == 1) ? SENT : (sj >= 1 ? load_cell<AFFINE>(p1, c - 1) : EDGE);
                Cell3 f2 = (c == W) ? SENT : (i >= 2 ? load_cell<AFFINE>(p1, c + 1) : EDGE);
                uint32_t dd;
                Cell3 cc = cell_update<AFFINE>(P, sim, fm, f1, f2, false, false, dd);
                store_cell<AFFINE>(cur, c, cc);
                const int64_t idx = (int64_t)i * rs + c;
                if (P.local) best.offer(cc.m, idx);
                if (dir) dir[idx] = (uint8_t)(dd | ((P.mark && cc.m == P.mark_value) ? DIR_MARK : 0));
                if (cand && (i == n || sj == m - 1)) {
                    cand[c] = cc.m;
                    if (AFFINE) { cand[(W + 2) + c] = cc.g1; cand[2 * (W + 2) + c] = cc.g2; }
                }
            }
            __syncthreads();
        }
    }

    // ---- optimum and traceback start (pairwise.rs:532-559, :845-881; banded.rs:377-414, :550-593)
    if (P.local) {
        for (int off = 16; off > 0; off >>= 1) {
            int32_t os = __shfl_down_sync(0xffffffffu, best.score, off);
            int64_t oi = __shfl_down_sync(0xffffffffu, best.idx, off);
            best.offer(os, oi);
        }
        if ((tid & 31) == 0) { red_score[tid >> 5] = best.score; red_idx[tid >> 5] = best.idx; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < (nthr + 31) / 32; k++) best.offer(red_score[k], red_idx[k]);
            V.score[p] = best.score;
            V.start_idx[p] = best.idx;
            V.start_state[p] = ST_MATCH;
            V.end_vals[3 * p] = best.score; V.end_vals[3 * p + 1] = NI; V.end_vals[3 * p + 2] = NI;
        }
    } else if (!BANDED) {
        if (tid == 0) {
            Cell3 c = load_cell<AFFINE>(buf + ((n + m) % 3) * slots * NS, n);
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

// First-optimal traceback over full direction sets: one thread per pair follows
// the highest-priority step of every cell (cell.rs:125-138, :219-254; trace.rs:59-90
// with max_number = 1) and emits one step byte per path cell, end to start.
template <bool AFFINE, bool BANDED>
__global__ void general_traceback_kernel(const DevParams P, const BatchView V) {
    const int64_t p = V.pair_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= V.pair_end) return;
    const int64_t n = V.map.len1(p), m = V.map.len2(p);
    int64_t rs = table_stride(P.dir_pad, m + 1);
    if (BANDED) rs = table_stride(P.dir_pad, V.bands[2 * p + 1] - V.bands[2 * p] + 3);
    const int64_t step[3] = {BANDED ? -rs : -(rs + 1), -1, BANDED ? -rs + 1 : -rs};
    const uint8_t *dir = V.dirs + V.dir_off[p];
    uint8_t *ops = P.stats ? nullptr : V.ops + V.map.ops(p);
    GapStats gs(P.terminal != 0);
    const int64_t cap = n + m;
    int64_t idx = V.start_idx[p], last = 0, len = 0;
    int st = V.start_state[p];
    while (len < cap) {
        const uint32_t d = dir[idx];
        int op, ns = 0;
        if (AFFINE) {
            if (st == ST_MATCH) {
                if (d & A_MM) { op = OP_DIAG; ns = ST_MATCH; }
                else if (d & A_G1M) { op = OP_DIAG; ns = ST_GAP1; }
                else if (d & A_G2M) { op = OP_DIAG; ns = ST_GAP2; }
                else break;
            } else if (st == ST_GAP1) {
                if (d & A_MG1) { op = OP_TO1; ns = ST_MATCH; }
                else if (d & A_G1G1) { op = OP_TO1; ns = ST_GAP1; }
                else break;
            } else {
                if (d & A_MG2) { op = OP_TO2; ns = ST_MATCH; }
                else if (d & A_G2G2) { op = OP_TO2; ns = ST_GAP2; }
                else break;
            }
        } else {
            if (d & L_MATCH) op = OP_DIAG;
            else if (d & L_GAP1) op = OP_TO1;
            else if (d & L_GAP2) op = OP_TO2;
            else break;
        }
        if (P.stats) gs.push(op); else ops[len] = (uint8_t)op;
        len++;
        last = idx;
        idx += step[op];
        st = ns;
    }
    V.trace_len[p] = len;
    if (P.stats) {
        gs.finish();
        V.first[2 * p] = gs.opens; V.first[2 * p + 1] = gs.exts;
    } else {
        V.first[2 * p] = (int32_t)(last / rs);
        V.first[2 * p + 1] = (int32_t)(last % rs);
    }
}

}  // namespace bt
