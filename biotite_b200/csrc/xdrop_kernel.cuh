// Local X-drop extension of a region: the native part of the
// reference's align_local_gapped, `align_region` in src/rust/sequence/align/localgapped.rs.
//
// Reference algorithm (localgapped.rs:182-310): the table is filled antidiagonal by
// antidiagonal; a score of 0 marks an unreached cell, the anchor (0,0) holds threshold + 1,
// the diagonal predecessor only counts if it is itself reached, and every candidate is pruned
// against the running maximum (`score >= max - threshold`), which is updated cell by cell in
// order of increasing row.  The row range of an antidiagonal is the hull of the ranges that
// survived on the two previous ones (:209-223), the fill ends when that hull is empty.
//
// Mapping.  One thread block per region, one thread per cell of the antidiagonal (chunks of
// the block size).  The candidates of an antidiagonal only depend on the two previous ones, so
// they are computed in parallel; the reference's sequential running maximum is reproduced
// exactly by an exclusive prefix maximum over the candidates, because a rejected candidate is
// by definition below the running maximum and cannot change it.  Scores live in three rolling
// antidiagonal buffers indexed by row; the direction bytes (the reference's 3- / 7-bit tie
// sets, cell.rs:44-80) go to a rows x cols table sized like the reference's doubling table
// (:232-256), cells that hold the final maximum are tagged DIR_MARK for the start search
// (:491-511, :621-641).  Two launches: the first (score only) yields the maximum, the table
// shape and the MemoryError condition, the second writes the directions.
#pragma once
#include "common.cuh"

namespace bt {

constexpr int XD_THREADS = 128;
constexpr int XD_INIT_SIZE = 100;  // localgapped.rs:45

struct XdropArgs {
    const void *codes1, *codes2;       // concatenated region codes
    const int64_t *off1, *off2;        // n_regions + 1
    int code_bytes;
    const int32_t *matrix;
    int n_cols, use_matrix;
    int32_t match, mismatch, gap_open, gap_ext, threshold;
    int64_t max_table_size;
    int32_t *roll;                     // rolling antidiagonal buffers
    const int64_t *roll_off;           // first int of region r (3 diagonals x states x (n + 1))
    int32_t *max_score;                // out (pass 1) / in (pass 2): raw maximum, anchor included
    int64_t *dims;                     // out (pass 1): rows, cols of the reference's table
    int32_t *status;                   // out (pass 1): 1 = maximum table size exceeded
    uint8_t *dirs;                     // pass 2: direction tables
    const int64_t *dir_off;            // first byte of region r (rows x cols, row stride = cols)
};

// cell.rs:291-323: maximum of three with the set of all arguments attaining it
__device__ __forceinline__ int32_t xd_max3(int32_t a, int32_t b, int32_t c, uint32_t &dir) {
    const int32_t mx = max(a, max(b, c));
    dir = (a == mx ? 1u : 0u) | (b == mx ? 2u : 0u) | (c == mx ? 4u : 0u);
    return mx;
}

template <bool AFFINE, bool TRACED>
__global__ void __launch_bounds__(XD_THREADS) xdrop_fill_kernel(const XdropArgs A) {
    const int64_t r = blockIdx.x;
    if (TRACED && A.status[r]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t b1 = A.off1[r], b2 = A.off2[r];
    const int64_t n = A.off1[r + 1] - b1, m = A.off2[r + 1] - b2;
    constexpr int STATES = AFFINE ? 3 : 1;
    int32_t *roll = A.roll + A.roll_off[r];
    const int64_t plane = n + 1;                       // ints per state and diagonal
    auto buf = [&](int64_t k, int state) -> int32_t * { return roll + ((k % 3) * STATES + state) * plane; };
    const int32_t init = wadd(A.threshold, 1);
    const int32_t final_max = TRACED ? A.max_score[r] : 0;
    const int64_t cols_final = TRACED ? A.dims[2 * r + 1] : 0;
    uint8_t *dirs = TRACED ? A.dirs + A.dir_off[r] : nullptr;

    __shared__ int64_t s_imin, s_imax;                 // row range of the current antidiagonal
    __shared__ int64_t s_emin1, s_emax1, s_emin2, s_emax2;   // evaluated ranges of k-1, k-2
    __shared__ long long s_acc_min, s_acc_max;         // surviving range of the current antidiagonal
    __shared__ int32_t s_max, s_warp_max[XD_THREADS / 32];
    __shared__ int s_stop;
    // thread 0 only: surviving ranges of k, k-1, k-2 and the emulated table shape
    int64_t i_min_k0 = 0, i_max_k0 = 0, i_min_k1 = 0, i_max_k1 = 0;
    int64_t rows = min(n + 1, (int64_t)XD_INIT_SIZE), cols = min(m + 1, (int64_t)XD_INIT_SIZE);
    if (tid == 0) {
        buf(0, 0)[0] = init;                           // anchor (localgapped.rs:194-195)
        if (AFFINE) { buf(0, 1)[0] = 0; buf(0, 2)[0] = 0; }
        if (TRACED && init == final_max) dirs[0] = DIR_MARK;
        s_max = init;
        s_emin1 = 0; s_emax1 = 0; s_emin2 = 1; s_emax2 = 0;   // k-1 = the anchor, k-2 = nothing
        s_stop = 0;
    }
    __syncthreads();

    for (int64_t k = 1; k <= n + m; k++) {
        if (tid == 0) {
            const int64_t i_min_k2 = i_min_k1, i_max_k2 = i_max_k1;
            i_min_k1 = i_min_k0; i_max_k1 = i_max_k0;
            int64_t i_min = max(min(i_min_k1, i_min_k2 + 1), k > m ? k - m : (int64_t)0);
            int64_t i_max = min(max(i_max_k1 + 1, i_max_k2 + 1), n);
            int stop = i_min > i_max;
            if (!stop && !TRACED) {
                // growth of the reference's table (localgapped.rs:232-256)
                const int64_t j_max = k - i_min;
                if (i_max >= rows) {
                    if (rows * 2 > A.max_table_size / cols) { A.status[r] = 1; stop = 1; }
                    else rows *= 2;
                }
                if (!stop && j_max >= cols) {
                    if (cols * 2 > A.max_table_size / rows) { A.status[r] = 1; stop = 1; }
                    else cols *= 2;
                }
            }
            s_imin = i_min; s_imax = i_max; s_stop = stop;
            s_acc_min = (long long)k; s_acc_max = 0;   // `k` = unset sentinel (:214-216)
        }
        __syncthreads();
        if (s_stop) break;
        const int64_t i_min = s_imin, i_max = s_imax;
        const int64_t emin1 = s_emin1, emax1 = s_emax1, emin2 = s_emin2, emax2 = s_emax2;
        int32_t running = s_max;
        const int32_t *p1m = buf(k - 1, 0), *p2m = buf(k - 2, 0);
        int32_t *cm = buf(k, 0);
        for (int64_t base = i_min; base <= i_max; base += XD_THREADS) {
            const int64_t i = base + tid, j = k - i;
            const bool active = i <= i_max;
            int32_t ms = 0, g1s = 0, g2s = 0;
            uint32_t dir = 0;
            int32_t cellmax = INT32_MIN;
            if (active) {
                const bool has1 = j > 0 && i >= emin1 && i <= emax1;            // (i, j-1)
                const bool has2 = i > 0 && i - 1 >= emin1 && i - 1 <= emax1;    // (i-1, j)
                const bool hasd = i > 0 && j > 0 && i - 1 >= emin2 && i - 1 <= emax2;  // (i-1, j-1)
                int32_t sim = 0;
                if (i > 0 && j > 0) {
                    const uint64_t c1 = load_code(A.codes1, b1 + i - 1, A.code_bytes);
                    const uint64_t c2 = load_code(A.codes2, b2 + j - 1, A.code_bytes);
                    sim = A.use_matrix ? A.matrix[c1 * A.n_cols + c2] : (c1 == c2 ? A.match : A.mismatch);
                }
                const int32_t dm = hasd ? p2m[i - 1] : 0;
                const int32_t l_m = has1 ? p1m[i] : 0, u_m = has2 ? p1m[i - 1] : 0;
                if (!AFFINE) {
                    const int32_t a = dm != 0 ? wadd(dm, sim) : 0;
                    ms = xd_max3(a, wadd(l_m, A.gap_open), wadd(u_m, A.gap_open), dir);
                    cellmax = ms;
                } else {
                    const int32_t dg1 = hasd ? buf(k - 2, 1)[i - 1] : 0, dg2 = hasd ? buf(k - 2, 2)[i - 1] : 0;
                    const int32_t l_g1 = has1 ? buf(k - 1, 1)[i] : 0, u_g2 = has2 ? buf(k - 1, 2)[i - 1] : 0;
                    const int32_t mm = dm != 0 ? wadd(dm, sim) : 0, g1m = dg1 != 0 ? wadd(dg1, sim) : 0,
                                  g2m = dg2 != 0 ? wadd(dg2, sim) : 0;
                    ms = xd_max3(mm, g1m, g2m, dir);
                    const int32_t mg1 = wadd(l_m, A.gap_open), g1g1 = wadd(l_g1, A.gap_ext);
                    const int32_t mg2 = wadd(u_m, A.gap_open), g2g2 = wadd(u_g2, A.gap_ext);
                    g1s = max(mg1, g1g1); g2s = max(mg2, g2g2);
                    dir |= (mg1 >= g1g1 ? (uint32_t)A_MG1 : 0u) | (g1g1 >= mg1 ? (uint32_t)A_G1G1 : 0u);
                    dir |= (mg2 >= g2g2 ? (uint32_t)A_MG2 : 0u) | (g2g2 >= mg2 ? (uint32_t)A_G2G2 : 0u);
                    cellmax = max(ms, max(g1s, g2s));
                }
            }
            // exclusive prefix maximum over the chunk, seeded with the running maximum
            int32_t incl = cellmax;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl = max(incl, o);
            }
            if (lane == 31) s_warp_max[warp] = incl;
            __syncthreads();
            int32_t before = running;
            for (int w = 0; w < warp; w++) before = max(before, s_warp_max[w]);
            int32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane > 0) before = max(before, excl);
            int32_t chunk_max = running;
            for (int w = 0; w < XD_THREADS / 32; w++) chunk_max = max(chunk_max, s_warp_max[w]);
            if (active) {
                // pruning against the running maximum (:473-480, :593-619)
                bool valid = false;
                int32_t mx = before, req = wsub(mx, A.threshold);
                int32_t om = 0, og1 = 0, og2 = 0;
                if (ms >= req) { om = ms; valid = true; if (ms > mx) { mx = ms; req = wsub(mx, A.threshold); } }
                if (AFFINE) {
                    if (g1s >= req) { og1 = g1s; valid = true; if (g1s > mx) { mx = g1s; req = wsub(mx, A.threshold); } }
                    if (g2s >= req) { og2 = g2s; valid = true; }
                }
                cm[i] = om;
                if (AFFINE) { buf(k, 1)[i] = og1; buf(k, 2)[i] = og2; }
                if (valid) {
                    atomicMin(&s_acc_min, (long long)i);
                    atomicMax(&s_acc_max, (long long)i);
                    if (TRACED)
                        dirs[i * cols_final + j] = (uint8_t)(dir | ((om == final_max && om != 0) ? (uint32_t)DIR_MARK : 0u));
                }
            }
            running = chunk_max;
            __syncthreads();
        }
        if (tid == 0) {
            i_min_k0 = (int64_t)s_acc_min; i_max_k0 = (int64_t)s_acc_max;
            s_max = running;
            s_emin2 = emin1; s_emax2 = emax1;
            s_emin1 = i_min; s_emax1 = i_max;
        }
        __syncthreads();
    }
    if (tid == 0 && !TRACED) {
        A.max_score[r] = s_max;
        A.dims[2 * r] = rows; A.dims[2 * r + 1] = cols;
    }
}

}  // namespace bt
