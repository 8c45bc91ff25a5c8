// Per-query top-k of a score matrix on the device (BASELINE config 3 at specification size:
// 10^3 queries x 10^7 database sequences are 40 GB of int32 scores - only the best hits of every
// query should leave the GPU).  The reference has no such call: a caller of align_optimal(...,
// score_only=True) in a loop would sort its own list; the order used here is that of a stable
// descending sort: higher score first, equal scores by ascending database index.
//
// One block per row.  Four passes of an 8-bit radix select find the k-th largest value v of the
// row, then an ordered compaction keeps every element above v and the first (by index) of those
// equal to v until k are taken.  The row (4 bytes x chunk size, a few hundred KB) stays in L2
// between the passes.  Output per row: k (score, column) pairs in ascending column order; the
// host sorts the k survivors.
#pragma once
#include "common.cuh"

namespace bt {

constexpr int TK_THREADS = 256;

__device__ __forceinline__ uint32_t topk_key(int32_t v) { return (uint32_t)v ^ 0x80000000u; }   // unsigned order = signed order

// exclusive prefix sum of one int per thread over the block; `total` receives the block sum
__device__ __forceinline__ int topk_block_scan(int v, int &total, int *warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    int before = 0, all = 0;
    for (int w = 0; w < TK_THREADS / 32; w++) {
        const int s = warp_sums[w];
        if (w < warp) before += s;
        all += s;
    }
    __syncthreads();
    total = all;
    return before + x - v;
}

__global__ void __launch_bounds__(TK_THREADS)
topk_rows_kernel(const int32_t *scores, int64_t n_cols, int k, int32_t *out_scores, int32_t *out_cols) {
    __shared__ int hist[256];
    __shared__ int warp_sums[TK_THREADS / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_need;
    const int32_t *row = scores + (int64_t)blockIdx.x * n_cols;
    int32_t *os = out_scores + (int64_t)blockIdx.x * k;
    int32_t *oc = out_cols + (int64_t)blockIdx.x * k;
    const int kk = (int)min((int64_t)k, n_cols);
    // ---- radix select of the kk-th largest key: prefix = the bytes fixed so far, need = rank inside the bucket
    uint32_t prefix = 0, mask = 0;
    int need = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        for (int64_t c = threadIdx.x; c < n_cols; c += TK_THREADS) {
            const uint32_t key = topk_key(row[c]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 255;
            for (; b > 0; b--) {              // from the largest byte value down
                if (acc + hist[b] >= need) break;
                acc += hist[b];
            }
            s_prefix = prefix | ((uint32_t)b << shift);
            s_need = need - acc;
        }
        __syncthreads();
        prefix = s_prefix; need = s_need;
        mask |= 0xffu << shift;
        __syncthreads();
    }
    // prefix is the key of the kk-th largest element; `need` of the elements equal to it are taken
    const uint32_t vkey = prefix;
    int taken = 0, eq_taken = 0;
    for (int64_t c0 = 0; c0 < n_cols; c0 += TK_THREADS) {
        const int64_t c = c0 + threadIdx.x;
        uint32_t key = 0;
        int32_t v = 0;
        bool gt = false, eq = false;
        if (c < n_cols) { v = row[c]; key = topk_key(v); gt = key > vkey; eq = key == vkey; }
        int eq_total, all_total;
        const int eq_rank = topk_block_scan(eq ? 1 : 0, eq_total, warp_sums);
        const bool take = gt || (eq && eq_taken + eq_rank < need);
        const int pos = topk_block_scan(take ? 1 : 0, all_total, warp_sums);
        if (take) { os[taken + pos] = v; oc[taken + pos] = (int32_t)c; }
        taken += all_total;
        eq_taken += eq_total;
    }
    // rows shorter than k: pad with the smallest score and column -1
    for (int q = kk + threadIdx.x; q < k; q += TK_THREADS) { os[q] = INT32_MIN; oc[q] = -1; }
}

__global__ void scatter_scores_kernel(const int32_t *ids, const int32_t *values, int64_t count, int32_t *scores) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < count) scores[ids[q]] = values[q];
}

}  // namespace bt
