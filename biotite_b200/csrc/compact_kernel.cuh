// Compaction of first-optimal traces for the trip back to the host.
//
// The traceback kernels write one step byte per alignment column into a buffer that has room
// for len1 + len2 bytes per pair (about 600 for BASELINE config 2, of which about 310 are used
// and each carries 2 bits).  What the reference's FFI returns - the (L, 2) trace of
// trace.rs:103-130 - is reconstructible from the first cell and the steps alone, so only
//     trace_len (int32), first (2 x int32) and ceil(L / 16) words of 2-bit steps
// cross PCIe: about 95 bytes per pair instead of 620.  Step k of a pair (end-to-start order, as
// in the byte form) sits in bits 2 (k % 16) .. of word k / 16; the words of consecutive pairs of
// a window follow each other without gaps, so that one cudaMemcpyAsync of exactly the used
// bytes returns them.
//
// Three launches per window: word counts per block of CP_BLOCK pairs, an exclusive scan of the
// block sums, and the packing proper, in which a block walks all words of its pairs so that
// both the byte reads and the word stores are coalesced.
#pragma once
#include "common.cuh"

namespace bt {

constexpr int CP_BLOCK = 256;   // pairs (and threads) per block

__host__ __device__ __forceinline__ int64_t ops2_words(int64_t len) { return (len + 15) >> 4; }

struct CompactArgs {
    const int64_t *trace_len;   // device, per pair, as written by the traceback kernels
    const uint8_t *ops;         // step bytes
    PairMap map;                // pair -> first step byte
    int64_t pair_begin, n;      // pairs [pair_begin, pair_begin + n)
    int32_t *len32;             // out: trace_len as int32, same indexing as trace_len
    int64_t *block_sums;        // scratch: n_blocks entries
    uint32_t *packed;           // out: the words of pair pair_begin start at packed[0]
    int64_t *total;             // out (device): number of words written
};

__global__ void __launch_bounds__(CP_BLOCK) compact_count_kernel(const CompactArgs A) {
    __shared__ int64_t warp_sum[CP_BLOCK / 32];
    const int64_t k = (int64_t)blockIdx.x * CP_BLOCK + threadIdx.x;
    int64_t w = 0;
    if (k < A.n) {
        const int64_t len = A.trace_len[A.pair_begin + k];
        A.len32[A.pair_begin + k] = (int32_t)len;
        w = ops2_words(len);
    }
    for (int off = 16; off > 0; off >>= 1) w += __shfl_down_sync(0xffffffffu, w, off);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t s = 0;
        for (int q = 0; q < CP_BLOCK / 32; q++) s += warp_sum[q];
        A.block_sums[blockIdx.x] = s;
    }
}

// exclusive scan of the block sums in place (one block), total to A.total
__global__ void __launch_bounds__(1024) compact_scan_kernel(const CompactArgs A, int64_t n_blocks) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t k = base + threadIdx.x;
        const int64_t v = k < n_blocks ? A.block_sums[k] : 0;
        int64_t x = v;
        for (int off = 1; off < 32; off <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= off) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t s = warp_sum[threadIdx.x];
            for (int off = 1; off < 32; off <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, s, off);
                if (threadIdx.x >= off) s += y;
            }
            warp_sum[threadIdx.x] = s;   // inclusive over the warps
        }
        __syncthreads();
        const int64_t before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0) + x - v;
        if (k < n_blocks) A.block_sums[k] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *A.total = carry;
}

__global__ void __launch_bounds__(CP_BLOCK) compact_pack_kernel(const CompactArgs A) {
    __shared__ int64_t first_word[CP_BLOCK + 1];   // exclusive scan of the block's word counts
    __shared__ int64_t warp_sum[CP_BLOCK / 32];
    const int64_t k = (int64_t)blockIdx.x * CP_BLOCK + threadIdx.x;
    const int64_t len = k < A.n ? A.trace_len[A.pair_begin + k] : 0;
    const int64_t w = ops2_words(len);
    int64_t x = w;
    for (int off = 1; off < 32; off <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) >= off) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    int64_t before = 0;
    for (int q = 0; q < (int)(threadIdx.x >> 5); q++) before += warp_sum[q];
    first_word[threadIdx.x + 1] = before + x;
    if (threadIdx.x == 0) first_word[0] = 0;
    __syncthreads();
    const int64_t block_words = first_word[CP_BLOCK];
    uint32_t *out = A.packed + A.block_sums[blockIdx.x];
    for (int64_t q = threadIdx.x; q < block_words; q += CP_BLOCK) {
        // pair of word q: last entry of first_word that is <= q
        int lo = 0, hi = CP_BLOCK;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (first_word[mid] <= q) lo = mid; else hi = mid;
        }
        const int64_t p = A.pair_begin + (int64_t)blockIdx.x * CP_BLOCK + lo;
        const int64_t step0 = (q - first_word[lo]) * 16;
        const int64_t plen = A.trace_len[p];
        const uint8_t *src = A.ops + A.map.ops(p) + step0;
        const int cnt = (int)min((int64_t)16, plen - step0);
        uint32_t word = 0;
        for (int b = 0; b < cnt; b++) word |= (uint32_t)(src[b] & 3u) << (2 * b);
        out[q] = word;
    }
}

inline cudaError_t compact_launch(const CompactArgs &A, cudaStream_t stream, int64_t *launches) {
    if (A.n <= 0) return cudaMemsetAsync(A.total, 0, 8, stream);
    const int64_t n_blocks = (A.n + CP_BLOCK - 1) / CP_BLOCK;
    compact_count_kernel<<<(unsigned)n_blocks, CP_BLOCK, 0, stream>>>(A);
    compact_scan_kernel<<<1, 1024, 0, stream>>>(A, n_blocks);
    compact_pack_kernel<<<(unsigned)n_blocks, CP_BLOCK, 0, stream>>>(A);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// Host: (L, 2) rows of one pair from the 2-bit form (trace.rs:103-130); mirrors expand_ops of
// host_trace.hpp.  Returns the number of rows written.
inline int64_t expand_ops2(int64_t len, int32_t first_i, int32_t first_j, const uint32_t *words,
                           bool banded, int64_t lower, int64_t *out) {
    if (len <= 0) return 0;
    int64_t s0 = (int64_t)first_i - 1;
    int64_t s1 = banded ? (int64_t)first_j + s0 + lower - 1 : (int64_t)first_j - 1;
    int64_t rows = 0;
    if (!(s0 == -1 && s1 == -1)) { out[0] = s0; out[1] = s1; rows = 1; }
    for (int64_t k = len - 2; k >= 0; k--) {
        const uint32_t op = (words[k >> 4] >> (2 * (k & 15))) & 3u;
        if (op == OP_DIAG) { s0++; s1++; out[2 * rows] = s0; out[2 * rows + 1] = s1; }
        else if (op == OP_TO1) { s1++; out[2 * rows] = -1; out[2 * rows + 1] = s1; }
        else { s0++; out[2 * rows] = s0; out[2 * rows + 1] = -1; }
        rows++;
    }
    return rows;
}

}  // namespace bt
