// Batch planner on the device (the packer's front end): what interseq_plan_window_records /
// interseq_order_batch do on host threads, as kernels on the stream that runs the fill.
//
//   keys    : key(p) = len1 << 16 | len2, value = pair id              (dplan_keys_kernel, or the pair
//             generators below, which also produce the index arrays of cross / triangular batches)
//   sort    : stable LSD radix sort of (key, id), ascending               (cub::DeviceRadixSort)
//   groups  : slot s of the window takes sorted[cnt - 1 - s] (longest first); a group is 64
//             consecutive slots; its cost is stripes(nmax) * mmax row steps (dplan_groups_kernel)
//   sort    : groups by descending cost, stable                           (cub::DeviceRadixSort)
//   layout  : per window (a fixed number of groups) exclusive prefix sums of the groups' scratch
//             needs -> WarpDesc; slot -> pair id table                    (dplan_layout_kernel,
//                                                                          dplan_order_kernel)
//
// The result is identical to the host planner's (same keys, both sorts stable), which the GPU tests
// check through bt_debug_device_plan.  Nothing here synchronises with the host: the streaming path
// plans window w on the compute stream right before it packs it, so the host does no per-pair work
// at all between the caller's arrays and the kernels.
//
// The radix sort is CUB's (header-only, part of the CUDA toolkit): it orders 32-bit keys for the
// planner and is not on the alignment path the kernels of this library implement.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "interseq_kernel.cuh"
#include "pool.cuh"

namespace bt {

constexpr int DP_THREADS = 256;
constexpr uint32_t DP_COST_MAX = 0x7fffffu;       // stripes * mmax <= 500 * 8000 < 2^23
constexpr uint32_t DP_KEY_OTHER = 0xffffffffu;    // pairs the packed kernels cannot take sort last

struct DplanTotals { int64_t dir_words, rowbuf, rowcodes, colcodes; };

__global__ void dplan_keys_kernel(const PairMap map, int64_t p0, int64_t cnt, uint32_t *keys, int32_t *vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < cnt; q += stride) {
        const int64_t p = p0 + q;
        keys[q] = (uint32_t)map.len1(p) << 16 | (uint32_t)map.len2(p);
        vals[q] = (int32_t)p;
    }
}

// How the pairs of a generated batch are enumerated (include/b200align.h, BtPairKind).
enum : int { DP_PAIRS_CROSS = 0, DP_PAIRS_TRIANGLE = 1 };

// first pair of row i of the upper triangle (i <= j) of an n x n matrix, row-major
__host__ __device__ __forceinline__ int64_t triangle_row_begin(int64_t i, int64_t n) {
    return i * n - i * (i - 1) / 2;
}

// row of pair p of the triangle: largest i with triangle_row_begin(i) <= p
__device__ __forceinline__ int64_t triangle_row_of(int64_t p, int64_t n) {
    // i = floor(((2n + 1) - sqrt((2n + 1)^2 - 8p)) / 2), corrected for rounding
    const double b = 2.0 * (double)n + 1.0;
    int64_t i = (int64_t)((b - sqrt(b * b - 8.0 * (double)p)) * 0.5);
    i = max((int64_t)0, min(i, n - 1));
    while (i > 0 && triangle_row_begin(i, n) > p) i--;
    while (i + 1 < n && triangle_row_begin(i + 1, n) <= p) i++;
    return i;
}

// Pair generator: pair p of a CROSS batch aligns sequence p / n2 of pool 1 with sequence p % n2 of
// pool 2; pair p of a TRIANGLE batch is the p-th (i, j), i <= j, of one pool in row-major order,
// oriented shorter sequence first (the caller guarantees a symmetric matrix).  Writes the index
// arrays, the planner's keys (DP_KEY_OTHER for pairs outside the packed kernels' exact range:
// empty or over-long sequences, shorter sequence above `short_max`), and adds up the DP cells.
__global__ void dplan_generate_kernel(int kind, const int64_t *off1, int64_t n1, const int64_t *off2, int64_t n2,
                                      int64_t n_pairs, int32_t short_max, int32_t long_max, int32_t *idx1,
                                      int32_t *idx2, uint32_t *keys, int32_t *vals,
                                      unsigned long long *cells_out, unsigned long long *others_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long cells = 0, others = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride) {
        int64_t a, b;
        if (kind == DP_PAIRS_CROSS) { a = p / n2; b = p - a * n2; }
        else { a = triangle_row_of(p, n1); b = a + (p - triangle_row_begin(a, n1)); }
        int64_t la = off1[a + 1] - off1[a], lb = off2[b + 1] - off2[b];
        if (kind == DP_PAIRS_TRIANGLE && la > lb) { const int64_t t = a; a = b; b = t; const int64_t u = la; la = lb; lb = u; }
        idx1[p] = (int32_t)a; idx2[p] = (int32_t)b;
        cells += (unsigned long long)(la * lb);
        const bool packed = la >= 1 && lb >= 1 && la <= long_max && lb <= long_max && min(la, lb) <= short_max;
        keys[p] = packed ? ((uint32_t)la << 16 | (uint32_t)lb) : DP_KEY_OTHER;
        vals[p] = (int32_t)p;
        others += packed ? 0 : 1;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, d);
        others += __shfl_down_sync(0xffffffffu, others, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) atomicAdd(cells_out, cells);
        if (others) atomicAdd(others_out, others);
    }
}

// One warp per group of 64 slots over the descending order of `cnt` sorted keys.
__global__ void dplan_groups_kernel(const uint32_t *skeys, int64_t cnt, int64_t groups, uint32_t *gkey, int32_t *gval,
                                    uint32_t *gnm) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= groups) return;
    const int64_t s0 = g * 64 + lane, s1 = s0 + 32;
    const uint32_t k0 = s0 < cnt ? skeys[cnt - 1 - s0] : 0u, k1 = s1 < cnt ? skeys[cnt - 1 - s1] : 0u;
    const unsigned n = __reduce_max_sync(0xffffffffu, max(k0 >> 16, k1 >> 16));
    const unsigned m = __reduce_max_sync(0xffffffffu, max(k0 & 0xffffu, k1 & 0xffffu));
    if (lane == 0) {
        const uint32_t nmax = max(n, 1u), mmax = max(m, 1u);
        const uint32_t cost = (nmax + IS_R - 1) / IS_R * mmax;
        gkey[g] = DP_COST_MAX - min(cost, DP_COST_MAX);
        gval[g] = (int32_t)g;
        gnm[g] = nmax << 16 | mmax;
    }
}

// exclusive prefix sums of four 64-bit values across a block of DP_THREADS threads
__device__ __forceinline__ void dplan_block_scan4(int64_t (&v)[4], int64_t (&total)[4], int64_t (*warp_sums)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int64_t x = v[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        incl[c] = x;
        if (lane == 31) warp_sums[warp][c] = x;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int64_t before = 0, all = 0;
        for (int w = 0; w < DP_THREADS / 32; w++) {
            const int64_t s = warp_sums[w][c];
            if (w < warp) before += s;
            all += s;
        }
        v[c] = before + incl[c] - v[c];
        total[c] = all;
    }
    __syncthreads();
}

// One block per window of `groups_per_window` groups (in the final order `perm`): descriptors with
// offsets relative to the window's scratch, and the window's scratch totals.
__global__ void __launch_bounds__(DP_THREADS)
dplan_layout_kernel(const int32_t *perm, const uint32_t *gnm, int64_t groups, int64_t groups_per_window, int no_dirs,
                    int dir_words_per_lane, WarpDesc *desc, DplanTotals *totals) {
    __shared__ int64_t warp_sums[DP_THREADS / 32][4];
    const int64_t g0 = (int64_t)blockIdx.x * groups_per_window, g1 = min(groups, g0 + groups_per_window);
    int64_t carry[4] = {0, 0, 0, 0};
    for (int64_t base = g0; base < g1; base += DP_THREADS) {
        const int64_t g = base + threadIdx.x;
        const bool valid = g < g1;
        const uint32_t nm = valid ? gnm[perm[g]] : 0u;
        const int64_t nmax = nm >> 16, mmax = nm & 0xffffu, stripes = (nmax + IS_R - 1) / IS_R;
        int64_t v[4] = {no_dirs ? 0 : stripes * mmax * 32 * dir_words_per_lane, mmax * 32, stripes * IS_R * 32, mmax * 32};
        int64_t total[4];
        dplan_block_scan4(v, total, warp_sums);
        if (valid) {
            WarpDesc d;
            d.dir_off = carry[0] + v[0]; d.rowbuf_off = carry[1] + v[1];
            d.rowcode_off = carry[2] + v[2]; d.colcode_off = carry[3] + v[3];
            d.nmax = (int32_t)nmax; d.mmax = (int32_t)mmax; d.stripes = (int32_t)stripes; d.pad = 0;
            desc[g] = d;
        }
#pragma unroll
        for (int c = 0; c < 4; c++) carry[c] += total[c];
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = DplanTotals{carry[0], carry[1], carry[2], carry[3]};
}

// slot -> pair id (-1 = padding) of the groups in their final order
__global__ void dplan_order_kernel(const int32_t *perm, const int32_t *svals, int64_t cnt, int64_t groups, int32_t *order) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; slot < groups * 64; slot += stride) {
        const int64_t s = (int64_t)perm[slot >> 6] * 64 + (slot & 63);
        order[slot] = s < cnt ? svals[cnt - 1 - s] : -1;
    }
}

// Buffers + launch sequence of the planner; one instance per stream in flight.
struct DevicePlanner {
    DevBuf keys_a, keys_b, vals_a, vals_b, gkey_a, gkey_b, gval_a, gval_b, gnm, temp;
    size_t temp_bytes = 0;
    int64_t cap_pairs = 0;

    static unsigned blocks_for(int64_t items) {
        return (unsigned)std::max<int64_t>(1, std::min<int64_t>((items + DP_THREADS - 1) / DP_THREADS, 148 * 16));
    }
    cudaError_t reserve(int64_t max_pairs) {
        if (max_pairs <= cap_pairs) return cudaSuccess;
        const int64_t max_groups = (max_pairs + 63) / 64;
        cudaError_t e;
        if ((e = keys_a.alloc((size_t)max_pairs * 4)) != cudaSuccess) return e;
        if ((e = keys_b.alloc((size_t)max_pairs * 4)) != cudaSuccess) return e;
        if ((e = vals_a.alloc((size_t)max_pairs * 4)) != cudaSuccess) return e;
        if ((e = vals_b.alloc((size_t)max_pairs * 4)) != cudaSuccess) return e;
        if ((e = gkey_a.alloc((size_t)max_groups * 4)) != cudaSuccess) return e;
        if ((e = gkey_b.alloc((size_t)max_groups * 4)) != cudaSuccess) return e;
        if ((e = gval_a.alloc((size_t)max_groups * 4)) != cudaSuccess) return e;
        if ((e = gval_b.alloc((size_t)max_groups * 4)) != cudaSuccess) return e;
        if ((e = gnm.alloc((size_t)max_groups * 4)) != cudaSuccess) return e;
        size_t need = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, need, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                            (const int32_t *)nullptr, (int32_t *)nullptr, max_pairs, 0, 32);
        if (e != cudaSuccess) return e;
        temp_bytes = need;
        if ((e = temp.alloc(need)) != cudaSuccess) return e;
        cap_pairs = max_pairs;
        return cudaSuccess;
    }
    // keys of the identity / indexed pairs [p0, p0 + cnt) of `dm` into keys_a / vals_a
    cudaError_t make_keys(const PairMap &dm, int64_t p0, int64_t cnt, cudaStream_t s, int64_t *launches) {
        dplan_keys_kernel<<<blocks_for(cnt), DP_THREADS, 0, s>>>(dm, p0, cnt, keys_a.as<uint32_t>(), vals_a.as<int32_t>());
        if (launches) (*launches)++;
        return cudaGetLastError();
    }
    // sorts keys_a / vals_a (cnt_all entries) into keys_b / vals_b
    cudaError_t sort_pairs(int64_t cnt_all, cudaStream_t s) {
        size_t bytes = temp_bytes;
        return cub::DeviceRadixSort::SortPairs(temp.ptr, bytes, keys_a.as<uint32_t>(), keys_b.as<uint32_t>(),
                                               vals_a.as<int32_t>(), vals_b.as<int32_t>(), cnt_all, 0, 32, s);
    }
    // groups, their order, descriptors and slot table of the first `cnt` sorted pairs
    cudaError_t layout(int64_t cnt, int64_t groups_per_window, int no_dirs, int dir_words_per_lane, WarpDesc *desc,
                       int32_t *order, DplanTotals *totals, cudaStream_t s, int64_t *launches) {
        const int64_t groups = (cnt + 63) / 64;
        if (groups == 0) return cudaSuccess;
        dplan_groups_kernel<<<(unsigned)((groups * 32 + DP_THREADS - 1) / DP_THREADS), DP_THREADS, 0, s>>>(
            keys_b.as<uint32_t>(), cnt, groups, gkey_a.as<uint32_t>(), gval_a.as<int32_t>(), gnm.as<uint32_t>());
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        size_t bytes = temp_bytes;
        e = cub::DeviceRadixSort::SortPairs(temp.ptr, bytes, gkey_a.as<uint32_t>(), gkey_b.as<uint32_t>(),
                                            gval_a.as<int32_t>(), gval_b.as<int32_t>(), groups, 0, 23, s);
        if (e != cudaSuccess) return e;
        const int64_t gpw = std::max<int64_t>(1, std::min(groups_per_window, groups));
        dplan_layout_kernel<<<(unsigned)((groups + gpw - 1) / gpw), DP_THREADS, 0, s>>>(
            gval_b.as<int32_t>(), gnm.as<uint32_t>(), groups, gpw, no_dirs, dir_words_per_lane, desc, totals);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        dplan_order_kernel<<<blocks_for(groups * 64), DP_THREADS, 0, s>>>(gval_b.as<int32_t>(), vals_b.as<int32_t>(), cnt,
                                                                          groups, order);
        if (launches) *launches += 3;
        return cudaGetLastError();
    }
    // the whole plan of the consecutive pairs [p0, p0 + cnt) as one window
    cudaError_t plan_window(const PairMap &dm, int64_t p0, int64_t cnt, int no_dirs, int dir_words_per_lane,
                            WarpDesc *desc, int32_t *order, cudaStream_t s, int64_t *launches) {
        cudaError_t e = make_keys(dm, p0, cnt, s, launches);
        if (e == cudaSuccess) e = sort_pairs(cnt, s);
        if (e == cudaSuccess) e = layout(cnt, (cnt + 63) / 64, no_dirs, dir_words_per_lane, desc, order, nullptr, s, launches);
        return e;
    }
};

}  // namespace bt
