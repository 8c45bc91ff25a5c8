// libb200align - host side of the C ABI declared in include/b200align.h.
//
// Routing: every batch is served by hand-written sm_100a kernels only.
//   interseq_s16x2 : short pairs, affine/linear, packed 16-bit DPX lanes, 4-bit
//                    directions + GPU traceback (interseq_kernel.cuh)
//   general_i32    : everything else - any length, band, code width, scoring
//                    range; full direction sets (general_kernel.cuh)
// There is no CPU path: without a device every compute call fails.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200align.h"
#include "common.cuh"
#include "pool.cuh"
#include "general_kernel.cuh"
#include "wavefront_kernel.cuh"
#include "strip_kernel.cuh"
#include "host_trace.hpp"
#include "interseq_kernel.cuh"
#include "banded_kernel.cuh"
#include "microbench.cuh"
#include "emit_kernel.cuh"
#include "compact_kernel.cuh"
#include "xdrop_kernel.cuh"
#include "device_plan.cuh"
#include "topk_kernel.cuh"

using namespace bt;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local std::string g_error;

static int fail(int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return status;
}

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess) {                                                      \
            int st__ = err__ == cudaErrorMemoryAllocation ? BT_ERR_OOM : BT_ERR_CUDA;    \
            cudaGetLastError();                                                          \
            return fail(st__, "%s failed: %s", #expr, cudaGetErrorString(err__));        \
        }                                                                                \
    } while (0)

extern "C" const char *bt_last_error(void) { return g_error.c_str(); }
extern "C" const char *bt_version(void) { return "b200align 0.1 (sm_100a)"; }

extern "C" int bt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int bt_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return BT_OK;
}
extern "C" int bt_get_device(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}

static void *pinned_alloc(size_t bytes, unsigned flags) {
    void *p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, flags);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(BT_ERR_OOM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
extern "C" void *bt_pinned_alloc(size_t bytes) { return pinned_alloc(bytes, cudaHostAllocDefault); }
extern "C" void bt_pinned_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

// Makes `device` current for the calling thread and restores the caller's device on scope exit
// (a batch remembers the device it was created on; the calls that touch it may come from
// threads whose current device is another one).
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != device && cudaSetDevice(device) == cudaSuccess) changed = true;
    }
    ~DeviceGuard() { if (changed && prev >= 0) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------
// device buffers
// ---------------------------------------------------------------------------------------
static int check_params(const BtParams *P) {
    if (!P) return fail(BT_ERR_INVALID_ARG, "params is NULL");
    if (P->gap_open > 0 || (P->affine && P->gap_ext > 0))
        return fail(BT_ERR_INVALID_ARG, "Gap penalty must be negative");
    if (P->code_bytes != 1 && P->code_bytes != 2 && P->code_bytes != 4 && P->code_bytes != 8)
        return fail(BT_ERR_INVALID_ARG, "Unsupported dtype: code_bytes=%d", P->code_bytes);
    if (P->use_matrix && (P->n_rows <= 0 || P->n_cols <= 0))
        return fail(BT_ERR_INVALID_ARG, "substitution matrix has an empty shape");
    return BT_OK;
}

// pairwise.rs:585-602, banded.rs:310-314, :437-441
static int32_t reference_neg_inf(const BtParams &P, int32_t min_score) {
    int32_t ni = INT32_MIN;
    if (P.affine) {
        ni = wsub(wsub(ni, P.gap_open), P.gap_ext);
        if (min_score < 0) ni = wsub(ni, min_score);
    } else if (P.banded) {
        ni = wsub(ni, P.gap_open);
        if (min_score < 0) ni = wsub(ni, min_score);
    }
    return ni;
}

// ---------------------------------------------------------------------------------------
// batch object
// ---------------------------------------------------------------------------------------
enum Route { ROUTE_GENERAL = 0, ROUTE_INTERSEQ = 1, ROUTE_BANDED = 2 };

struct Chunk {
    int64_t begin, end;   // pair range
    int64_t dir_bytes;    // direction storage of the chunk
    int max_n, max_m, max_w;
};

struct BtBatch {
    BtParams P{};
    DevParams D{};
    int64_t n_pairs = 0;
    int score_only = 0;   // no directions, no traceback
    int stats = 0;        // traceback emits gap statistics, no step bytes
    int device = 0;
    Route route = ROUTE_GENERAL;
    bool wavefront = false;   // general route: intra-pair wavefront kernel (else the anti-diagonal kernel)
    int64_t cells = 0, launches = 0;
    int64_t total1 = 0, total2 = 0;
    int min_score = 0, max_score = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<int64_t> h_off1, h_off2, h_bands, h_dir_off, h_cand_off, h_ops_off;
    std::vector<int32_t> h_idx1, h_idx2;   // pair -> sequence of pool 1 / 2 (empty: identity)
    int64_t n_seq1 = 0, n_seq2 = 0, ops_total = 0;
    PairMap host_map() const {
        PairMap m;
        m.off1 = h_off1.data(); m.off2 = h_off2.data();
        m.idx1 = h_idx1.empty() ? nullptr : h_idx1.data();
        m.idx2 = h_idx2.empty() ? nullptr : h_idx2.data();
        m.ops_off = h_ops_off.empty() ? nullptr : h_ops_off.data();
        return m;
    }
    PairMap dev_map() const {
        PairMap m;
        m.off1 = off1.as<int64_t>(); m.off2 = off2.as<int64_t>();
        m.idx1 = (h_idx1.empty() && !dev_indexed) ? nullptr : idx1.as<int32_t>();
        m.idx2 = (h_idx2.empty() && !dev_indexed) ? nullptr : idx2.as<int32_t>();
        m.ops_off = h_ops_off.empty() ? nullptr : ops_off.as<int64_t>();
        return m;
    }
    std::vector<Chunk> chunks;
    DevBuf codes1, codes2, off1, off2, idx1, idx2, ops_off, bands, matrix, dir_off, cand_off, cand, dirs, gscratch;
    DevBuf score, start_idx, start_state, end_vals, trace_len, first, ops, flag, is_start;
    InterseqPlan interseq;  // state of the packed 16-bit route
    BandedPlan banded;      // state of the banded inter-sequence route
    bool ran = false;
    // generated batches (bt_batch_create_generated): the index arrays exist on the device only;
    // pairs outside the packed kernels' range run in `other` (pair other_ids[k] = its pair k)
    bool dev_indexed = false;
    int pair_kind = -1;
    std::unique_ptr<BtBatch> other;
    std::vector<int32_t> other_ids;
    // per-phase device timing: events[3k..3k+2] bracket fill and traceback of step k
    std::vector<cudaEvent_t> phase_events;
    size_t phase_used = 0;
    float fill_ms = 0.f, traceback_ms = 0.f;
    cudaError_t mark_phase() {
        if (phase_used == phase_events.size()) {
            cudaEvent_t e;
            cudaError_t err = cudaEventCreate(&e);
            if (err != cudaSuccess) return err;
            phase_events.push_back(e);
        }
        return cudaEventRecord(phase_events[phase_used++], stream);
    }
    ~BtBatch() {
        for (cudaEvent_t e : phase_events) cudaEventDestroy(e);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        // drain before the members (device buffers) return to the pool: error paths leave
        // copies and kernels queued
        if (stream) { cudaStreamSynchronize(stream); cudaStreamDestroy(stream); }
        cudaGetLastError();
    }
};

static const int64_t kDirBudget = 12ll << 30;   // direction bytes resident per chunk
static const int kMaxSmemInts = 50 * 1024;      // 200 KB of wavefront buffers + matrix

__global__ void validate_codes_kernel(const void *codes, int64_t n, int bytes, uint64_t limit,
                                      int *flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) bad |= load_code(codes, i, bytes) >= limit;
    if (bad) *flag = 1;
}

static BatchView make_view(BtBatch *b) {
    BatchView V{};
    V.codes1 = b->codes1.ptr; V.codes2 = b->codes2.ptr;
    V.map = b->dev_map();
    V.bands = b->P.banded ? b->bands.as<int64_t>() : nullptr;
    V.matrix = b->P.use_matrix ? b->matrix.as<int32_t>() : nullptr;
    V.dirs = b->dirs.as<uint8_t>();
    V.dir_off = b->dir_off.as<int64_t>();
    V.cand = b->cand.as<int32_t>();
    V.cand_off = b->cand_off.as<int64_t>();
    V.gscratch = nullptr; V.gscratch_stride = 0;
    V.score = b->score.as<int32_t>();
    V.start_idx = b->start_idx.as<int64_t>();
    V.start_state = b->start_state.as<int32_t>();
    V.end_vals = b->end_vals.as<int32_t>();
    V.trace_len = b->trace_len.as<int64_t>();
    V.first = b->first.as<int32_t>();
    V.ops = b->ops.as<uint8_t>();
    return V;
}

template <bool AFFINE, bool BANDED>
static cudaError_t launch_general_fill(const DevParams &D, const BatchView &V, int grid, int threads,
                                       size_t smem, cudaStream_t s) {
    auto kernel = general_fill_kernel<AFFINE, BANDED>;
    if (smem > 32 * 1024) {   // opt in early: static shared memory counts against the 48 KiB default too
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, threads, smem, s>>>(D, V);
    return cudaGetLastError();
}

static cudaError_t dispatch_general_fill(const DevParams &D, const BatchView &V, int grid, int threads,
                                         size_t smem, cudaStream_t s) {
    if (D.affine) {
        return D.banded ? launch_general_fill<true, true>(D, V, grid, threads, smem, s)
                        : launch_general_fill<true, false>(D, V, grid, threads, smem, s);
    }
    return D.banded ? launch_general_fill<false, true>(D, V, grid, threads, smem, s)
                    : launch_general_fill<false, false>(D, V, grid, threads, smem, s);
}

static cudaError_t dispatch_general_traceback(const DevParams &D, const BatchView &V, int64_t count,
                                              cudaStream_t s) {
    if (D.dir_pad) {
        // padded tables (wavefront route): one warp per pair walks a cached tile
        const int grid = (int)((count + 3) / 4);
        if (D.lean) {
            wavefront_traceback_kernel<true, false, true><<<grid, 128, 0, s>>>(D, V);
        } else if (D.affine) {
            if (D.banded) wavefront_traceback_kernel<true, true><<<grid, 128, 0, s>>>(D, V);
            else wavefront_traceback_kernel<true, false><<<grid, 128, 0, s>>>(D, V);
        } else {
            if (D.banded) wavefront_traceback_kernel<false, true><<<grid, 128, 0, s>>>(D, V);
            else wavefront_traceback_kernel<false, false><<<grid, 128, 0, s>>>(D, V);
        }
        return cudaGetLastError();
    }
    const int threads = 128;
    const int grid = (int)((count + threads - 1) / threads);
    if (D.affine) {
        if (D.banded) general_traceback_kernel<true, true><<<grid, threads, 0, s>>>(D, V);
        else general_traceback_kernel<true, false><<<grid, threads, 0, s>>>(D, V);
    } else {
        if (D.banded) general_traceback_kernel<false, true><<<grid, threads, 0, s>>>(D, V);
        else general_traceback_kernel<false, false><<<grid, threads, 0, s>>>(D, V);
    }
    return cudaGetLastError();
}

template <bool AFFINE, bool BANDED, int MODE, bool TRACED>
static cudaError_t launch_wavefront_fill(const DevParams &D, const BatchView &V, const WfArgs &A, int grid, int threads,
                                         size_t smem, cudaStream_t s) {
    // rows per lane: 2 under a band (its windows move two columns per lane), 4 on full tables
    auto kernel = wavefront_fill_kernel<AFFINE, BANDED, MODE, TRACED, BANDED ? 2 : 4>;
    if (smem > 32 * 1024) {   // opt in early: static shared memory counts against the 48 KiB default too
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, threads, smem, s>>>(D, V, A);
    return cudaGetLastError();
}

template <bool AFFINE, bool BANDED>
static cudaError_t dispatch_wavefront_mode(const DevParams &D, const BatchView &V, const WfArgs &A, int grid,
                                           int threads, size_t smem, cudaStream_t s) {
    if (D.local)
        return D.score_only ? launch_wavefront_fill<AFFINE, BANDED, WF_LOCAL, false>(D, V, A, grid, threads, smem, s)
                            : launch_wavefront_fill<AFFINE, BANDED, WF_LOCAL, true>(D, V, A, grid, threads, smem, s);
    if (!BANDED && !D.terminal)
        return D.score_only ? launch_wavefront_fill<AFFINE, false, WF_SEMI, false>(D, V, A, grid, threads, smem, s)
                            : launch_wavefront_fill<AFFINE, false, WF_SEMI, true>(D, V, A, grid, threads, smem, s);
    return D.score_only ? launch_wavefront_fill<AFFINE, BANDED, WF_GLOBAL, false>(D, V, A, grid, threads, smem, s)
                        : launch_wavefront_fill<AFFINE, BANDED, WF_GLOBAL, true>(D, V, A, grid, threads, smem, s);
}

template <int MODE, bool TRACED>
static cudaError_t launch_strip_fill(const DevParams &D, const BatchView &V, const WfArgs &A, int grid, int threads,
                                     size_t smem, cudaStream_t s) {
    auto kernel = A.lean_rows == 8 ? strip_fill_kernel<MODE, TRACED, 8> : strip_fill_kernel<MODE, TRACED, 4>;
    if (smem > 32 * 1024) {   // opt in early: static shared memory counts against the 48 KiB default too
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<grid, threads, smem, s>>>(D, V, A);
    return cudaGetLastError();
}

static cudaError_t dispatch_wavefront_fill(const DevParams &D, const BatchView &V, const WfArgs &A, int grid,
                                           int threads, size_t smem, cudaStream_t s) {
    if (D.lean) {
        // lean first-optimal kernel (strip_kernel.cuh): same launch shape, same outputs
        if (D.local)
            return D.score_only ? launch_strip_fill<WF_LOCAL, false>(D, V, A, grid, threads, smem, s)
                                : launch_strip_fill<WF_LOCAL, true>(D, V, A, grid, threads, smem, s);
        if (!D.terminal)
            return D.score_only ? launch_strip_fill<WF_SEMI, false>(D, V, A, grid, threads, smem, s)
                                : launch_strip_fill<WF_SEMI, true>(D, V, A, grid, threads, smem, s);
        return D.score_only ? launch_strip_fill<WF_GLOBAL, false>(D, V, A, grid, threads, smem, s)
                            : launch_strip_fill<WF_GLOBAL, true>(D, V, A, grid, threads, smem, s);
    }
    if (D.affine) {
        return D.banded ? dispatch_wavefront_mode<true, true>(D, V, A, grid, threads, smem, s)
                        : dispatch_wavefront_mode<true, false>(D, V, A, grid, threads, smem, s);
    }
    return D.banded ? dispatch_wavefront_mode<false, true>(D, V, A, grid, threads, smem, s)
                    : dispatch_wavefront_mode<false, false>(D, V, A, grid, threads, smem, s);
}

// Fill of pairs [begin, end) with the intra-pair wavefront kernel.  `max_n`, `max_m`: longest
// sequences among them; `gscratch` holds the stripe boundaries when they do not fit shared memory.
static int wavefront_launch(const BtParams &P, const DevParams &D, BatchView V, int64_t begin, int64_t end,
                            int max_n, int max_m, DevBuf &gscratch, cudaStream_t stream, int64_t *launches) {
    const int NS = P.affine ? 3 : 1;
    const int mat_ints = P.use_matrix ? P.n_rows * P.n_cols : 0;
    // rows per lane (launch_wavefront_fill); the lean kernel takes 8 once the pairs span several stripes
    // and there are enough of them to fill the machine (half the per-step overhead per cell), else 4
    // a handful of pairs cannot fill the SMs: more, shorter stripes (more warps in flight) win there
    const int R = P.banded ? 2 : (D.lean && max_n > 128 && end - begin >= 64 ? 8 : 4);
    const int rows = 32 * R;
    const int nw = std::max(1, std::min(WF_MAX_WARPS, (max_n + rows - 1) / rows));
    WfArgs A{};
    A.lean_rows = R;
    A.bnd_cols = max_m + 8;
    A.mat_smem = (mat_ints > 0 && mat_ints <= 4096) ? 1 : 0;
    const int64_t bnd_ints = (int64_t)nw * A.bnd_cols * NS;
    const int64_t smem_ints = (A.mat_smem ? mat_ints : 0) + bnd_ints;
    int64_t per_launch = end - begin;
    size_t smem = 0;
    if (smem_ints <= 24 * 1024) {
        smem = (size_t)smem_ints * 4;            // boundaries in shared memory (up to 96 KB)
    } else {
        smem = (size_t)(A.mat_smem ? mat_ints : 0) * 4;
        A.bnd_block = bnd_ints;
        per_launch = std::max<int64_t>(1, std::min<int64_t>(per_launch, (2ll << 30) / (bnd_ints * 4)));
        if (gscratch.bytes < (size_t)(per_launch * bnd_ints * 4))
            CUDA_TRY(gscratch.alloc((size_t)(per_launch * bnd_ints * 4)));
        A.bnd = gscratch.as<int32_t>();
    }
    for (int64_t p0 = begin; p0 < end; p0 += per_launch) {
        V.pair_begin = p0; V.pair_end = std::min(end, p0 + per_launch);
        const int grid = (int)(V.pair_end - V.pair_begin);
        CUDA_TRY(dispatch_wavefront_fill(D, V, A, grid, 32 * nw, smem, stream));
        if (launches) (*launches)++;
    }
    return BT_OK;
}

static int run_wavefront_chunk(BtBatch *b, const Chunk &c, const DevParams &D, BatchView V) {
    return wavefront_launch(b->P, D, V, c.begin, c.end, c.max_n, c.max_m, b->gscratch, b->stream, &b->launches);
}

// Runs fill (+ first-optimal traceback) of the general route for every chunk.
static int run_general(BtBatch *b, bool traceback) {
    const bool affine = b->P.affine != 0;
    const int NS = affine ? 3 : 1;
    const int mat_ints = b->P.use_matrix ? b->P.n_rows * b->P.n_cols : 0;
    for (const Chunk &c : b->chunks) {
        BatchView V = make_view(b);
        DevParams D = b->D;
        // batches return the first-optimal trace only: the lean kernel's direction bytes suffice
        {
            const char *forced = getenv("B200ALIGN_GENERAL");
            D.lean = b->wavefront && (traceback || b->score_only) && !D.mark &&
                     !(forced && strcmp(forced, "full") == 0) &&
                     strip_eligible(b->P, (int64_t)c.max_n + c.max_m, b->min_score, b->max_score);
        }
        if (b->wavefront) {
            // never-filled band cells and the padding behind every row read as "no direction"
            if (!b->score_only) CUDA_TRY(cudaMemsetAsync(b->dirs.ptr, 0, (size_t)c.dir_bytes, b->stream));
            CUDA_TRY(b->mark_phase());
            int st = run_wavefront_chunk(b, c, D, V);
            if (st) return st;
            CUDA_TRY(b->mark_phase());
            if (traceback && !b->score_only) {
                V.pair_begin = c.begin; V.pair_end = c.end;
                CUDA_TRY(dispatch_general_traceback(D, V, c.end - c.begin, b->stream));
                b->launches++;
            }
            CUDA_TRY(b->mark_phase());
            continue;
        }
        const int slots = b->P.banded ? c.max_w + 2 : c.max_n + 1;
        const int64_t wave_ints = 3ll * slots * NS;
        D.matrix_in_smem = (mat_ints > 0 && mat_ints <= 4096) ? 1 : 0;
        const int64_t smem_ints = wave_ints + (D.matrix_in_smem ? mat_ints : 0);
        const int diag_len = b->P.banded ? (c.max_w + 1) / 2 : std::min(c.max_n, c.max_m);
        int threads = std::min(512, std::max(64, (diag_len + 31) / 32 * 32));
        if (!b->score_only && b->P.banded)
            CUDA_TRY(cudaMemsetAsync(b->dirs.ptr, 0, (size_t)c.dir_bytes, b->stream));
        CUDA_TRY(b->mark_phase());
        if (smem_ints <= kMaxSmemInts) {
            V.pair_begin = c.begin; V.pair_end = c.end;
            CUDA_TRY(dispatch_general_fill(D, V, (int)(c.end - c.begin), threads,
                                           (size_t)smem_ints * 4, b->stream));
            b->launches++;
        } else {
            // wavefront buffers in global memory, a bounded number of blocks at a time
            const int64_t per_launch = 512;
            if (b->gscratch.bytes < (size_t)(per_launch * wave_ints * 4))
                CUDA_TRY(b->gscratch.alloc((size_t)(per_launch * wave_ints * 4)));
            V.gscratch = b->gscratch.as<int32_t>();
            V.gscratch_stride = wave_ints;
            for (int64_t p0 = c.begin; p0 < c.end; p0 += per_launch) {
                V.pair_begin = p0; V.pair_end = std::min(c.end, p0 + per_launch);
                CUDA_TRY(dispatch_general_fill(D, V, (int)(V.pair_end - V.pair_begin), threads,
                                               (size_t)(D.matrix_in_smem ? mat_ints : 0) * 4,
                                               b->stream));
                b->launches++;
            }
        }
        CUDA_TRY(b->mark_phase());
        if (traceback && !b->score_only) {
            V.pair_begin = c.begin; V.pair_end = c.end;
            CUDA_TRY(dispatch_general_traceback(D, V, c.end - c.begin, b->stream));
            b->launches++;
        }
        CUDA_TRY(b->mark_phase());
    }
    return BT_OK;
}

static int build_chunks(BtBatch *b) {
    const int64_t n = b->n_pairs;
    b->h_dir_off.assign(n, 0);
    b->h_cand_off.assign(n, 0);
    b->chunks.clear();
    Chunk cur{0, 0, 0, 0, 0, 0};
    int64_t cand_total = 0;
    b->cells = 0;
    const PairMap hm = b->host_map();
    for (int64_t k = 0; k < n; k++) {
        const int64_t l1 = hm.len1(k), l2 = hm.len2(k);
        if (l1 < 0 || l2 < 0) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
        if (l1 > INT32_MAX / 2 - 2 || l2 > INT32_MAX / 2 - 2)
            return fail(BT_ERR_INVALID_ARG, "sequence too long");
        int64_t bytes, w = 0;
        if (b->P.banded) {
            const int64_t lo = b->h_bands[2 * k], up = b->h_bands[2 * k + 1];
            w = up - lo + 1;
            if (w < 1) return fail(BT_ERR_INVALID_ARG, "The width of the band is 0");
            if (lo < -l1 || up > l2)
                return fail(BT_ERR_INVALID_ARG, "Alignment band is out of range (crop it to the sequences)");
            bytes = (l1 + 1) * table_stride(b->D.dir_pad, w + 2);
            // cells: in-band, in-sequence positions (banded.rs:199-207)
            for (int64_t i = 0; i < l1; i++) {
                int64_t j0 = std::max<int64_t>(i + lo, 0), j1 = std::min<int64_t>(i + up + 1, l2);
                if (j1 > j0) b->cells += j1 - j0;
            }
            b->h_cand_off[k] = cand_total;
            cand_total += 3 * (w + 2);
        } else {
            bytes = (l1 + 1) * table_stride(b->D.dir_pad, l2 + 1);
            b->cells += l1 * l2;
        }
        if (b->score_only) bytes = 0;
        if (cur.end > cur.begin && cur.dir_bytes + bytes > kDirBudget) {
            b->chunks.push_back(cur);
            cur = Chunk{k, k, 0, 0, 0, 0};
        }
        b->h_dir_off[k] = cur.dir_bytes;
        cur.dir_bytes += bytes;
        cur.end = k + 1;
        cur.max_n = std::max<int>(cur.max_n, (int)l1);
        cur.max_m = std::max<int>(cur.max_m, (int)l2);
        cur.max_w = std::max<int>(cur.max_w, (int)w);
    }
    if (cur.end > cur.begin) b->chunks.push_back(cur);
    int64_t max_dir = 0;
    for (const Chunk &c : b->chunks) max_dir = std::max(max_dir, c.dir_bytes);
    if (!b->score_only) CUDA_TRY(b->dirs.alloc((size_t)max_dir));
    if (b->P.banded) CUDA_TRY(b->cand.alloc((size_t)cand_total * 4));
    return BT_OK;
}

static int batch_create_impl(const BtParams *params, const void *codes1, const int64_t *off1, int64_t n_seq1,
                             const void *codes2, const int64_t *off2, int64_t n_seq2, const int32_t *idx1,
                             const int32_t *idx2, int64_t n_pairs, const int32_t *matrix, const int64_t *bands,
                             int32_t score_only, bool force_general, BtBatch **batch_out) {
    if (!batch_out) return fail(BT_ERR_INVALID_ARG, "batch_out is NULL");
    *batch_out = nullptr;
    const bool trace_create = getenv("B200ALIGN_TRACE") != nullptr;
    auto now_ms = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double tc0 = now_ms();
    int st = check_params(params);
    if (st) return st;
    if (n_pairs < 0 || n_seq1 < 0 || n_seq2 < 0 || !off1 || !off2) return fail(BT_ERR_INVALID_ARG, "bad pair offsets");
    if ((idx1 == nullptr) != (idx2 == nullptr)) return fail(BT_ERR_INVALID_ARG, "give both index arrays or none");
    if (!idx1 && (n_seq1 != n_pairs || n_seq2 != n_pairs))
        return fail(BT_ERR_INVALID_ARG, "without index arrays both pools must hold n_pairs sequences");
    if (n_pairs > INT32_MAX) return fail(BT_ERR_INVALID_ARG, "too many pairs for one batch (split it)");
    if (params->use_matrix && !matrix) return fail(BT_ERR_INVALID_ARG, "matrix is NULL");
    if (params->banded && !bands) return fail(BT_ERR_INVALID_ARG, "bands is NULL");
    if (bt_device_count() <= 0)
        return fail(BT_ERR_CUDA, "no CUDA device available (libb200align has no CPU fallback)");

    NvtxRange nvtx_create("bt_batch_create: checks, plan, H2D");
    std::unique_ptr<BtBatch> b(new BtBatch());
    b->P = *params;
    b->n_pairs = n_pairs;
    if (score_only < 0 || score_only > 2) return fail(BT_ERR_INVALID_ARG, "score_only must be 0, 1 or 2");
    b->score_only = score_only == 1 ? 1 : 0;
    b->stats = score_only == 2 ? 1 : 0;
    b->device = bt_get_device();
    b->n_seq1 = n_seq1; b->n_seq2 = n_seq2;
    b->h_off1.assign(off1, off1 + n_seq1 + 1);
    b->h_off2.assign(off2, off2 + n_seq2 + 1);
    if (params->banded) b->h_bands.assign(bands, bands + 2 * n_pairs);
    b->total1 = b->h_off1[n_seq1] - b->h_off1[0];
    b->total2 = b->h_off2[n_seq2] - b->h_off2[0];
    for (int64_t k = 0; k < n_seq1; k++)
        if (off1[k + 1] < off1[k]) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
    for (int64_t k = 0; k < n_seq2; k++)
        if (off2[k + 1] < off2[k]) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
    if (idx1) {
        // copy + range check on a few threads (10^8 pairs: 0.8 GB of indices)
        b->h_idx1.resize((size_t)n_pairs);
        b->h_idx2.resize((size_t)n_pairs);
        {
            const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), n_pairs / 1048576));
            std::vector<int> bad((size_t)n_threads, 0);
            interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t lo, int64_t hi) {
                int any = 0;
                for (int64_t k = lo; k < hi; k++) {
                    const int32_t a = idx1[k], c = idx2[k];
                    any |= (a < 0) | (a >= n_seq1) | (c < 0) | (c >= n_seq2);
                    b->h_idx1[(size_t)k] = a; b->h_idx2[(size_t)k] = c;
                }
                bad[(size_t)t] = any;
            });
            for (int v : bad) if (v) return fail(BT_ERR_INVALID_ARG, "pair index out of range");
        }
        if (score_only == 0) {   // step bytes of pair k start at the prefix sum of len1 + len2
            b->h_ops_off.resize((size_t)n_pairs);
            int64_t acc = 0;
            const PairMap hm0 = b->host_map();
            for (int64_t k = 0; k < n_pairs; k++) {
                b->h_ops_off[(size_t)k] = acc;
                acc += (b->h_off1[idx1[k] + 1] - b->h_off1[idx1[k]]) + (b->h_off2[idx2[k] + 1] - b->h_off2[idx2[k]]);
            }
            (void)hm0;
            b->ops_total = acc;
        }
    } else {
        b->ops_total = b->total1 + b->total2;
    }
    if (b->h_off1[0] != 0 || b->h_off2[0] != 0)
        return fail(BT_ERR_INVALID_ARG, "offsets must start at 0");

    int32_t mn = 0, mx = 0;
    if (params->use_matrix) {
        const int64_t cnt = (int64_t)params->n_rows * params->n_cols;
        mn = mx = matrix[0];
        for (int64_t k = 1; k < cnt; k++) { mn = std::min(mn, matrix[k]); mx = std::max(mx, matrix[k]); }
    } else {
        mn = std::min(params->match_score, params->mismatch_score);
        mx = std::max(params->match_score, params->mismatch_score);
    }
    b->min_score = mn; b->max_score = mx;

    DevParams &D = b->D;
    D.affine = params->affine; D.local = params->local;
    D.terminal = params->banded ? 1 : params->terminal_penalty;
    D.use_matrix = params->use_matrix; D.banded = params->banded; D.score_only = b->score_only;
    D.gap_open = params->gap_open; D.gap_ext = params->affine ? params->gap_ext : params->gap_open;
    D.match = params->match_score; D.mismatch = params->mismatch_score;
    D.neg_inf = reference_neg_inf(*params, mn);
    D.n_rows = params->n_rows; D.n_cols = params->n_cols; D.code_bytes = params->code_bytes;
    D.mark = 0; D.mark_value = 0; D.matrix_in_smem = 0; D.stats = b->stats;

    CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&b->ev0));
    CUDA_TRY(cudaEventCreate(&b->ev1));

    // route: packed 16-bit inter-sequence kernels when they are provably exact
    b->route = ROUTE_GENERAL;
    const PairMap hm = b->host_map();
    const double tc1 = now_ms();
    if (!force_general && interseq_eligible(*params, hm, n_pairs, mn, mx))
        b->route = ROUTE_INTERSEQ;
    else if (!force_general && !b->stats && banded_eligible(*params, hm, b->h_bands, n_pairs, mn, mx))
        b->route = ROUTE_BANDED;
    const double tc2 = now_ms();
    if (trace_create) fprintf(stderr, "[b200align] create: host copies / checks %.1f ms, routing %.1f ms\n", tc1 - tc0, tc2 - tc1);

    const size_t cb = (size_t)params->code_bytes;
    CUDA_TRY(b->codes1.alloc((size_t)b->total1 * cb));
    CUDA_TRY(b->codes2.alloc((size_t)b->total2 * cb));
    CUDA_TRY(b->off1.alloc((size_t)(n_seq1 + 1) * 8));
    CUDA_TRY(b->off2.alloc((size_t)(n_seq2 + 1) * 8));
    CUDA_TRY(b->score.alloc((size_t)n_pairs * 4));
    CUDA_TRY(b->flag.alloc(4));
    CUDA_TRY(cudaMemsetAsync(b->flag.ptr, 0, 4, b->stream));
    cudaStream_t s = b->stream;
    CUDA_TRY(cudaMemcpyAsync(b->codes1.ptr, codes1, (size_t)b->total1 * cb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->codes2.ptr, codes2, (size_t)b->total2 * cb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->off1.ptr, off1, (size_t)(n_seq1 + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->off2.ptr, off2, (size_t)(n_seq2 + 1) * 8, cudaMemcpyHostToDevice, s));
    if (idx1) {
        CUDA_TRY(b->idx1.alloc((size_t)n_pairs * 4));
        CUDA_TRY(b->idx2.alloc((size_t)n_pairs * 4));
        CUDA_TRY(cudaMemcpyAsync(b->idx1.ptr, idx1, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(b->idx2.ptr, idx2, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, s));
        if (!b->h_ops_off.empty()) {
            CUDA_TRY(b->ops_off.alloc((size_t)n_pairs * 8));
            CUDA_TRY(cudaMemcpyAsync(b->ops_off.ptr, b->h_ops_off.data(), (size_t)n_pairs * 8, cudaMemcpyHostToDevice, s));
        }
    }
    if (params->use_matrix) {
        const size_t mb = (size_t)params->n_rows * params->n_cols * 4;
        CUDA_TRY(b->matrix.alloc(mb));
        CUDA_TRY(cudaMemcpyAsync(b->matrix.ptr, matrix, mb, cudaMemcpyHostToDevice, s));
    }
    if (!b->score_only) {
        CUDA_TRY(b->trace_len.alloc((size_t)n_pairs * 8));
        CUDA_TRY(b->first.alloc((size_t)n_pairs * 8));
        if (!b->stats) CUDA_TRY(b->ops.alloc((size_t)b->ops_total));
    }

    if (b->route == ROUTE_INTERSEQ) {
        CUDA_TRY(b->is_start.alloc((size_t)n_pairs * 8));
        st = interseq_prepare(b->interseq, *params, hm, n_pairs, b->stats ? 2 : b->score_only, mn, mx,
                              &b->cells, 1, 4, s);
        if (trace_create) fprintf(stderr, "[b200align] create: uploads + packed plan %.1f ms\n", now_ms() - tc2);
        if (st == BT_ERR_OOM) return fail(st, "out of device memory (inter-sequence plan)");
        if (st) return fail(st, "inter-sequence plan failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else if (b->route == ROUTE_BANDED) {
        CUDA_TRY(b->bands.alloc((size_t)n_pairs * 16));
        CUDA_TRY(cudaMemcpyAsync(b->bands.ptr, bands, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, s));
        st = banded_prepare(b->banded, *params, hm, b->h_bands, n_pairs, b->score_only,
                            &b->cells, s);
        if (st == BT_ERR_OOM) return fail(st, "out of device memory (banded plan)");
        if (st) return fail(st, "banded plan failed: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
        // The wavefront kernel takes row 0 / column 0 from closed forms, which equal the
        // reference's running sums as long as no sum can wrap: a bound on every |value|.
        {
            int64_t span = 0;
            for (int64_t k = 0; k < n_pairs; k++) span = std::max(span, hm.len1(k) + hm.len2(k));
            const int64_t step = std::max<int64_t>({std::abs((int64_t)mn), std::abs((int64_t)mx),
                                                    std::abs((int64_t)params->gap_open), std::abs((int64_t)params->gap_ext)});
            const char *forced = getenv("B200ALIGN_GENERAL");
            b->wavefront = step * (span + 4) < (1ll << 30) && !(forced && strcmp(forced, "antidiag") == 0);
            D.dir_pad = b->wavefront ? 1 : 0;
        }
        st = build_chunks(b.get());
        if (st) return st;
        CUDA_TRY(b->dir_off.alloc((size_t)n_pairs * 8));
        CUDA_TRY(b->cand_off.alloc((size_t)n_pairs * 8));
        CUDA_TRY(b->start_idx.alloc((size_t)n_pairs * 8));
        CUDA_TRY(b->start_state.alloc((size_t)n_pairs * 4));
        CUDA_TRY(b->end_vals.alloc((size_t)n_pairs * 12));
        CUDA_TRY(cudaMemcpyAsync(b->dir_off.ptr, b->h_dir_off.data(), (size_t)n_pairs * 8,
                                 cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(b->cand_off.ptr, b->h_cand_off.data(), (size_t)n_pairs * 8,
                                 cudaMemcpyHostToDevice, s));
        if (params->banded) {
            CUDA_TRY(b->bands.alloc((size_t)n_pairs * 16));
            CUDA_TRY(cudaMemcpyAsync(b->bands.ptr, bands, (size_t)n_pairs * 16, cudaMemcpyHostToDevice, s));
        }
    }
    // code validation on the device (the reference panics on an out-of-range lookup, scoring.rs:63)
    if (params->use_matrix) {
        if (b->total1 > 0) {
            validate_codes_kernel<<<(int)std::min<int64_t>(1184, (b->total1 + 255) / 256), 256, 0, s>>>(
                b->codes1.ptr, b->total1, params->code_bytes, (uint64_t)params->n_rows, b->flag.as<int>());
        }
        if (b->total2 > 0) {
            validate_codes_kernel<<<(int)std::min<int64_t>(1184, (b->total2 + 255) / 256), 256, 0, s>>>(
                b->codes2.ptr, b->total2, params->code_bytes, (uint64_t)params->n_cols, b->flag.as<int>());
        }
        CUDA_TRY(cudaGetLastError());
    }
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, b->flag.ptr, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag) return fail(BT_ERR_INVALID_CODE, "a sequence code is outside the substitution matrix");
    *batch_out = b.release();
    return BT_OK;
}

extern "C" int bt_batch_create(const BtParams *params, const void *codes1, const int64_t *off1,
                               const void *codes2, const int64_t *off2, int64_t n_pairs,
                               const int32_t *matrix, const int64_t *bands, int32_t score_only,
                               BtBatch **batch_out) {
    return batch_create_impl(params, codes1, off1, n_pairs, codes2, off2, n_pairs, nullptr, nullptr, n_pairs, matrix,
                             bands, score_only, false, batch_out);
}

extern "C" int bt_batch_create_indexed(const BtParams *params, const void *codes1, const int64_t *off1,
                                       int64_t n_seq1, const void *codes2, const int64_t *off2, int64_t n_seq2,
                                       const int32_t *idx1, const int32_t *idx2, int64_t n_pairs,
                                       const int32_t *matrix, const int64_t *bands, int32_t score_only,
                                       BtBatch **batch_out) {
    if (!idx1 || !idx2) return fail(BT_ERR_INVALID_ARG, "index arrays are NULL");
    return batch_create_impl(params, codes1, off1, n_seq1, codes2, off2, n_seq2, idx1, idx2, n_pairs, matrix, bands,
                             score_only, false, batch_out);
}

extern "C" int64_t bt_batch_ops_bytes(const BtBatch *b) { return b ? b->ops_total : 0; }

extern "C" int bt_batch_run(BtBatch *b, float *device_ms) {
    if (!b) return fail(BT_ERR_INVALID_ARG, "batch is NULL");
    DeviceGuard device_guard(b->device);
    NvtxRange nvtx_run("bt_batch_run");
    b->launches = 0;
    b->phase_used = 0;
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    int st = BT_OK;
    if (b->n_pairs > 0) {
        if (b->route == ROUTE_INTERSEQ) {
            InterseqIO io{b->codes1.ptr, b->codes2.ptr, b->dev_map(),
                          b->matrix.as<int32_t>(), b->score.as<int32_t>(), b->trace_len.as<int64_t>(),
                          b->first.as<int32_t>(), b->ops.as<uint8_t>(), b->is_start.as<int32_t>()};
            cudaError_t e = cudaSuccess;
            for (const Window &win : b->interseq.windows) {
                if (e == cudaSuccess) e = b->mark_phase();
                if (e == cudaSuccess) { NvtxRange r("pack + fill"); e = interseq_fill_window(b->interseq, win, 0, io, b->stream, &b->launches); }
                if (e == cudaSuccess) e = b->mark_phase();
                if (e == cudaSuccess) { NvtxRange r("traceback"); e = interseq_traceback_window(b->interseq, win, 0, io, b->stream, &b->launches); }
                if (e == cudaSuccess) e = b->mark_phase();
            }
            if (e != cudaSuccess) st = fail(BT_ERR_CUDA, "inter-sequence kernels: %s", cudaGetErrorString(e));
        } else if (b->route == ROUTE_BANDED) {
            BandedIO io{b->codes1.ptr, b->codes2.ptr, b->dev_map(),
                        b->bands.as<int64_t>(), b->matrix.as<int32_t>(), b->score.as<int32_t>(),
                        b->trace_len.as<int64_t>(), b->first.as<int32_t>(), b->ops.as<uint8_t>()};
            cudaError_t e = cudaSuccess;
            for (const BandWindow &win : b->banded.windows) {
                if (e == cudaSuccess) e = b->mark_phase();
                if (e == cudaSuccess) { NvtxRange r("banded fill"); e = banded_fill_window(b->banded, win, io, b->stream, &b->launches); }
                if (e == cudaSuccess) e = b->mark_phase();
                if (e == cudaSuccess) { NvtxRange r("banded traceback"); e = banded_traceback_window(b->banded, win, io, b->stream, &b->launches); }
                if (e == cudaSuccess) e = b->mark_phase();
            }
            if (e != cudaSuccess) st = fail(BT_ERR_CUDA, "banded kernels: %s", cudaGetErrorString(e));
        } else {
            st = run_general(b, true);
        }
    }
    if (st) return st;
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (device_ms) CUDA_TRY(cudaEventElapsedTime(device_ms, b->ev0, b->ev1));
    b->fill_ms = b->traceback_ms = 0.f;
    for (size_t k = 0; k + 2 < b->phase_used; k += 3) {
        float f = 0.f, t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&f, b->phase_events[k], b->phase_events[k + 1]));
        CUDA_TRY(cudaEventElapsedTime(&t, b->phase_events[k + 1], b->phase_events[k + 2]));
        b->fill_ms += f;
        b->traceback_ms += t;
    }
    if (b->other) {   // the few pairs of a generated batch that need the 32-bit kernels
        float ms = 0.f;
        st = bt_batch_run(b->other.get(), &ms);
        if (st) return st;
        if (device_ms) *device_ms += ms;
        b->fill_ms += b->other->fill_ms; b->traceback_ms += b->other->traceback_ms;
        b->launches += b->other->launches;
    }
    b->ran = true;
    return BT_OK;
}

extern "C" int bt_batch_timings(const BtBatch *b, float *fill_ms, float *traceback_ms) {
    if (!b || !b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    if (fill_ms) *fill_ms = b->fill_ms;
    if (traceback_ms) *traceback_ms = b->traceback_ms;
    return BT_OK;
}

extern "C" int bt_batch_fetch_scores(BtBatch *b, int32_t *scores_out) {
    if (!b || !scores_out) return fail(BT_ERR_INVALID_ARG, "NULL argument");
    if (!b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    DeviceGuard device_guard(b->device);
    CUDA_TRY(cudaMemcpyAsync(scores_out, b->score.ptr, (size_t)b->n_pairs * 4, cudaMemcpyDeviceToHost,
                             b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (b->other) {
        std::vector<int32_t> part(b->other_ids.size());
        int st = bt_batch_fetch_scores(b->other.get(), part.data());
        if (st) return st;
        for (size_t k = 0; k < part.size(); k++) scores_out[b->other_ids[k]] = part[k];
    }
    return BT_OK;
}

// ---------------------------------------------------------------------------------------
// Generated batches: the pairs of a cross product / of the upper triangle of one pool are
// enumerated, planned and packed on the device (device_plan.cuh); the host never holds a
// per-pair array.  Score-only batches of the packed 16-bit route; the (few) pairs whose scores
// could leave 16 bits are collected by the generator and run as an indexed batch of the 32-bit
// kernels.
// ---------------------------------------------------------------------------------------
__global__ void symmetrize_kernel(const int32_t *scores, const int32_t *idx1, const int32_t *idx2, int64_t n_pairs,
                                  int64_t n, int32_t *matrix) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride) {
        const int64_t a = idx1[p], b = idx2[p];
        const int32_t v = scores[p];
        matrix[a * n + b] = v;
        matrix[b * n + a] = v;
    }
}

extern "C" int bt_batch_create_generated(const BtParams *params, const void *codes1, const int64_t *off1,
                                         int64_t n_seq1, const void *codes2, const int64_t *off2, int64_t n_seq2,
                                         int32_t pair_kind, const int32_t *matrix, int32_t score_only,
                                         BtBatch **batch_out) {
    if (!batch_out) return fail(BT_ERR_INVALID_ARG, "batch_out is NULL");
    *batch_out = nullptr;
    int st = check_params(params);
    if (st) return st;
    NvtxRange nvtx_create("bt_batch_create_generated: H2D, device plan");
    if (pair_kind != DP_PAIRS_CROSS && pair_kind != DP_PAIRS_TRIANGLE) return fail(BT_ERR_INVALID_ARG, "unknown pair_kind");
    if (score_only != 1) return fail(BT_ERR_INVALID_ARG, "generated batches are score-only (score_only = 1)");
    if (!off1 || !off2 || n_seq1 < 0 || n_seq2 < 0) return fail(BT_ERR_INVALID_ARG, "bad pool offsets");
    if (pair_kind == DP_PAIRS_TRIANGLE && n_seq1 != n_seq2) return fail(BT_ERR_INVALID_ARG, "a triangular batch takes one pool (twice)");
    if (n_seq1 > INT32_MAX || n_seq2 > INT32_MAX) return fail(BT_ERR_INVALID_ARG, "too many sequences");
    const double pairs_d = pair_kind == DP_PAIRS_CROSS ? (double)n_seq1 * (double)n_seq2 : (double)n_seq1 * ((double)n_seq1 + 1.0) / 2.0;
    if (pairs_d > (double)INT32_MAX) return fail(BT_ERR_INVALID_ARG, "too many pairs for one batch (split it)");
    const int64_t n_pairs = pair_kind == DP_PAIRS_CROSS ? n_seq1 * n_seq2 : n_seq1 * (n_seq1 + 1) / 2;
    // the modes the packed kernels serve (interseq_eligible); everything else: the indexed entry points
    if (params->banded || !params->use_matrix || !matrix || params->code_bytes != 1 || params->n_rows > 255 ||
        params->n_cols > 255 || (!params->affine && !params->local && !params->terminal_penalty))
        return fail(BT_ERR_INVALID_ARG, "generated batches need the packed kernels' modes (matrix scoring, 8-bit codes)");
    if (off1[0] != 0 || off2[0] != 0) return fail(BT_ERR_INVALID_ARG, "offsets must start at 0");
    for (int64_t k = 0; k < n_seq1; k++) if (off1[k + 1] < off1[k]) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
    for (int64_t k = 0; k < n_seq2; k++) if (off2[k + 1] < off2[k]) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
    int32_t mn = matrix[0], mx = matrix[0];
    for (int64_t k = 1; k < (int64_t)params->n_rows * params->n_cols; k++) { mn = std::min(mn, matrix[k]); mx = std::max(mx, matrix[k]); }
    if (params->affine ? (mn < -128 || mx > 127 || params->n_cols > 128)
                       : ((int64_t)params->n_rows * params->n_cols * IS_ROWW > 96 * 1024))
        return fail(BT_ERR_INVALID_ARG, "generated batches need a matrix the packed kernels can stage");
    if (bt_device_count() <= 0) return fail(BT_ERR_CUDA, "no CUDA device available (libb200align has no CPU fallback)");
    // exact 16-bit range (the proof of interseq_eligible, per pair): the longer sequence at most
    // IS_MAX_LEN, the shorter one small enough for the highest finite value
    int32_t long_max = IS_MAX_LEN, short_max = IS_MAX_LEN;
    {
        const int64_t open = params->gap_open, ext = params->affine ? params->gap_ext : params->gap_open;
        const int64_t lo_s = std::min<int64_t>(mn, 0), hi_s = std::max<int64_t>(mx, 0);
        const int64_t neg = -32768 - open - ext - lo_s;
        while (long_max > 0 && !(3 * open + (2 * (int64_t)long_max + 2 + IS_R) * ext + 2 * lo_s > neg + 1)) long_max /= 2;
        if (hi_s > 0) short_max = (int32_t)std::min<int64_t>(long_max, std::max<int64_t>(0, (32766 + open - hi_s) / hi_s - 1));
        else short_max = long_max;
    }

    std::unique_ptr<BtBatch> b(new BtBatch());
    b->P = *params; b->n_pairs = n_pairs; b->score_only = 1; b->stats = 0;
    b->device = bt_get_device();
    b->n_seq1 = n_seq1; b->n_seq2 = n_seq2;
    b->h_off1.assign(off1, off1 + n_seq1 + 1);
    b->h_off2.assign(off2, off2 + n_seq2 + 1);
    b->total1 = off1[n_seq1]; b->total2 = off2[n_seq2];
    b->min_score = mn; b->max_score = mx;
    b->dev_indexed = true; b->pair_kind = pair_kind;
    b->route = ROUTE_INTERSEQ;
    DevParams &D = b->D;
    D.affine = params->affine; D.local = params->local; D.terminal = params->terminal_penalty;
    D.use_matrix = 1; D.banded = 0; D.score_only = 1;
    D.gap_open = params->gap_open; D.gap_ext = params->affine ? params->gap_ext : params->gap_open;
    D.neg_inf = reference_neg_inf(*params, mn);
    D.n_rows = params->n_rows; D.n_cols = params->n_cols; D.code_bytes = 1;
    CUDA_TRY(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&b->ev0));
    CUDA_TRY(cudaEventCreate(&b->ev1));
    cudaStream_t s = b->stream;
    const size_t mb = (size_t)params->n_rows * params->n_cols * 4;
    CUDA_TRY(b->codes1.alloc((size_t)b->total1));
    CUDA_TRY(b->codes2.alloc((size_t)b->total2));
    CUDA_TRY(b->off1.alloc((size_t)(n_seq1 + 1) * 8));
    CUDA_TRY(b->off2.alloc((size_t)(n_seq2 + 1) * 8));
    CUDA_TRY(b->matrix.alloc(mb));
    CUDA_TRY(b->score.alloc((size_t)std::max<int64_t>(n_pairs, 1) * 4));
    CUDA_TRY(b->is_start.alloc((size_t)std::max<int64_t>(n_pairs, 1) * 8));
    CUDA_TRY(b->idx1.alloc((size_t)std::max<int64_t>(n_pairs, 1) * 4));
    CUDA_TRY(b->idx2.alloc((size_t)std::max<int64_t>(n_pairs, 1) * 4));
    CUDA_TRY(b->flag.alloc(32));
    CUDA_TRY(cudaMemsetAsync(b->flag.ptr, 0, 32, s));
    CUDA_TRY(cudaMemcpyAsync(b->codes1.ptr, codes1, (size_t)b->total1, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->codes2.ptr, codes2, (size_t)b->total2, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->off1.ptr, off1, (size_t)(n_seq1 + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->off2.ptr, off2, (size_t)(n_seq2 + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(b->matrix.ptr, matrix, mb, cudaMemcpyHostToDevice, s));
    if (b->total1 > 0)
        validate_codes_kernel<<<(int)std::min<int64_t>(1184, (b->total1 + 255) / 256), 256, 0, s>>>(
            b->codes1.ptr, b->total1, 1, (uint64_t)params->n_rows, b->flag.as<int>());
    if (b->total2 > 0)
        validate_codes_kernel<<<(int)std::min<int64_t>(1184, (b->total2 + 255) / 256), 256, 0, s>>>(
            b->codes2.ptr, b->total2, 1, (uint64_t)params->n_cols, b->flag.as<int>());
    CUDA_TRY(cudaGetLastError());

    InterseqPlan &plan = b->interseq;
    interseq_setup(plan, *params, n_pairs, 1, mn);
    const int64_t per_window = interseq_pairs_per_window(std::max<int64_t>(n_pairs, 1), 4, plan.warps_per_sm);
    const int64_t groups_per_window = per_window / 64;
    // flag words: [0] invalid code, [2..3] cells, [4..5] pairs for the 32-bit kernels
    unsigned long long *d_cells = (unsigned long long *)(b->flag.as<char>() + 8);
    unsigned long long *d_others = d_cells + 1;
    struct { int32_t flag, pad; unsigned long long cells, others; } counts{};
    DevicePlanner dp;
    // declared after the planner: on every exit the stream is drained before its buffers return to the pool
    struct Drain { cudaStream_t s; ~Drain() { cudaStreamSynchronize(s); cudaGetLastError(); } } drain{s};
    if (n_pairs > 0) {
        CUDA_TRY(dp.reserve(n_pairs));
        dplan_generate_kernel<<<DevicePlanner::blocks_for(n_pairs), DP_THREADS, 0, s>>>(
            pair_kind, b->off1.as<int64_t>(), n_seq1, b->off2.as<int64_t>(), n_seq2, n_pairs, short_max, long_max,
            b->idx1.as<int32_t>(), b->idx2.as<int32_t>(), dp.keys_a.as<uint32_t>(), dp.vals_a.as<int32_t>(), d_cells, d_others);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(dp.sort_pairs(n_pairs, s));
    }
    CUDA_TRY(cudaMemcpyAsync(&counts, b->flag.ptr, sizeof(counts), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (counts.flag) return fail(BT_ERR_INVALID_CODE, "a sequence code is outside the substitution matrix");
    b->cells = (int64_t)counts.cells;
    const int64_t n_other = (int64_t)counts.others, n_packed = n_pairs - n_other;
    const int64_t groups = (n_packed + 63) / 64;
    const int64_t n_windows = groups ? (groups + groups_per_window - 1) / groups_per_window : 0;
    plan.n_groups = groups;
    if (groups > 0) {
        DevBuf d_totals;
        CUDA_TRY(plan.d_order.alloc((size_t)groups * 64 * 4));
        CUDA_TRY(plan.d_desc.alloc((size_t)groups * sizeof(WarpDesc)));
        CUDA_TRY(d_totals.alloc((size_t)n_windows * sizeof(DplanTotals)));
        int64_t launches = 0;
        CUDA_TRY(dp.layout(n_packed, groups_per_window, 1, params->affine ? 4 : 2, plan.d_desc.as<WarpDesc>(),
                           plan.d_order.as<int32_t>(), d_totals.as<DplanTotals>(), s, &launches));
        std::vector<DplanTotals> totals((size_t)n_windows);
        CUDA_TRY(cudaMemcpyAsync(totals.data(), d_totals.ptr, (size_t)n_windows * sizeof(DplanTotals), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        plan.windows.clear();
        for (int64_t w = 0; w < n_windows; w++) {
            Window win{};
            win.pair_begin = 0; win.pair_end = 0;
            win.group_begin = w * groups_per_window; win.group_end = std::min(groups, (w + 1) * groups_per_window);
            win.dir_words = 0; win.rowbuf_entries = totals[(size_t)w].rowbuf;
            win.rowcodes = totals[(size_t)w].rowcodes; win.colcodes = totals[(size_t)w].colcodes;
            plan.max_rowbuf = std::max(plan.max_rowbuf, win.rowbuf_entries);
            plan.max_rowcodes = std::max(plan.max_rowcodes, win.rowcodes);
            plan.max_colcodes = std::max(plan.max_colcodes, win.colcodes);
            plan.windows.push_back(win);
        }
        SlotBuffers &sb = plan.slots[0];
        CUDA_TRY(sb.rowcodes.alloc((size_t)plan.max_rowcodes * 2));
        CUDA_TRY(sb.colcodes.alloc((size_t)(plan.max_colcodes + IS_SLACK * 32) * 2));
        CUDA_TRY(sb.rowbuf.alloc((size_t)(plan.max_rowbuf + IS_SLACK * 32) * 12));
        CUDA_TRY(sb.dirs.alloc(16));
        CUDA_TRY(sb.ctl.alloc(16));
        if (plan.C.mode == IS_SEMI) {
            CUDA_TRY(sb.lastrow.alloc((size_t)plan.max_rowbuf * 8));
            CUDA_TRY(sb.lastcol.alloc((size_t)plan.max_rowcodes * 8));
        }
    }
    if (n_other > 0) {
        // ids of the other pairs: the tail of the sorted order (ascending ids: the sort is stable)
        b->other_ids.resize((size_t)n_other);
        CUDA_TRY(cudaMemcpyAsync(b->other_ids.data(), dp.vals_b.as<int32_t>() + n_packed, (size_t)n_other * 4, cudaMemcpyDeviceToHost, s));
        std::vector<int32_t> i1((size_t)n_other), i2((size_t)n_other);
        CUDA_TRY(cudaStreamSynchronize(s));
        for (int64_t k = 0; k < n_other; k++) {
            const int64_t p = b->other_ids[(size_t)k];
            int64_t a, c;
            if (pair_kind == DP_PAIRS_CROSS) { a = p / n_seq2; c = p - a * n_seq2; }
            else {
                // row of the triangle by bisection over triangle_row_begin
                int64_t lo = 0, hi = n_seq1 - 1;
                while (lo < hi) { const int64_t mid = (lo + hi + 1) / 2; if (triangle_row_begin(mid, n_seq1) <= p) lo = mid; else hi = mid - 1; }
                a = lo; c = a + (p - triangle_row_begin(a, n_seq1));
                if (off1[a + 1] - off1[a] > off2[c + 1] - off2[c]) std::swap(a, c);
            }
            i1[(size_t)k] = (int32_t)a; i2[(size_t)k] = (int32_t)c;
        }
        BtBatch *raw = nullptr;
        st = batch_create_impl(params, codes1, off1, n_seq1, codes2, off2, n_seq2, i1.data(), i2.data(), n_other, matrix,
                               nullptr, 1, true, &raw);
        if (st) return st;
        b->other.reset(raw);
    }
    *batch_out = b.release();
    return BT_OK;
}

// Scores of a TRIANGLE batch as the symmetric (n, n) int32 matrix (built on the device).
extern "C" int bt_batch_fetch_score_matrix(BtBatch *b, int32_t *matrix_out) {
    if (!b || !matrix_out) return fail(BT_ERR_INVALID_ARG, "NULL argument");
    if (!b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    if (b->pair_kind != DP_PAIRS_TRIANGLE) return fail(BT_ERR_INVALID_ARG, "not a triangular generated batch");
    DeviceGuard device_guard(b->device);
    const int64_t n = b->n_seq1;
    if (n == 0) return BT_OK;
    DevBuf d_matrix;
    struct Drain { cudaStream_t s; ~Drain() { cudaStreamSynchronize(s); cudaGetLastError(); } } guard{b->stream};
    CUDA_TRY(d_matrix.alloc((size_t)(n * n) * 4));
    symmetrize_kernel<<<DevicePlanner::blocks_for(b->n_pairs), DP_THREADS, 0, b->stream>>>(
        b->score.as<int32_t>(), b->idx1.as<int32_t>(), b->idx2.as<int32_t>(), b->n_pairs, n, d_matrix.as<int32_t>());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(matrix_out, d_matrix.ptr, (size_t)(n * n) * 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (b->other) {
        std::vector<int32_t> part(b->other_ids.size());
        int st = bt_batch_fetch_scores(b->other.get(), part.data());
        if (st) return st;
        for (size_t k = 0; k < part.size(); k++) {
            const int64_t a = b->other->h_idx1[k], c = b->other->h_idx2[k];
            matrix_out[a * n + c] = part[k];
            matrix_out[c * n + a] = part[k];
        }
    }
    return BT_OK;
}

// The k best database hits of every query of a CROSS batch: the selection runs on the device over
// the (n_seq1, n_seq2) score matrix, 8 k bytes per query come back instead of 4 n_seq2.
extern "C" int bt_batch_fetch_topk(BtBatch *b, int32_t k, int32_t *scores_out, int32_t *index_out) {
    if (!b || !scores_out || !index_out) return fail(BT_ERR_INVALID_ARG, "NULL argument");
    if (!b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    if (b->pair_kind != DP_PAIRS_CROSS) return fail(BT_ERR_INVALID_ARG, "not a cross-product generated batch");
    if (k < 1) return fail(BT_ERR_INVALID_ARG, "k must be at least 1");
    DeviceGuard device_guard(b->device);
    const int64_t rows = b->n_seq1, cols = b->n_seq2;
    if (rows == 0) return BT_OK;
    cudaStream_t s = b->stream;
    struct Drain { cudaStream_t s; ~Drain() { cudaStreamSynchronize(s); cudaGetLastError(); } } guard{s};
    DevBuf d_ids, d_vals, d_scores, d_cols;
    if (b->other) {
        // scores of the pairs the 32-bit kernels ran go into the matrix first
        const int64_t cnt = (int64_t)b->other_ids.size();
        CUDA_TRY(d_ids.alloc((size_t)cnt * 4));
        CUDA_TRY(cudaMemcpyAsync(d_ids.ptr, b->other_ids.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice, s));
        // the other batch has its own stream: its scores are complete (bt_batch_run synchronised it)
        scatter_scores_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s>>>(d_ids.as<int32_t>(), b->other->score.as<int32_t>(),
                                                                            cnt, b->score.as<int32_t>());
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(d_scores.alloc((size_t)(rows * k) * 4));
    CUDA_TRY(d_cols.alloc((size_t)(rows * k) * 4));
    topk_rows_kernel<<<(unsigned)rows, TK_THREADS, 0, s>>>(b->score.as<int32_t>(), cols, k, d_scores.as<int32_t>(),
                                                           d_cols.as<int32_t>());
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(scores_out, d_scores.ptr, (size_t)(rows * k) * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(index_out, d_cols.ptr, (size_t)(rows * k) * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return BT_OK;
}

// Test hook: the device planner's result for the pairs of a plain batch as ONE window (slot ->
// pair id, and nmax << 16 | mmax per group), to be compared with the host planner / a restatement.
extern "C" int bt_debug_device_plan(const int64_t *off1, const int64_t *off2, int64_t n_pairs, int32_t score_only,
                                    int32_t affine, int32_t *order_out, uint32_t *group_nm_out, int64_t *dir_off_out) {
    if (!off1 || !off2 || n_pairs < 1 || !order_out) return fail(BT_ERR_INVALID_ARG, "NULL argument");
    if (bt_device_count() <= 0) return fail(BT_ERR_CUDA, "no CUDA device available (libb200align has no CPU fallback)");
    const int64_t groups = (n_pairs + 63) / 64;
    DevBuf d_off1, d_off2, d_order, d_desc;
    DevicePlanner dp;
    cudaStream_t s = nullptr;
    struct Drain { cudaStream_t *s; ~Drain() { if (*s) { cudaStreamSynchronize(*s); cudaStreamDestroy(*s); } cudaGetLastError(); } } guard{&s};
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    CUDA_TRY(d_off1.alloc((size_t)(n_pairs + 1) * 8));
    CUDA_TRY(d_off2.alloc((size_t)(n_pairs + 1) * 8));
    CUDA_TRY(d_order.alloc((size_t)groups * 64 * 4));
    CUDA_TRY(d_desc.alloc((size_t)groups * sizeof(WarpDesc)));
    CUDA_TRY(cudaMemcpyAsync(d_off1.ptr, off1, (size_t)(n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_off2.ptr, off2, (size_t)(n_pairs + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(dp.reserve(n_pairs));
    PairMap dm;
    dm.off1 = d_off1.as<int64_t>(); dm.off2 = d_off2.as<int64_t>();
    CUDA_TRY(dp.plan_window(dm, 0, n_pairs, score_only, affine ? 4 : 2, d_desc.as<WarpDesc>(), d_order.as<int32_t>(), s, nullptr));
    std::vector<WarpDesc> desc((size_t)groups);
    CUDA_TRY(cudaMemcpyAsync(order_out, d_order.ptr, (size_t)groups * 64 * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(desc.data(), d_desc.ptr, (size_t)groups * sizeof(WarpDesc), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    for (int64_t g = 0; g < groups; g++) {
        if (group_nm_out) group_nm_out[g] = (uint32_t)desc[(size_t)g].nmax << 16 | (uint32_t)desc[(size_t)g].mmax;
        if (dir_off_out) dir_off_out[g] = desc[(size_t)g].dir_off;
    }
    return BT_OK;
}

extern "C" int bt_batch_fetch_traces(BtBatch *b, int64_t *trace_len_out, int32_t *first_out,
                                     uint8_t *ops_out) {
    if (!b) return fail(BT_ERR_INVALID_ARG, "batch is NULL");
    if (b->score_only) return fail(BT_ERR_INVALID_ARG, "batch was created score_only");
    if (!b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    DeviceGuard device_guard(b->device);
    cudaStream_t s = b->stream;
    if (trace_len_out)
        CUDA_TRY(cudaMemcpyAsync(trace_len_out, b->trace_len.ptr, (size_t)b->n_pairs * 8, cudaMemcpyDeviceToHost, s));
    if (first_out)
        CUDA_TRY(cudaMemcpyAsync(first_out, b->first.ptr, (size_t)b->n_pairs * 8, cudaMemcpyDeviceToHost, s));
    if (ops_out) {
        if (b->stats) return fail(BT_ERR_INVALID_ARG, "a statistics batch has no step bytes");
        CUDA_TRY(cudaMemcpyAsync(ops_out, b->ops.ptr, (size_t)b->ops_total, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return BT_OK;
}

// ---------------------------------------------------------------------------------------
// wire formats straight from the device-resident compact traces (emit_kernel.cuh)
// ---------------------------------------------------------------------------------------
static int emit_args(BtBatch *b, const uint8_t *swapped, DevBuf &d_swapped, EmitArgs *A) {
    if (!b) return fail(BT_ERR_INVALID_ARG, "batch is NULL");
    if (b->score_only || b->stats) return fail(BT_ERR_INVALID_ARG, "the batch holds no traces");
    if (!b->ran) return fail(BT_ERR_INVALID_ARG, "bt_batch_run has not been called");
    if (swapped) {
        CUDA_TRY(d_swapped.alloc((size_t)b->n_pairs));
        CUDA_TRY(cudaMemcpyAsync(d_swapped.ptr, swapped, (size_t)b->n_pairs, cudaMemcpyHostToDevice, b->stream));
    }
    A->codes1 = b->codes1.ptr; A->codes2 = b->codes2.ptr;
    A->map = b->dev_map();
    A->bands = b->P.banded ? b->bands.as<int64_t>() : nullptr;
    A->swapped = swapped ? d_swapped.as<uint8_t>() : nullptr;
    A->trace_len = b->trace_len.as<int64_t>(); A->first = b->first.as<int32_t>(); A->ops = b->ops.as<uint8_t>();
    A->code_bytes = b->P.code_bytes;
    A->n_pairs = b->n_pairs;
    return BT_OK;
}

extern "C" int bt_batch_fetch_cigar(BtBatch *b, int32_t flags, const uint8_t *swapped, int32_t *n_ops_out,
                                    uint32_t *cigar_out) {
    if (!n_ops_out || !cigar_out) return fail(BT_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard device_guard(b ? b->device : 0);
    DevBuf d_swapped, d_nops, d_cigar;
    EmitArgs A{};
    int st = emit_args(b, swapped, d_swapped, &A);
    if (st) return st;
    if (b->n_pairs == 0) return BT_OK;
    CUDA_TRY(d_nops.alloc((size_t)b->n_pairs * 4));
    CUDA_TRY(d_cigar.alloc((size_t)std::max<int64_t>(b->ops_total, 1) * 4));
    emit_cigar_kernel<<<(unsigned)((b->n_pairs + 127) / 128), 128, 0, b->stream>>>(A, flags, d_nops.as<int32_t>(),
                                                                                   d_cigar.as<uint32_t>());
    CUDA_TRY(cudaGetLastError());
    b->launches++;
    CUDA_TRY(cudaMemcpyAsync(n_ops_out, d_nops.ptr, (size_t)b->n_pairs * 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(cigar_out, d_cigar.ptr, (size_t)b->ops_total * 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return BT_OK;
}

extern "C" int bt_batch_fetch_gapped(BtBatch *b, const uint8_t *symbols1, int32_t n_symbols1, const uint8_t *symbols2,
                                     int32_t n_symbols2, const uint8_t *swapped, int64_t *n_columns_out,
                                     uint8_t *gapped1_out, uint8_t *gapped2_out) {
    if (!symbols1 || !symbols2 || !n_columns_out || !gapped1_out || !gapped2_out || n_symbols1 < 1 || n_symbols2 < 1)
        return fail(BT_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard device_guard(b ? b->device : 0);
    DevBuf d_swapped, d_sym, d_cols, d_g1, d_g2;
    EmitArgs A{};
    int st = emit_args(b, swapped, d_swapped, &A);
    if (st) return st;
    if (b->n_pairs == 0) return BT_OK;
    const size_t total = (size_t)std::max<int64_t>(b->ops_total, 1);
    CUDA_TRY(d_sym.alloc((size_t)n_symbols1 + (size_t)n_symbols2));
    CUDA_TRY(d_cols.alloc((size_t)b->n_pairs * 8));
    CUDA_TRY(d_g1.alloc(total));
    CUDA_TRY(d_g2.alloc(total));
    CUDA_TRY(cudaMemcpyAsync(d_sym.ptr, symbols1, (size_t)n_symbols1, cudaMemcpyHostToDevice, b->stream));
    CUDA_TRY(cudaMemcpyAsync(d_sym.as<uint8_t>() + n_symbols1, symbols2, (size_t)n_symbols2, cudaMemcpyHostToDevice, b->stream));
    emit_gapped_kernel<<<(unsigned)((b->n_pairs + 127) / 128), 128, 0, b->stream>>>(
        A, d_sym.as<uint8_t>(), n_symbols1, d_sym.as<uint8_t>() + n_symbols1, n_symbols2, d_g1.as<uint8_t>(),
        d_g2.as<uint8_t>(), d_cols.as<int64_t>());
    CUDA_TRY(cudaGetLastError());
    b->launches++;
    CUDA_TRY(cudaMemcpyAsync(n_columns_out, d_cols.ptr, (size_t)b->n_pairs * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(gapped1_out, d_g1.ptr, (size_t)b->ops_total, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaMemcpyAsync(gapped2_out, d_g2.ptr, (size_t)b->ops_total, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return BT_OK;
}

// Host planning of a resident packed batch, without any device work (CPU tests of the planner).
extern "C" int bt_debug_plan(const BtParams *params, const int64_t *off1, const int64_t *off2, int64_t n_pairs,
                             int32_t score_only, int32_t waves_per_window, int32_t *order_out, int64_t order_cap,
                             int64_t *n_groups_out, int64_t *window_groups_out, int64_t *window_dir_bytes_out,
                             int64_t window_cap, int64_t *n_windows_out, int64_t *cells_out,
                             uint64_t *layout_hash_out) {
    int st = check_params(params);
    if (st) return st;
    if (!off1 || !off2 || n_pairs < 1 || !n_groups_out || !n_windows_out)
        return fail(BT_ERR_INVALID_ARG, "NULL argument");
    PairMap hm;
    hm.off1 = off1; hm.off2 = off2;
    const PairScan scan = interseq_scan(hm, n_pairs);
    if (!scan.in_range) return fail(BT_ERR_RANGE, "a sequence is outside the packed kernels' length range");
    InterseqPlan plan;
    int64_t cells = 0;
    interseq_plan_host(plan, *params, hm, n_pairs, score_only, 0, 0, &cells, std::max(1, (int)waves_per_window));
    *n_groups_out = plan.n_groups;
    *n_windows_out = (int64_t)plan.windows.size();
    if (cells_out) *cells_out = cells;
    if (layout_hash_out) {
        // FNV-1a over everything the kernels get from the plan besides the order: group descriptors,
        // window records and the scratch sizes
        uint64_t h = 1469598103934665603ull;
        auto mix = [&h](const void *p, size_t n) {
            const unsigned char *c = (const unsigned char *)p;
            for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ull; }
        };
        if (!plan.desc.empty()) mix(plan.desc.data(), plan.desc.size() * sizeof(WarpDesc));
        if (!plan.windows.empty()) mix(plan.windows.data(), plan.windows.size() * sizeof(Window));
        const int64_t sizes[4] = {plan.max_dir_words, plan.max_rowbuf, plan.max_rowcodes, plan.max_colcodes};
        mix(sizes, sizeof(sizes));
        *layout_hash_out = h;
    }
    if (order_out) {
        if ((int64_t)plan.order.size() > order_cap) return fail(BT_ERR_INVALID_ARG, "order_out is too small");
        memcpy(order_out, plan.order.data(), plan.order.size() * 4);
    }
    if (window_groups_out || window_dir_bytes_out) {
        if ((int64_t)plan.windows.size() > window_cap) return fail(BT_ERR_INVALID_ARG, "window arrays are too small");
        for (size_t w = 0; w < plan.windows.size(); w++) {
            if (window_groups_out) window_groups_out[w] = plan.windows[w].group_end - plan.windows[w].group_begin;
            if (window_dir_bytes_out) window_dir_bytes_out[w] = plan.windows[w].dir_words * 4;
        }
    }
    return BT_OK;
}

extern "C" int64_t bt_batch_cells(const BtBatch *b) { return b ? b->cells : 0; }
extern "C" int64_t bt_batch_launches(const BtBatch *b) { return b ? b->launches : 0; }
extern "C" int64_t bt_batch_windows(const BtBatch *b) { return b ? (int64_t)(b->phase_used / 3) : 0; }
extern "C" const char *bt_batch_kernel(const BtBatch *b) {
    if (!b) return "";
    return b->route == ROUTE_INTERSEQ ? "interseq_s16x2" : (b->route == ROUTE_BANDED ? "banded_i32" : "general_i32");
}
extern "C" void bt_batch_destroy(BtBatch *b) {
    if (!b) return;
    DeviceGuard device_guard(b->device);   // the buffers go back to the pool of their own device
    if (b->stream) cudaStreamSynchronize(b->stream);
    delete b;
}

// Where the results of a one-shot batch go.  Classic form: trace_len int64, one step byte per
// column at off1[k] + off2[k].  Compact form (compact_kernel.cuh): trace_len int32 and 2-bit
// steps, the words of pair k starting at ops2[ops2_off[k]].
struct BatchOut {
    int32_t *scores = nullptr;
    int32_t *first = nullptr;
    int64_t *trace_len64 = nullptr;
    uint8_t *ops = nullptr;
    bool compact = false;
    int32_t *trace_len32 = nullptr;
    uint32_t *ops2 = nullptr;
    int64_t ops2_cap = 0;       // words
    int64_t *ops2_off = nullptr;
    int64_t *ops2_words = nullptr;   // words used (highest end), optional
};

// First word of the compact steps of a run of pairs starting at pair p: an upper bound of
// sum ceil(L / 16) over the pairs before it that depends on the inputs only.
static inline int64_t ops2_base(const int64_t *off1, const int64_t *off2, int64_t p) {
    return (off1[p] + off2[p]) / 16 + 2 * p;
}

extern "C" int64_t bt_ops2_capacity(const int64_t *off1, const int64_t *off2, int64_t n_pairs) {
    if (!off1 || !off2 || n_pairs < 0) return 0;
    return ops2_base(off1, off2, n_pairs) + 16;
}

// One-shot path for the packed inter-sequence route: windows of consecutive pairs are
// streamed through two slots (stream + scratch each), so the H2D copy of window k+1 and the
// D2H copy of window k-1 overlap the kernels of window k.
static int align_batch_pipelined(const BtParams *params, const void *codes1, const int64_t *off1,
                                 const void *codes2, const int64_t *off2, int64_t n_pairs,
                                 const int32_t *matrix, int32_t score_only, int32_t mn, int32_t mx,
                                 const PairScan &scan, const BatchOut &out) {
    const int64_t *h_off1 = off1, *h_off2 = off2;   // the caller's host arrays stay valid during the call
    const int64_t total1 = h_off1[n_pairs], total2 = h_off2[n_pairs];
    const bool trace = getenv("B200ALIGN_TRACE") != nullptr;
    const bool compact = out.compact && !score_only;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    NvtxRange nvtx_call("bt_align_batch (streamed windows)");
    DevBuf d_codes1, d_codes2, d_off1, d_off2, d_matrix, d_score, d_len, d_first, d_ops, d_flag, d_start;
    DevBuf d_len32, d_packed, d_bsums[2], d_total;
    PinnedBuf h_total;
    DevicePlanner dplanner;   // declared before the streams: they are drained before its buffers go back to the pool
    CUDA_TRY(d_codes1.alloc((size_t)total1));
    CUDA_TRY(d_codes2.alloc((size_t)total2));
    CUDA_TRY(d_off1.alloc((size_t)(n_pairs + 1) * 8));
    CUDA_TRY(d_off2.alloc((size_t)(n_pairs + 1) * 8));
    const size_t mb = (size_t)params->n_rows * params->n_cols * 4;
    CUDA_TRY(d_matrix.alloc(mb));
    CUDA_TRY(d_score.alloc((size_t)n_pairs * 4));
    CUDA_TRY(d_flag.alloc(4));
    CUDA_TRY(d_start.alloc((size_t)n_pairs * 8));
    if (!score_only) {
        CUDA_TRY(d_len.alloc((size_t)n_pairs * 8));
        CUDA_TRY(d_first.alloc((size_t)n_pairs * 8));
        CUDA_TRY(d_ops.alloc((size_t)(total1 + total2)));
    }
    // Four streams: all input copies are queued up front on `h2d` (the device code buffers hold
    // the whole batch, so a window's input never waits for a scratch slot), the two compute
    // streams alternate windows / scratch slots, and `d2h` returns a window's results as soon as
    // its kernels are done - no compute stream ever waits for a copy of another window.
    // Declared after the buffers: its destructor drains the streams before any buffer goes back
    // to the pool, also on the error paths.
    struct Streams {
        cudaStream_t s[2] = {nullptr, nullptr}, h2d = nullptr, d2h = nullptr;
        std::vector<cudaEvent_t> in_ready, plan_ready, k_done, d2h_done;
        ~Streams() {
            for (cudaStream_t x : s) if (x) { cudaStreamSynchronize(x); cudaStreamDestroy(x); }
            if (h2d) { cudaStreamSynchronize(h2d); cudaStreamDestroy(h2d); }
            if (d2h) { cudaStreamSynchronize(d2h); cudaStreamDestroy(d2h); }
            for (cudaEvent_t e : in_ready) cudaEventDestroy(e);
            for (cudaEvent_t e : plan_ready) cudaEventDestroy(e);
            for (cudaEvent_t e : k_done) cudaEventDestroy(e);
            for (cudaEvent_t e : d2h_done) cudaEventDestroy(e);
            cudaGetLastError();
        }
    } st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st.s[0], cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&st.s[1], cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&st.h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&st.d2h, cudaStreamNonBlocking));

    const double t1 = now();
    // plan the windows on worker threads while the copies are in flight; windows are launched
    // as soon as they are planned
    InterseqPlan plan;
    PairMap hm;
    hm.off1 = h_off1; hm.off2 = h_off2;
    interseq_setup(plan, *params, n_pairs, score_only, mn);
    interseq_enable_tags(plan, *params, scan, mn, mx);
    const char *waves_env = getenv("B200ALIGN_WAVES");
    const int waves = waves_env ? std::max(1, atoi(waves_env)) : 3;
    // windows are planned on the device, on the compute stream, right before they are packed
    // (B200ALIGN_PLAN=host: the host-thread planner, kept for A/B measurements)
    const char *plan_env = getenv("B200ALIGN_PLAN");
    const bool device_plan = !(plan_env && strcmp(plan_env, "host") == 0);
    StreamingPlanner planner(plan, hm, n_pairs, waves, params->affine != 0, device_plan);
    if (device_plan) {
        int64_t max_np = 0;
        for (int64_t w = 0; w < planner.n_windows; w++)
            max_np = std::max(max_np, planner.bounds[(size_t)w + 1] - planner.bounds[(size_t)w]);
        CUDA_TRY(dplanner.reserve(max_np));
    }
    if (compact) {
        int64_t max_np = 0;
        for (int64_t w = 0; w < planner.n_windows; w++)
            max_np = std::max(max_np, planner.bounds[(size_t)w + 1] - planner.bounds[(size_t)w]);
        CUDA_TRY(d_len32.alloc((size_t)n_pairs * 4));
        CUDA_TRY(d_packed.alloc((size_t)(ops2_base(h_off1, h_off2, n_pairs) + 16) * 4));
        for (DevBuf &b : d_bsums) CUDA_TRY(b.alloc((size_t)((max_np + CP_BLOCK - 1) / CP_BLOCK + 1) * 8));
        CUDA_TRY(d_total.alloc((size_t)planner.n_windows * 8));
        CUDA_TRY(h_total.alloc((size_t)planner.n_windows * 8));
    }
    cudaEvent_t tev0 = nullptr;
    if (trace) { cudaEventCreate(&tev0); cudaEventRecord(tev0, st.h2d); }
    CUDA_TRY(cudaMemsetAsync(d_flag.ptr, 0, 4, st.h2d));
    CUDA_TRY(cudaMemcpyAsync(d_matrix.ptr, matrix, mb, cudaMemcpyHostToDevice, st.h2d));
    // input codes of window w on the copy stream; in_ready[w] fires when they have arrived
    auto enqueue_inputs = [&](int64_t w) -> cudaError_t {
        const int64_t p0 = planner.bounds[(size_t)w], p1 = planner.bounds[(size_t)w + 1];
        const int64_t a1 = h_off1[p0], b1 = h_off1[p1], a2 = h_off2[p0], b2 = h_off2[p1];
        // the window's slice of the offsets travels with its codes (a caller's pageable offset
        // arrays make these copies synchronous: small slices keep the first window's wait short)
        cudaError_t e = cudaMemcpyAsync(d_off1.as<int64_t>() + p0, off1 + p0, (size_t)(p1 - p0 + 1) * 8, cudaMemcpyHostToDevice, st.h2d);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_off2.as<int64_t>() + p0, off2 + p0, (size_t)(p1 - p0 + 1) * 8, cudaMemcpyHostToDevice, st.h2d);
        if (e == cudaSuccess) e = cudaMemcpyAsync((char *)d_codes1.ptr + a1, (const char *)codes1 + a1, (size_t)(b1 - a1), cudaMemcpyHostToDevice, st.h2d);
        if (e == cudaSuccess) e = cudaMemcpyAsync((char *)d_codes2.ptr + a2, (const char *)codes2 + a2, (size_t)(b2 - a2), cudaMemcpyHostToDevice, st.h2d);
        if (e == cudaSuccess) e = cudaEventRecord(st.in_ready[(size_t)w], st.h2d);
        return e;
    };
    for (int64_t w = 0; w < planner.n_windows; w++) {
        cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
        st.in_ready.push_back(e0);
        CUDA_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
        st.plan_ready.push_back(e2);
        CUDA_TRY(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        st.k_done.push_back(e1);
        cudaEvent_t e3 = nullptr;
        CUDA_TRY(cudaEventCreateWithFlags(&e3, cudaEventDisableTiming));
        st.d2h_done.push_back(e3);
    }
    // the copy engine serves its queue in order: input codes, then group order / descriptors
    // of a window (uploaded as soon as it is planned), then the next window's input codes, ...
    // (the host enqueues all of this within the first few milliseconds)
    CUDA_TRY(enqueue_inputs(0));
    CUDA_TRY(planner.allocate(1, scan.nmax, scan.mmax));
    planner.start();
    const double t2 = now();

    PairMap dm;
    dm.off1 = d_off1.as<int64_t>(); dm.off2 = d_off2.as<int64_t>();
    InterseqIO io{d_codes1.ptr, d_codes2.ptr, dm, d_matrix.as<int32_t>(),
                  d_score.as<int32_t>(), d_len.as<int64_t>(), d_first.as<int32_t>(), d_ops.as<uint8_t>(),
                  d_start.as<int32_t>(), d_flag.as<int>()};
    int64_t launches = 0;
    // B200ALIGN_TRACE: per window [start, h2d done, kernels done, d2h done]
    std::vector<cudaEvent_t> tev(trace ? (size_t)planner.n_windows * 4 : 0, nullptr);
    auto mark = [&](size_t w, int what, cudaStream_t s) {
        if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, s); tev[w * 4 + what] = e; }
    };
    int64_t words_end = 0;
    std::vector<double> host_t(trace ? (size_t)planner.n_windows * 3 : 0, 0.0);   // host clock: planned, enqueued, drained
    // results of window w to the host: everything a window returns is a contiguous range of the
    // output arrays.  The compact form waits (on the host) for the window's kernels, because the
    // size of its step words is only known then; by that time the next window is already queued.
    auto fill_ops2_off = [&](size_t w) {
        if (!out.ops2_off || !out.trace_len32) return;
        const Window &win = plan.windows[w];
        int64_t run = ops2_base(h_off1, h_off2, win.pair_begin);
        for (int64_t k = win.pair_begin; k < win.pair_end; k++) {
            out.ops2_off[k] = run;
            run += ops2_words(out.trace_len32[k]);
        }
    };
    auto drain = [&](size_t w) -> int {
        NvtxRange nvtx_drain("window D2H");
        const Window &win = plan.windows[w];
        const int64_t p0 = win.pair_begin, np = win.pair_end - win.pair_begin;
        cudaStream_t r = st.d2h;
        int64_t words = 0;
        if (compact) {
            CUDA_TRY(cudaEventSynchronize(st.k_done[w]));
            words = h_total.as<int64_t>()[w];
        } else {
            CUDA_TRY(cudaStreamWaitEvent(r, st.k_done[w], 0));
        }
        if (out.scores)
            CUDA_TRY(cudaMemcpyAsync(out.scores + p0, d_score.as<int32_t>() + p0, (size_t)np * 4, cudaMemcpyDeviceToHost, r));
        if (!score_only) {
            if (out.first)
                CUDA_TRY(cudaMemcpyAsync(out.first + 2 * p0, d_first.as<int32_t>() + 2 * p0, (size_t)np * 8, cudaMemcpyDeviceToHost, r));
            if (compact) {
                const int64_t base = ops2_base(h_off1, h_off2, p0);
                if (out.trace_len32)
                    CUDA_TRY(cudaMemcpyAsync(out.trace_len32 + p0, d_len32.as<int32_t>() + p0, (size_t)np * 4, cudaMemcpyDeviceToHost, r));
                if (out.ops2) {
                    if (base + words > out.ops2_cap) return fail(BT_ERR_INVALID_ARG, "ops2 buffer is too small (see bt_ops2_capacity)");
                    if (words > 0)
                        CUDA_TRY(cudaMemcpyAsync(out.ops2 + base, d_packed.as<uint32_t>() + base, (size_t)words * 4, cudaMemcpyDeviceToHost, r));
                }
                words_end = std::max(words_end, base + words);
            } else {
                const int64_t a = h_off1[p0] + h_off2[p0], b = h_off1[win.pair_end] + h_off2[win.pair_end];
                if (out.trace_len64)
                    CUDA_TRY(cudaMemcpyAsync(out.trace_len64 + p0, d_len.as<int64_t>() + p0, (size_t)np * 8, cudaMemcpyDeviceToHost, r));
                if (out.ops)
                    CUDA_TRY(cudaMemcpyAsync(out.ops + a, d_ops.as<uint8_t>() + a, (size_t)(b - a), cudaMemcpyDeviceToHost, r));
            }
        }
        mark(w, 3, r);
        if (compact) {
            CUDA_TRY(cudaEventRecord(st.d2h_done[w], r));
            // first word of every pair of the window BEFORE this one, whose lengths have arrived by
            // now: the window's base plus the running sum (keeps this pass off the end of the call)
            if (w > 0) {
                CUDA_TRY(cudaEventSynchronize(st.d2h_done[w - 1]));
                fill_ops2_off(w - 1);
            }
        }
        if (trace) host_t[w * 3 + 2] = now() - t0;
        return BT_OK;
    };
    for (size_t w = 0; w < plan.windows.size(); w++) {
        // one compute stream: the persistent fill kernels fill the machine, and a traceback beside
        // the next window's fill slows that fill by as much as it saves (measured)
        const int slot = 0;
        cudaStream_t s = st.s[0];
        { NvtxRange r("wait for window plan"); CUDA_TRY(planner.take((int64_t)w, st.h2d)); }
        if (trace) host_t[w * 3] = now() - t0;
        CUDA_TRY(cudaEventRecord(st.plan_ready[w], st.h2d));
        if ((int64_t)w + 1 < planner.n_windows) CUDA_TRY(enqueue_inputs((int64_t)w + 1));
        CUDA_TRY(cudaStreamWaitEvent(s, st.plan_ready[w], 0));
        mark(w, 0, s);
        const Window &win = plan.windows[w];
        const int64_t np = win.pair_end - win.pair_begin;
        CUDA_TRY(cudaStreamWaitEvent(s, st.in_ready[w], 0));
        mark(w, 1, s);
        if (device_plan) {
            NvtxRange r("device plan");
            CUDA_TRY(dplanner.plan_window(dm, win.pair_begin, np, plan.C.score_only, params->affine ? 4 : 2,
                                          plan.d_desc.as<WarpDesc>() + win.group_begin,
                                          plan.d_order.as<int32_t>() + win.group_begin * 64, s, &launches));
        }
        { NvtxRange r("pack + fill"); CUDA_TRY(interseq_fill_window(plan, win, slot, io, s, &launches)); }
        { NvtxRange r("traceback"); CUDA_TRY(interseq_traceback_window(plan, win, slot, io, s, &launches)); }
        if (compact) {
            CompactArgs ca{d_len.as<int64_t>(), d_ops.as<uint8_t>(), dm, win.pair_begin, np, d_len32.as<int32_t>(),
                           d_bsums[slot].as<int64_t>(),
                           d_packed.as<uint32_t>() + ops2_base(h_off1, h_off2, win.pair_begin),
                           d_total.as<int64_t>() + w};
            CUDA_TRY(compact_launch(ca, s, &launches));
            CUDA_TRY(cudaMemcpyAsync(h_total.as<int64_t>() + w, d_total.as<int64_t>() + w, 8, cudaMemcpyDeviceToHost, s));
        }
        mark(w, 2, s);
        CUDA_TRY(cudaEventRecord(st.k_done[w], s));
        if (trace) host_t[w * 3 + 1] = now() - t0;
        if (!compact) { int rc = drain(w); if (rc) return rc; }
        else if (w > 0) { int rc = drain(w - 1); if (rc) return rc; }
    }
    if (compact && !plan.windows.empty()) { int rc = drain(plan.windows.size() - 1); if (rc) return rc; }
    int flag = 0;
    const double t3 = now();
    CUDA_TRY(cudaStreamSynchronize(st.h2d));
    CUDA_TRY(cudaStreamSynchronize(st.s[0]));
    CUDA_TRY(cudaStreamSynchronize(st.s[1]));
    CUDA_TRY(cudaStreamSynchronize(st.d2h));
    if (trace) {
        for (size_t w = 0; w * 4 + 3 < tev.size(); w++) {
            if (!tev[w * 4] || !tev[w * 4 + 1] || !tev[w * 4 + 2] || !tev[w * 4 + 3]) continue;
            float a = 0, b = 0, c = 0, d = 0;
            cudaEventElapsedTime(&a, tev0, tev[w * 4]); cudaEventElapsedTime(&b, tev0, tev[w * 4 + 1]);
            cudaEventElapsedTime(&c, tev0, tev[w * 4 + 2]); cudaEventElapsedTime(&d, tev0, tev[w * 4 + 3]);
            fprintf(stderr, "[b200align]   window %zu: start %.2f  h2d done %.2f  kernels done %.2f  d2h done %.2f ms | host: planned %.2f enqueued %.2f drained %.2f\n",
                    w, a, b, c, d, host_t[w * 3], host_t[w * 3 + 1], host_t[w * 3 + 2]);
        }
        for (cudaEvent_t e : tev) if (e) cudaEventDestroy(e);
        cudaEventDestroy(tev0);
    }
    if (trace)
        fprintf(stderr, "[b200align] alloc %.2f ms, plan %.2f ms, enqueue %.2f ms, drain %.2f ms, windows %zu\n",
                t1 - t0, t2 - t1, t3 - t2, now() - t3, plan.windows.size());
    CUDA_TRY(cudaMemcpy(&flag, d_flag.ptr, 4, cudaMemcpyDeviceToHost));
    if (flag) return fail(BT_ERR_INVALID_CODE, "a sequence code is outside the substitution matrix");
    if (compact) {
        if (out.ops2_words) *out.ops2_words = words_end;
        if (!plan.windows.empty()) fill_ops2_off(plan.windows.size() - 1);
    }
    return BT_OK;
}

// The resident route behind the compact result form: run the batch, compact its traces as a
// whole, copy exactly the used words.
static int batch_fetch_compact(BtBatch *b, const BatchOut &out) {
    if (b->score_only || b->stats) return fail(BT_ERR_INVALID_ARG, "the batch holds no traces");
    if (!b->h_idx1.empty()) return fail(BT_ERR_INVALID_ARG, "compact traces of indexed batches are not supported");
    const int64_t n = b->n_pairs;
    DevBuf d_len32, d_packed, d_bsums, d_total;
    struct Drain { cudaStream_t s; ~Drain() { cudaStreamSynchronize(s); cudaGetLastError(); } } guard{b->stream};
    const int64_t cap = ops2_base(b->h_off1.data(), b->h_off2.data(), n) + 16;
    CUDA_TRY(d_len32.alloc((size_t)std::max<int64_t>(n, 1) * 4));
    CUDA_TRY(d_packed.alloc((size_t)cap * 4));
    CUDA_TRY(d_bsums.alloc((size_t)((n + CP_BLOCK - 1) / CP_BLOCK + 1) * 8));
    CUDA_TRY(d_total.alloc(8));
    CompactArgs ca{b->trace_len.as<int64_t>(), b->ops.as<uint8_t>(), b->dev_map(), 0, n, d_len32.as<int32_t>(),
                   d_bsums.as<int64_t>(), d_packed.as<uint32_t>(), d_total.as<int64_t>()};
    CUDA_TRY(compact_launch(ca, b->stream, &b->launches));
    int64_t words = 0;
    CUDA_TRY(cudaMemcpyAsync(&words, d_total.ptr, 8, cudaMemcpyDeviceToHost, b->stream));
    if (out.trace_len32) CUDA_TRY(cudaMemcpyAsync(out.trace_len32, d_len32.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (out.first) CUDA_TRY(cudaMemcpyAsync(out.first, b->first.ptr, (size_t)n * 8, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    if (out.ops2) {
        if (words > out.ops2_cap) return fail(BT_ERR_INVALID_ARG, "ops2 buffer is too small (see bt_ops2_capacity)");
        if (words > 0) CUDA_TRY(cudaMemcpyAsync(out.ops2, d_packed.ptr, (size_t)words * 4, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
    }
    if (out.ops2_words) *out.ops2_words = words;
    if (out.ops2_off && out.trace_len32) {
        int64_t run = 0;
        for (int64_t k = 0; k < n; k++) { out.ops2_off[k] = run; run += ops2_words(out.trace_len32[k]); }
    }
    return BT_OK;
}

static int align_batch_impl(const BtParams *params, const void *codes1, const int64_t *off1,
                            const void *codes2, const int64_t *off2, int64_t n_pairs,
                            const int32_t *matrix, const int64_t *bands, int32_t score_only,
                            const BatchOut &out) {
    int st = check_params(params);
    if (st) return st;
    if (n_pairs >= 64 && off1 && off2 && params->use_matrix && matrix && !params->banded && score_only != 2 &&
        bt_device_count() > 0) {
        const int64_t cnt = (int64_t)params->n_rows * params->n_cols;
        int32_t mn = matrix[0], mx = matrix[0];
        for (int64_t k = 1; k < cnt; k++) { mn = std::min(mn, matrix[k]); mx = std::max(mx, matrix[k]); }
        PairMap hm;
        hm.off1 = off1; hm.off2 = off2;
        // every length in [1, IS_MAX_LEN] (checked by the scan) makes the offsets monotone
        const PairScan scan = interseq_scan(hm, n_pairs);
        // the streaming path sizes its two scratch slots from the longest sequences of the batch;
        // batches with very long pairs go through the resident path, whose windows follow the
        // actual direction-table sizes
        const double slot_dirs = score_only ? 0.0 : (double)(interseq_pairs_per_window(n_pairs, 3, 12) / 64) *
                                                        (double)((scan.nmax + IS_R - 1) / IS_R) * (double)scan.mmax *
                                                        128.0 * (params->affine ? 4 : 2);
        if (off1[0] == 0 && off2[0] == 0 && slot_dirs <= (double)IS_DIR_BUDGET &&
            interseq_eligible(*params, hm, n_pairs, mn, mx, &scan))
            return align_batch_pipelined(params, codes1, off1, codes2, off2, n_pairs, matrix, score_only, mn, mx, scan, out);
    }
    BtBatch *raw = nullptr;
    st = bt_batch_create(params, codes1, off1, codes2, off2, n_pairs, matrix, bands, score_only, &raw);
    if (st) return st;
    std::unique_ptr<BtBatch> b(raw);
    st = bt_batch_run(b.get(), nullptr);
    if (st) return st;
    if (out.scores) { st = bt_batch_fetch_scores(b.get(), out.scores); if (st) return st; }
    if (score_only == 1) return BT_OK;
    if (out.compact && score_only == 0) return batch_fetch_compact(b.get(), out);
    if (out.trace_len64 || out.first || out.ops)
        st = bt_batch_fetch_traces(b.get(), out.trace_len64, out.first, score_only == 2 ? nullptr : out.ops);
    return st;
}

extern "C" int bt_align_batch(const BtParams *params, const void *codes1, const int64_t *off1,
                              const void *codes2, const int64_t *off2, int64_t n_pairs,
                              const int32_t *matrix, const int64_t *bands, int32_t score_only,
                              int32_t *scores_out, int64_t *trace_len_out, int32_t *first_out,
                              uint8_t *ops_out) {
    BatchOut out;
    out.scores = scores_out; out.trace_len64 = trace_len_out; out.first = first_out; out.ops = ops_out;
    return align_batch_impl(params, codes1, off1, codes2, off2, n_pairs, matrix, bands, score_only, out);
}

extern "C" int bt_align_batch_compact(const BtParams *params, const void *codes1, const int64_t *off1,
                                      const void *codes2, const int64_t *off2, int64_t n_pairs,
                                      const int32_t *matrix, const int64_t *bands, int32_t score_only,
                                      int32_t *scores_out, int32_t *trace_len_out, int32_t *first_out,
                                      uint32_t *ops2_out, int64_t ops2_capacity, int64_t *ops2_off_out,
                                      int64_t *ops2_words_out) {
    if (score_only != 0 && score_only != 1) return fail(BT_ERR_INVALID_ARG, "score_only must be 0 or 1");
    if (!score_only && ops2_out && !trace_len_out) return fail(BT_ERR_INVALID_ARG, "trace_len_out is NULL");
    BatchOut out;
    out.compact = true;
    out.scores = scores_out; out.trace_len32 = trace_len_out; out.first = first_out;
    out.ops2 = ops2_out; out.ops2_cap = ops2_capacity; out.ops2_off = ops2_off_out; out.ops2_words = ops2_words_out;
    return align_batch_impl(params, codes1, off1, codes2, off2, n_pairs, matrix, bands, score_only, out);
}

// ---------------------------------------------------------------------------------------
// One batch over several GPUs of the node, from one process: the pairs
// are cut into contiguous shards of equal DP cells, one host thread per device runs its shard
// through the one-shot path (own streams, own pooled buffers) and writes the results straight
// into disjoint slices of the caller's arrays.  No data-path collective: pairs are independent
// (pairwise.rs:161 takes one pair).  The step words of shard d start at the input-only bound
// ops2_base(first pair of d) + 16 d, so the shards cannot overlap.
// ---------------------------------------------------------------------------------------
extern "C" int bt_align_batch_multi(const int32_t *devices, int32_t n_devices, const BtParams *params,
                                    const void *codes1, const int64_t *off1, const void *codes2,
                                    const int64_t *off2, int64_t n_pairs, const int32_t *matrix,
                                    const int64_t *bands, int32_t score_only, int32_t *scores_out,
                                    int32_t *trace_len_out, int32_t *first_out, uint32_t *ops2_out,
                                    int64_t ops2_capacity, int64_t *ops2_off_out, int64_t *ops2_words_out,
                                    int64_t *shard_begin_out) {
    int st = check_params(params);
    if (st) return st;
    if (!off1 || !off2 || n_pairs < 0) return fail(BT_ERR_INVALID_ARG, "bad pair offsets");
    if (score_only != 0 && score_only != 1) return fail(BT_ERR_INVALID_ARG, "score_only must be 0 or 1");
    const int available = bt_device_count();
    if (available <= 0) return fail(BT_ERR_CUDA, "no CUDA device available (libb200align has no CPU fallback)");
    std::vector<int32_t> devs;
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else for (int d = 0; d < available; d++) devs.push_back(d);
    for (int32_t d : devs)
        if (d < 0 || d >= available) return fail(BT_ERR_INVALID_ARG, "device %d does not exist (%d devices)", d, available);
    const int nd = (int)devs.size();
    if (off1[0] != 0 || off2[0] != 0) return fail(BT_ERR_INVALID_ARG, "offsets must start at 0");
    for (int64_t k = 0; k < n_pairs; k++)
        if (off1[k + 1] < off1[k] || off2[k + 1] < off2[k]) return fail(BT_ERR_INVALID_ARG, "offsets must be non-decreasing");
    if (!score_only && ops2_out && ops2_capacity < bt_ops2_capacity(off1, off2, n_pairs) + 16 * nd)
        return fail(BT_ERR_INVALID_ARG, "ops2 buffer is too small (bt_ops2_capacity + 16 words per device)");
    // shards of equal DP cells (band: rows x width)
    std::vector<int64_t> cut((size_t)nd + 1, n_pairs);
    cut[0] = 0;
    {
        double total = 0.0;
        auto cells = [&](int64_t k) {
            const double n = (double)(off1[k + 1] - off1[k]), m = (double)(off2[k + 1] - off2[k]);
            return params->banded && bands ? n * (double)(bands[2 * k + 1] - bands[2 * k] + 1) : n * m;
        };
        for (int64_t k = 0; k < n_pairs; k++) total += cells(k);
        double run = 0.0;
        int d = 1;
        for (int64_t k = 0; k < n_pairs && d < nd; k++) {
            run += cells(k);
            while (d < nd && run >= total * d / nd) cut[(size_t)d++] = k + 1;
        }
    }
    if (shard_begin_out) for (int d = 0; d <= nd; d++) shard_begin_out[d] = cut[(size_t)d];
    const size_t cb = (size_t)params->code_bytes;
    std::vector<int> status((size_t)nd, BT_OK);
    std::vector<std::string> message((size_t)nd);
    std::vector<int64_t> words_end((size_t)nd, 0);
    auto work = [&](int d) {
        const int64_t p0 = cut[(size_t)d], p1 = cut[(size_t)d + 1], np = p1 - p0;
        if (np <= 0) return;
        if (cudaSetDevice(devs[(size_t)d]) != cudaSuccess) {
            cudaGetLastError();
            status[(size_t)d] = BT_ERR_CUDA; message[(size_t)d] = "cudaSetDevice failed";
            return;
        }
        // offsets of the shard, rebased to its first pair
        std::vector<int64_t> o1((size_t)np + 1), o2((size_t)np + 1);
        for (int64_t k = 0; k <= np; k++) { o1[(size_t)k] = off1[p0 + k] - off1[p0]; o2[(size_t)k] = off2[p0 + k] - off2[p0]; }
        const int64_t base = ops2_base(off1, off2, p0) + 16 * d;
        BatchOut out;
        out.compact = true;
        out.scores = scores_out ? scores_out + p0 : nullptr;
        if (!score_only) {
            out.trace_len32 = trace_len_out ? trace_len_out + p0 : nullptr;
            out.first = first_out ? first_out + 2 * p0 : nullptr;
            out.ops2 = ops2_out ? ops2_out + base : nullptr;
            out.ops2_cap = ops2_base(off1, off2, p1) + 16 * (d + 1) - base;
            out.ops2_off = ops2_off_out ? ops2_off_out + p0 : nullptr;
            out.ops2_words = &words_end[(size_t)d];
        }
        const int rc = align_batch_impl(params, (const char *)codes1 + (size_t)off1[p0] * cb, o1.data(),
                                        (const char *)codes2 + (size_t)off2[p0] * cb, o2.data(), np, matrix,
                                        bands ? bands + 2 * p0 : nullptr, score_only, out);
        if (rc) { status[(size_t)d] = rc; message[(size_t)d] = g_error; return; }
        if (out.ops2_off) for (int64_t k = 0; k < np; k++) out.ops2_off[k] += base;   // word index in the caller's array
        words_end[(size_t)d] += base;
    };
    {
        DeviceGuard keep(bt_get_device() >= 0 ? bt_get_device() : 0);   // the calling thread keeps its device
        std::vector<std::thread> threads;
        for (int d = 1; d < nd; d++) threads.emplace_back(work, d);
        work(0);
        for (std::thread &t : threads) t.join();
    }
    for (int d = 0; d < nd; d++)
        if (status[(size_t)d]) return fail(status[(size_t)d], "device %d: %s", devs[(size_t)d], message[(size_t)d].c_str());
    if (ops2_words_out) {
        int64_t end = 0;
        for (int64_t w : words_end) end = std::max(end, w);
        *ops2_words_out = end;
    }
    return BT_OK;
}

extern "C" int bt_expand_traces2(int64_t n_pairs, const int32_t *trace_len, const int32_t *first,
                                 const uint32_t *ops2, const int64_t *ops2_off, const int64_t *bands,
                                 const int64_t *rows_off, int64_t *out) {
    if (!trace_len || !first || !ops2 || !ops2_off || !rows_off || !out)
        return fail(BT_ERR_INVALID_ARG, "NULL argument");
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), n_pairs / 4096));
    std::vector<int64_t> bad((size_t)n_threads, -1);
    interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; k++) {
            const int64_t lower = bands ? bands[2 * k] : 0;
            const int64_t rows = expand_ops2(trace_len[k], first[2 * k], first[2 * k + 1], ops2 + ops2_off[k],
                                             bands != nullptr, lower, out + 2 * rows_off[k]);
            if (rows != trace_len[k] && bad[(size_t)t] < 0) bad[(size_t)t] = k;
        }
    });
    for (int64_t k : bad)
        if (k >= 0) return fail(BT_ERR_INVALID_ARG, "degenerate trace row in pair %lld", (long long)k);
    return BT_OK;
}

extern "C" int bt_expand_traces(int64_t n_pairs, const int64_t *trace_len, const int32_t *first,
                                const uint8_t *ops, const int64_t *ops_off, const int64_t *bands,
                                const int64_t *rows_off, int64_t *out) {
    if (!trace_len || !first || !ops || !ops_off || !rows_off || !out)
        return fail(BT_ERR_INVALID_ARG, "NULL argument");
    const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), n_pairs / 4096));
    std::vector<int64_t> bad((size_t)n_threads, -1);
    interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t lo, int64_t hi) {
        for (int64_t k = lo; k < hi; k++) {
            const int64_t lower = bands ? bands[2 * k] : 0;
            const int64_t rows = expand_ops(trace_len[k], first[2 * k], first[2 * k + 1], ops + ops_off[k],
                                            bands != nullptr, lower, out + 2 * rows_off[k]);
            if (rows != trace_len[k] && bad[(size_t)t] < 0) bad[(size_t)t] = k;
        }
    });
    for (int64_t k : bad)
        if (k >= 0) return fail(BT_ERR_INVALID_ARG, "degenerate trace row in pair %lld", (long long)k);
    return BT_OK;
}

// ---------------------------------------------------------------------------------------
// single pair: the reference's FFI contract, including co-optimal enumeration
// ---------------------------------------------------------------------------------------
struct BtTraceList {
    std::vector<std::vector<int64_t>> traces;  // each 2*L values
};

// ---------------------------------------------------------------------------------------
// X-drop extension of regions (align_local_gapped's native part, localgapped.rs:57-70)
// ---------------------------------------------------------------------------------------
template <bool TRACED>
static cudaError_t xdrop_launch(const XdropArgs &A, bool affine, int64_t n_regions, cudaStream_t s) {
    if (affine) xdrop_fill_kernel<true, TRACED><<<(unsigned)n_regions, XD_THREADS, 0, s>>>(A);
    else xdrop_fill_kernel<false, TRACED><<<(unsigned)n_regions, XD_THREADS, 0, s>>>(A);
    return cudaGetLastError();
}

extern "C" int bt_xdrop_regions(const BtParams *params, const void *codes1, const int64_t *off1,
                                const void *codes2, const int64_t *off2, int64_t n_regions,
                                const int32_t *matrix, int32_t threshold, int64_t max_table_size,
                                int32_t score_only, int64_t max_number, int32_t *scores_out,
                                int32_t *status_out, BtTraceList **lists_out) {
    int st = check_params(params);
    if (st) return st;
    if (n_regions < 0 || !off1 || !off2 || !scores_out || !status_out)
        return fail(BT_ERR_INVALID_ARG, "NULL argument");
    if (!score_only && !lists_out) return fail(BT_ERR_INVALID_ARG, "lists_out is NULL");
    if (params->use_matrix && !matrix) return fail(BT_ERR_INVALID_ARG, "matrix is NULL");
    if (threshold < 0) return fail(BT_ERR_INVALID_ARG, "The threshold value must be a non-negative integer");
    if (max_table_size <= 0) return fail(BT_ERR_INVALID_ARG, "Maximum table size must be a positve value");
    if (max_number < 1) return fail(BT_ERR_INVALID_ARG, "Maximum number of returned alignments must be at least 1");
    if (lists_out) for (int64_t r = 0; r < n_regions; r++) lists_out[r] = nullptr;
    if (n_regions == 0) return BT_OK;
    if (bt_device_count() <= 0) return fail(BT_ERR_CUDA, "no CUDA device available");
    const bool affine = params->affine != 0;
    const int cb = params->code_bytes;
    const int64_t total1 = off1[n_regions], total2 = off2[n_regions];
    struct StreamGuard {
        cudaStream_t s = nullptr;
        ~StreamGuard() { if (s) cudaStreamDestroy(s); }
    } sg;
    CUDA_TRY(cudaStreamCreateWithFlags(&sg.s, cudaStreamNonBlocking));
    cudaStream_t s = sg.s;
    DevBuf d_c1, d_c2, d_o1, d_o2, d_mat, d_roll, d_roll_off, d_max, d_dims, d_status, d_flag, d_dirs, d_dir_off;
    // destroyed before the buffers: nothing may be in flight when they return to the pool
    struct Drain { cudaStream_t s; ~Drain() { cudaStreamSynchronize(s); cudaGetLastError(); } } drain_guard{s};
    CUDA_TRY(d_c1.alloc((size_t)total1 * cb));
    CUDA_TRY(d_c2.alloc((size_t)total2 * cb));
    CUDA_TRY(d_o1.alloc((size_t)(n_regions + 1) * 8));
    CUDA_TRY(d_o2.alloc((size_t)(n_regions + 1) * 8));
    CUDA_TRY(d_flag.alloc(4));
    CUDA_TRY(cudaMemsetAsync(d_flag.ptr, 0, 4, s));
    if (total1) CUDA_TRY(cudaMemcpyAsync(d_c1.ptr, codes1, (size_t)total1 * cb, cudaMemcpyHostToDevice, s));
    if (total2) CUDA_TRY(cudaMemcpyAsync(d_c2.ptr, codes2, (size_t)total2 * cb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_o1.ptr, off1, (size_t)(n_regions + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_o2.ptr, off2, (size_t)(n_regions + 1) * 8, cudaMemcpyHostToDevice, s));
    if (params->use_matrix) {
        const size_t mb = (size_t)params->n_rows * params->n_cols * 4;
        CUDA_TRY(d_mat.alloc(mb));
        CUDA_TRY(cudaMemcpyAsync(d_mat.ptr, matrix, mb, cudaMemcpyHostToDevice, s));
        // the reference would panic on an out-of-range code (scoring.rs:63)
        if (total1) validate_codes_kernel<<<(int)std::min<int64_t>(592, (total1 + 255) / 256), 256, 0, s>>>(
            d_c1.ptr, total1, cb, (uint64_t)params->n_rows, d_flag.as<int>());
        if (total2) validate_codes_kernel<<<(int)std::min<int64_t>(592, (total2 + 255) / 256), 256, 0, s>>>(
            d_c2.ptr, total2, cb, (uint64_t)params->n_cols, d_flag.as<int>());
        CUDA_TRY(cudaGetLastError());
    }
    const int states = affine ? 3 : 1;
    std::vector<int64_t> roll_off((size_t)n_regions + 1, 0);
    for (int64_t r = 0; r < n_regions; r++) {
        if (off1[r + 1] < off1[r] || off2[r + 1] < off2[r]) return fail(BT_ERR_INVALID_ARG, "offsets must not decrease");
        roll_off[(size_t)r + 1] = roll_off[(size_t)r] + 3 * states * (off1[r + 1] - off1[r] + 1);
    }
    CUDA_TRY(d_roll.alloc((size_t)roll_off[(size_t)n_regions] * 4));
    CUDA_TRY(d_roll_off.alloc((size_t)(n_regions + 1) * 8));
    CUDA_TRY(cudaMemcpyAsync(d_roll_off.ptr, roll_off.data(), (size_t)(n_regions + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(d_max.alloc((size_t)n_regions * 4));
    CUDA_TRY(d_dims.alloc((size_t)n_regions * 16));
    CUDA_TRY(d_status.alloc((size_t)n_regions * 4));
    CUDA_TRY(cudaMemsetAsync(d_status.ptr, 0, (size_t)n_regions * 4, s));

    XdropArgs A{};
    A.codes1 = d_c1.ptr; A.codes2 = d_c2.ptr; A.off1 = d_o1.as<int64_t>(); A.off2 = d_o2.as<int64_t>();
    A.code_bytes = cb;
    A.matrix = d_mat.as<int32_t>(); A.n_cols = params->n_cols; A.use_matrix = params->use_matrix;
    A.match = params->match_score; A.mismatch = params->mismatch_score;
    A.gap_open = params->gap_open; A.gap_ext = affine ? params->gap_ext : params->gap_open;
    A.threshold = threshold; A.max_table_size = max_table_size;
    A.roll = d_roll.as<int32_t>(); A.roll_off = d_roll_off.as<int64_t>();
    A.max_score = d_max.as<int32_t>(); A.dims = d_dims.as<int64_t>(); A.status = d_status.as<int32_t>();
    CUDA_TRY(xdrop_launch<false>(A, affine, n_regions, s));
    std::vector<int32_t> h_max((size_t)n_regions);
    std::vector<int64_t> h_dims((size_t)n_regions * 2);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(h_max.data(), d_max.ptr, (size_t)n_regions * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(h_dims.data(), d_dims.ptr, (size_t)n_regions * 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(status_out, d_status.ptr, (size_t)n_regions * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(&flag, d_flag.ptr, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag) return fail(BT_ERR_INVALID_CODE, "a sequence code is outside the substitution matrix");
    for (int64_t r = 0; r < n_regions; r++) scores_out[r] = wsub(h_max[(size_t)r], wadd(threshold, 1));
    if (score_only) return BT_OK;

    // second pass: directions into tables shaped like the reference's (rows x cols)
    std::vector<int64_t> dir_off((size_t)n_regions + 1, 0);
    for (int64_t r = 0; r < n_regions; r++)
        dir_off[(size_t)r + 1] = dir_off[(size_t)r] + (status_out[r] ? 0 : h_dims[(size_t)2 * r] * h_dims[(size_t)2 * r + 1]);
    const int64_t dir_bytes = dir_off[(size_t)n_regions];
    cudaError_t e = d_dirs.alloc((size_t)dir_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(BT_ERR_OOM, "out of device memory (X-drop direction tables, %lld bytes)", (long long)dir_bytes); }
    CUDA_TRY(cudaMemsetAsync(d_dirs.ptr, 0, (size_t)std::max<int64_t>(dir_bytes, 1), s));
    CUDA_TRY(d_dir_off.alloc((size_t)(n_regions + 1) * 8));
    CUDA_TRY(cudaMemcpyAsync(d_dir_off.ptr, dir_off.data(), (size_t)(n_regions + 1) * 8, cudaMemcpyHostToDevice, s));
    A.dirs = d_dirs.as<uint8_t>(); A.dir_off = d_dir_off.as<int64_t>();
    CUDA_TRY(xdrop_launch<true>(A, affine, n_regions, s));
    std::vector<uint8_t> h_dirs((size_t)std::max<int64_t>(dir_bytes, 1));
    if (dir_bytes) CUDA_TRY(cudaMemcpyAsync(h_dirs.data(), d_dirs.ptr, (size_t)dir_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));

    // traceback on the host (trace.rs:21-142): starts = every cell holding the maximum, row-major
    for (int64_t r = 0; r < n_regions; r++) {
        if (status_out[r]) continue;
        const int64_t rows = h_dims[(size_t)2 * r], cols = h_dims[(size_t)2 * r + 1];
        const uint8_t *dir = h_dirs.data() + dir_off[(size_t)r];
        std::unique_ptr<BtTraceList> list(new BtTraceList());
        std::vector<Path> paths;
        const int64_t off[3] = {-cols - 1, -1, -cols};
        for (int64_t idx = 0; idx < rows * cols && (int64_t)paths.size() < max_number; idx++) {
            if (!(dir[idx] & DIR_MARK)) continue;
            enumerate_from(dir, affine, off, Start{idx, ST_MATCH}, max_number, max_number, paths);
        }
        if ((int64_t)paths.size() > max_number) paths.resize((size_t)max_number);
        for (const Path &path : paths) {
            std::vector<int64_t> trace_rows;
            path_to_trace(path, cols, false, 0, trace_rows);
            list->traces.push_back(std::move(trace_rows));
        }
        lists_out[r] = list.release();
    }
    return BT_OK;
}

extern "C" int64_t bt_tracelist_count(const BtTraceList *l) { return l ? (int64_t)l->traces.size() : 0; }
extern "C" int64_t bt_tracelist_length(const BtTraceList *l, int64_t k) {
    if (!l || k < 0 || k >= (int64_t)l->traces.size()) return -1;
    return (int64_t)l->traces[k].size() / 2;
}
extern "C" int bt_tracelist_copy(const BtTraceList *l, int64_t k, int64_t *out) {
    if (!l || k < 0 || k >= (int64_t)l->traces.size() || !out) return fail(BT_ERR_INVALID_ARG, "bad trace index");
    if (!l->traces[k].empty()) memcpy(out, l->traces[k].data(), l->traces[k].size() * 8);
    return BT_OK;
}
extern "C" void bt_tracelist_free(BtTraceList *l) { delete l; }

// Co-optimal enumeration on the host over a direction table copied back from the GPU
// (trace.rs:21-92; starts: pairwise.rs:532-559, :845-881, banded.rs:377-414, :550-593).
static void enumerate_traces(const BtParams &P, const uint8_t *dir, const int32_t *cand, const int32_t endv[3],
                             int32_t score, int64_t len1, int64_t len2, int64_t lower_diag, int64_t upper_diag,
                             int64_t rs, int64_t max_number, BtTraceList &list) {
    const bool banded = P.banded != 0, affine = P.affine != 0;
    const int64_t W = upper_diag - lower_diag + 1;
    const int64_t table = (len1 + 1) * rs;
    std::vector<Start> starts;
    if (P.local) {
        for (int64_t idx = 0; idx < table && (int64_t)starts.size() < max_number; idx++) {
            bool is_start = (dir[idx] & DIR_MARK) != 0;
            if (!is_start && banded && score == 0) {
                // never-filled in-band cells keep the edge value M = 0 (banded.rs:184)
                const int64_t i = idx / rs, c = idx % rs;
                if (c >= 1 && c <= W) {
                    const int64_t sj = c + i - 2 + lower_diag;
                    is_start = (i == 0) || sj < 0 || sj >= len2;
                }
            }
            if (is_start) starts.push_back({idx, ST_MATCH});
        }
    } else if (!banded) {
        const int64_t corner = len1 * rs + len2;
        if (affine) {
            for (int stt = 0; stt < 3; stt++)
                if (endv[stt] == score) starts.push_back({corner, stt});
        } else {
            starts.push_back({corner, 0});
        }
    } else {
        const int ns = affine ? 3 : 1;
        for (int stt = 0; stt < ns; stt++) {
            for (int64_t j = 1; j <= W; j++) {
                if (cand[(size_t)(stt * (W + 2) + j)] != score) continue;
                const int64_t sj = j + (len1 - 1) + lower_diag - 1;
                const int64_t i = sj < len2 ? len1 : len2 - j - lower_diag + 1;
                starts.push_back({i * rs + j, stt});
            }
        }
    }
    const int64_t off[3] = {banded ? -rs : -(rs + 1), -1, banded ? -rs + 1 : -rs};
    std::vector<Path> paths;
    for (const Start &s0 : starts) {
        enumerate_from(dir, affine, off, s0, max_number, max_number, paths);
        if ((int64_t)paths.size() >= max_number) break;
    }
    if ((int64_t)paths.size() > max_number) paths.resize((size_t)max_number);
    std::vector<int64_t> rows;
    for (const Path &path : paths) {
        path_to_trace(path, rs, banded, lower_diag, rows);
        list.traces.push_back(rows);
    }
}

// ---------------------------------------------------------------------------------------
// Single pair, low latency.  A call of align_optimal / align_banded on one pair is dominated by
// the trips around the kernel, not by the kernel: this path keeps one stream, one page-locked
// staging buffer each way and its device buffers per host thread, sends parameters, offsets,
// matrix and both sequences in ONE host-to-device copy, launches fill (+ traceback) and reads
// score, start and steps back in ONE device-to-host copy.  Codes are validated on the host (a
// few hundred bytes).  Returns 1 when the pair must take the general batch path instead (sums
// that could wrap, tables beyond the cached buffers' limit).
// ---------------------------------------------------------------------------------------
struct PairCtx {
    int device = -1;
    cudaStream_t stream = nullptr;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr, *d_dirs = nullptr, *h_dirs = nullptr;
    size_t h_in_cap = 0, h_out_cap = 0, d_in_cap = 0, d_out_cap = 0, d_dirs_cap = 0, h_dirs_cap = 0;
    DevBuf bnd;
    void release() {
        if (stream) cudaStreamDestroy(stream);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (h_dirs) cudaFreeHost(h_dirs);
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        if (d_dirs) cudaFree(d_dirs);
        bnd.release();
        cudaGetLastError();
        *this = PairCtx();
    }
    ~PairCtx() {
        // at thread / process exit the runtime may be gone already: errors are ignored
        if (device >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess) { cudaSetDevice(device); release(); if (cur >= 0) cudaSetDevice(cur); } }
        cudaGetLastError();
    }
    PairCtx() = default;
    PairCtx(const PairCtx &) = delete;
    PairCtx &operator=(PairCtx &&o) {
        device = o.device; stream = o.stream; h_in = o.h_in; h_out = o.h_out; d_in = o.d_in; d_out = o.d_out;
        d_dirs = o.d_dirs; h_dirs = o.h_dirs; h_in_cap = o.h_in_cap; h_out_cap = o.h_out_cap; d_in_cap = o.d_in_cap;
        d_out_cap = o.d_out_cap; d_dirs_cap = o.d_dirs_cap; h_dirs_cap = o.h_dirs_cap;
        return *this;
    }
};
static thread_local PairCtx g_pair_ctx;

template <bool HOST>
static cudaError_t pair_ensure(void **ptr, size_t *cap, size_t need) {
    if (*cap >= need) return cudaSuccess;
    if (*ptr) { if (HOST) cudaFreeHost(*ptr); else cudaFree(*ptr); }
    *ptr = nullptr; *cap = 0;
    const size_t want = std::max<size_t>(need + need / 2, 64 * 1024);
    cudaError_t e = HOST ? cudaHostAlloc(ptr, want, cudaHostAllocDefault) : cudaMalloc(ptr, want);
    if (e == cudaSuccess) *cap = want;
    return e;
}

static const int64_t kPairTableLimit = 4ll << 30;   // direction bytes the cached buffers may hold

static int align_pair_fast(const BtParams &P, const void *code1, int64_t len1, const void *code2, int64_t len2,
                           const int32_t *matrix, int64_t lower_diag, int64_t upper_diag, int32_t score_only,
                           int64_t max_number, int32_t *score_out, BtTraceList **traces_out) {
    if (len1 < 0 || len2 < 0 || len1 > INT32_MAX / 2 - 2 || len2 > INT32_MAX / 2 - 2) return 1;
    if (P.use_matrix && !matrix) return 1;
    const bool banded = P.banded != 0;
    const int64_t W = upper_diag - lower_diag + 1;
    if (banded && (W < 1 || lower_diag < -len1 || upper_diag > len2)) return 1;   // the batch path reports it
    // scoring range: the closed forms of the first row / column need sums that cannot wrap
    int32_t mn = 0, mx = 0;
    if (P.use_matrix) {
        const int64_t cnt = (int64_t)P.n_rows * P.n_cols;
        mn = mx = matrix[0];
        for (int64_t k = 1; k < cnt; k++) { mn = std::min(mn, matrix[k]); mx = std::max(mx, matrix[k]); }
    } else {
        mn = std::min(P.match_score, P.mismatch_score); mx = std::max(P.match_score, P.mismatch_score);
    }
    const int64_t step = std::max<int64_t>({std::abs((int64_t)mn), std::abs((int64_t)mx), std::abs((int64_t)P.gap_open),
                                            std::abs((int64_t)P.gap_ext)});
    if (step * (len1 + len2 + 4) >= (1ll << 30)) return 1;
    if (const char *forced = getenv("B200ALIGN_GENERAL")) { if (strcmp(forced, "antidiag") == 0) return 1; }
    const int64_t rs = table_stride(1, banded ? W + 2 : len2 + 1);
    const int64_t table = (len1 + 1) * rs;
    if (!score_only && table > kPairTableLimit) return 1;
    if (bt_device_count() <= 0) return fail(BT_ERR_CUDA, "no CUDA device available (libb200align has no CPU fallback)");
    // codes outside the matrix: the reference would panic (scoring.rs:63)
    const int cb = P.code_bytes;
    if (P.use_matrix) {
        auto bad = [cb](const void *c, int64_t n, uint64_t limit) {
            for (int64_t k = 0; k < n; k++) {
                uint64_t v;
                switch (cb) {
                case 1: v = ((const uint8_t *)c)[k]; break;
                case 2: v = ((const uint16_t *)c)[k]; break;
                case 4: v = ((const uint32_t *)c)[k]; break;
                default: v = ((const uint64_t *)c)[k]; break;
                }
                if (v >= limit) return true;
            }
            return false;
        };
        if (bad(code1, len1, (uint64_t)P.n_rows) || bad(code2, len2, (uint64_t)P.n_cols))
            return fail(BT_ERR_INVALID_CODE, "a sequence code is outside the substitution matrix");
    }

    PairCtx &C = g_pair_ctx;
    const int dev = bt_get_device();
    if (C.device != dev) {
        if (C.device >= 0) { DeviceGuard g(C.device); C.release(); }
        C.device = dev;
        CUDA_TRY(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
    }
    cudaStream_t s = C.stream;
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    // ---- input block: 8 int64 (offsets, band, table offsets) | matrix | codes1 | codes2
    const size_t mat_bytes = P.use_matrix ? (size_t)P.n_rows * P.n_cols * 4 : 0;
    const size_t in_mat = 64, in_c1 = in_mat + up16(mat_bytes), in_c2 = in_c1 + up16((size_t)len1 * cb);
    const size_t in_bytes = in_c2 + up16((size_t)len2 * cb);
    // ---- output block: results (64 bytes) | step bytes | end candidates of a band
    const size_t out_ops = 64, out_cand = out_ops + up16((size_t)(len1 + len2));
    const size_t out_bytes = out_cand + (banded ? (size_t)(3 * (W + 2)) * 4 : 0);
    CUDA_TRY(pair_ensure<true>(&C.h_in, &C.h_in_cap, in_bytes));
    CUDA_TRY(pair_ensure<false>(&C.d_in, &C.d_in_cap, in_bytes));
    CUDA_TRY(pair_ensure<true>(&C.h_out, &C.h_out_cap, out_bytes));
    CUDA_TRY(pair_ensure<false>(&C.d_out, &C.d_out_cap, out_bytes));
    if (!score_only) CUDA_TRY(pair_ensure<false>(&C.d_dirs, &C.d_dirs_cap, (size_t)table));
    {
        int64_t *hdr = (int64_t *)C.h_in;
        hdr[0] = 0; hdr[1] = len1; hdr[2] = 0; hdr[3] = len2; hdr[4] = lower_diag; hdr[5] = upper_diag; hdr[6] = 0; hdr[7] = 0;
        if (mat_bytes) memcpy((char *)C.h_in + in_mat, matrix, mat_bytes);
        if (len1) memcpy((char *)C.h_in + in_c1, code1, (size_t)len1 * cb);
        if (len2) memcpy((char *)C.h_in + in_c2, code2, (size_t)len2 * cb);
    }
    CUDA_TRY(cudaMemcpyAsync(C.d_in, C.h_in, in_bytes, cudaMemcpyHostToDevice, s));

    DevParams D{};
    D.affine = P.affine; D.local = P.local; D.terminal = banded ? 1 : P.terminal_penalty;
    D.use_matrix = P.use_matrix; D.banded = P.banded; D.score_only = score_only;
    D.gap_open = P.gap_open; D.gap_ext = P.affine ? P.gap_ext : P.gap_open;
    D.match = P.match_score; D.mismatch = P.mismatch_score;
    D.neg_inf = reference_neg_inf(P, mn);
    D.n_rows = P.n_rows; D.n_cols = P.n_cols; D.code_bytes = P.code_bytes;
    D.dir_pad = 1;
    {
        const char *forced = getenv("B200ALIGN_GENERAL");
        D.lean = (score_only || max_number == 1) && !(forced && strcmp(forced, "full") == 0) &&
                 strip_eligible(P, len1 + len2, mn, mx);
    }
    char *din = (char *)C.d_in, *dout = (char *)C.d_out;
    BatchView V{};
    V.codes1 = din + in_c1; V.codes2 = din + in_c2;
    V.map.off1 = (const int64_t *)din; V.map.off2 = (const int64_t *)din + 2;
    V.bands = banded ? (const int64_t *)din + 4 : nullptr;
    V.matrix = P.use_matrix ? (const int32_t *)(din + in_mat) : nullptr;
    V.dirs = (uint8_t *)C.d_dirs;
    V.dir_off = (const int64_t *)din + 6;
    V.cand = (int32_t *)(dout + out_cand);
    V.cand_off = (const int64_t *)din + 7;
    V.score = (int32_t *)dout; V.start_state = (int32_t *)(dout + 4); V.start_idx = (int64_t *)(dout + 8);
    V.end_vals = (int32_t *)(dout + 16); V.trace_len = (int64_t *)(dout + 32); V.first = (int32_t *)(dout + 40);
    V.ops = (uint8_t *)(dout + out_ops);
    V.pair_begin = 0; V.pair_end = 1;
    const bool first_only = max_number == 1;
    auto fill = [&]() -> int {
        if (!score_only) CUDA_TRY(cudaMemsetAsync(C.d_dirs, 0, (size_t)table, s));
        return wavefront_launch(P, D, V, 0, 1, (int)len1, (int)len2, C.bnd, s, nullptr);
    };
    int st = fill();
    if (st) return st;
    const char *hout = (const char *)C.h_out;
    if (score_only || first_only) {
        size_t back = 64;
        if (first_only && !score_only) {
            CUDA_TRY(dispatch_general_traceback(D, V, 1, s));
            back = out_ops + (size_t)(len1 + len2);
        }
        CUDA_TRY(cudaMemcpyAsync(C.h_out, C.d_out, back, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        *score_out = *(const int32_t *)hout;
        if (score_only) return BT_OK;
        std::unique_ptr<BtTraceList> list(new BtTraceList());
        const int64_t len = *(const int64_t *)(hout + 32);
        const int32_t *first = (const int32_t *)(hout + 40);
        std::vector<int64_t> rows((size_t)(2 * len));
        const int64_t nrows = expand_ops(len, first[0], first[1], (const uint8_t *)hout + out_ops, banded, lower_diag, rows.data());
        rows.resize((size_t)(2 * nrows));
        list->traces.push_back(std::move(rows));
        *traces_out = list.release();
        return BT_OK;
    }
    // ---- co-optimal enumeration: the whole direction table comes back
    int32_t score = 0;
    if (P.local) {
        // the optimum first, then a second fill that tags every cell holding it (pairwise.rs:536-551)
        CUDA_TRY(cudaMemcpyAsync(C.h_out, C.d_out, 64, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        score = *(const int32_t *)hout;
        D.mark = 1; D.mark_value = score;
        st = fill();
        if (st) return st;
    }
    CUDA_TRY(pair_ensure<true>(&C.h_dirs, &C.h_dirs_cap, (size_t)table));
    CUDA_TRY(cudaMemcpyAsync(C.h_dirs, C.d_dirs, (size_t)table, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(C.h_out, C.d_out, out_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    score = *(const int32_t *)hout;
    *score_out = score;
    std::unique_ptr<BtTraceList> list(new BtTraceList());
    enumerate_traces(P, (const uint8_t *)C.h_dirs, (const int32_t *)(hout + out_cand), (const int32_t *)(hout + 16), score,
                     len1, len2, lower_diag, upper_diag, rs, max_number, *list);
    *traces_out = list.release();
    return BT_OK;
}

extern "C" int bt_align_pair(const BtParams *params, const void *code1, int64_t len1, const void *code2,
                             int64_t len2, const int32_t *matrix, int64_t lower_diag, int64_t upper_diag,
                             int32_t score_only, int64_t max_number, int32_t *score_out,
                             BtTraceList **traces_out) {
    if (!score_out) return fail(BT_ERR_INVALID_ARG, "score_out is NULL");
    if (traces_out) *traces_out = nullptr;
    score_only = score_only ? 1 : 0;
    if (!score_only && !traces_out) return fail(BT_ERR_INVALID_ARG, "traces_out is NULL");
    if (max_number < 1) return fail(BT_ERR_INVALID_ARG, "Maximum number of returned alignments must be at least 1");
    {
        int stc = check_params(params);
        if (stc) return stc;
        const int fast = align_pair_fast(*params, code1, len1, code2, len2, matrix, lower_diag, upper_diag, score_only,
                                         max_number, score_out, traces_out);
        if (fast != 1) return fast;
    }
    const int64_t off1[2] = {0, len1}, off2[2] = {0, len2};
    const int64_t band[2] = {lower_diag, upper_diag};
    // the general batch path (anti-diagonal kernel where sums could wrap, chunked tables)
    BtParams P = *params;
    BtBatch *raw = nullptr;
    int st = batch_create_impl(&P, code1, off1, 1, code2, off2, 1, nullptr, nullptr, 1, matrix,
                               P.banded ? band : nullptr, score_only, true, &raw);
    if (st) return st;
    std::unique_ptr<BtBatch> b(raw);
    const bool first_only = (max_number == 1);
    b->phase_used = 0;
    st = run_general(b.get(), first_only);
    if (st) return st;
    b->ran = true;
    int32_t score = 0;
    CUDA_TRY(cudaMemcpyAsync(&score, b->score.ptr, 4, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    *score_out = score;
    if (score_only) return BT_OK;

    std::unique_ptr<BtTraceList> list(new BtTraceList());
    const bool banded = P.banded != 0;
    const int64_t W = upper_diag - lower_diag + 1;
    const int64_t rs = table_stride(b->D.dir_pad, banded ? W + 2 : len2 + 1);
    if (first_only) {
        int64_t len = 0; int32_t first[2] = {0, 0};
        CUDA_TRY(cudaMemcpyAsync(&len, b->trace_len.ptr, 8, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaMemcpyAsync(first, b->first.ptr, 8, cudaMemcpyDeviceToHost, b->stream));
        CUDA_TRY(cudaStreamSynchronize(b->stream));
        std::vector<uint8_t> ops((size_t)std::max<int64_t>(len, 1));
        if (len > 0) CUDA_TRY(cudaMemcpy(ops.data(), b->ops.ptr, (size_t)len, cudaMemcpyDeviceToHost));
        std::vector<int64_t> rows((size_t)(2 * len));
        int64_t nrows = expand_ops(len, first[0], first[1], ops.data(), banded, lower_diag, rows.data());
        rows.resize((size_t)(2 * nrows));
        list->traces.push_back(std::move(rows));
        *traces_out = list.release();
        return BT_OK;
    }

    // ---- co-optimal enumeration on the host over the direction table ------------------
    if (P.local) {
        // second pass tags every cell whose M equals the optimum (pairwise.rs:536-551)
        b->D.mark = 1; b->D.mark_value = score;
        b->phase_used = 0;
        st = run_general(b.get(), false);
        if (st) return st;
    }
    const int64_t table = (len1 + 1) * rs;
    std::vector<uint8_t> dir((size_t)table);
    CUDA_TRY(cudaMemcpyAsync(dir.data(), b->dirs.ptr, (size_t)table, cudaMemcpyDeviceToHost, b->stream));
    std::vector<int32_t> cand;
    int32_t endv[3] = {0, 0, 0};
    if (banded && !P.local) {
        cand.resize((size_t)(3 * (W + 2)));
        CUDA_TRY(cudaMemcpyAsync(cand.data(), b->cand.ptr, cand.size() * 4, cudaMemcpyDeviceToHost, b->stream));
    }
    CUDA_TRY(cudaMemcpyAsync(endv, b->end_vals.ptr, 12, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    enumerate_traces(P, dir.data(), cand.data(), endv, score, len1, len2, lower_diag, upper_diag, rs, max_number, *list);
    *traces_out = list.release();
    return BT_OK;
}

// ---------------------------------------------------------------------------------------
// roofline denominators
// ---------------------------------------------------------------------------------------
extern "C" int bt_microbench(int kind, double *lane_instr_per_s) {
    if (!lane_instr_per_s || kind < 0 || kind > 2) return fail(BT_ERR_INVALID_ARG, "bad microbench kind");
    if (bt_device_count() <= 0) return fail(BT_ERR_CUDA, "no CUDA device available");
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    DevBuf out;
    CUDA_TRY(out.alloc((size_t)threads * blocks * 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        if (kind == 0) microbench_kernel<0><<<blocks, threads>>>(out.as<uint32_t>(), iters, 12345u);
        else if (kind == 1) microbench_kernel<1><<<blocks, threads>>>(out.as<uint32_t>(), iters, 12345u);
        else microbench_kernel<2><<<blocks, threads>>>(out.as<uint32_t>(), iters, 12345u);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double instr = (double)threads * blocks * iters * microbench_instr_per_iter(kind);
        if (rep > 0) best = std::max(best, instr / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lane_instr_per_s = best;
    return BT_OK;
}

extern "C" void bt_pool_trim(void) { DevicePool::instance().trim(); }
