// Inter-sequence kernels for align_banded (reference: src/rust/sequence/align/banded.rs):
// the throughput path of BASELINE config 4 (10^5 nucleotide pairs of 10 kb, band +-128,
// affine, semi-global, score + trace).
//
// Mapping.  Pairs are sorted by length and dealt to warps in groups of 32; every lane owns
// ONE pair and keeps 32-bit scores (10 kb x match 5 leaves the 16-bit range).  The lane
// sweeps its band in stripes of R = 16 rows.  Inside a stripe the column coordinate is
//     c = seq_j - (r0 + lower)            (r0 = first row of the stripe, 0-based seq_i)
// so that row r of the stripe is in the band iff r <= c <= r + W - 1: the stripe visits
// c = 0 .. R + W - 2, a parallelogram of (W + R - 1) x R cells for W x R useful ones.
// State of the previous column (M, F1, H per row) lives in registers; the row above the
// stripe comes from a ring buffer indexed by r0 + c (only W + R entries are ever live) that
// holds two values per column: H of the stripe's last row and t = 2 F2' + flag, where
// F2' = max(M, F2 + ext) is already the vertical gap state of the row BELOW and flag says
// that M attained it (the direction bit of that row) - the producer does the next stripe's
// first vertical step, so 8 bytes per column cross a stripe seam instead of 12.  The entries
// of the next BD_STAGE columns are copied to shared memory with cp.async while the current
// column is computed: a ring entry is usually in HBM by the time it is needed again (the
// rings of all resident warps exceed L2), and a plain load one or two columns ahead left
// that latency exposed (the top stall site of the round-1 profile).
// Substitution scores come from the per-lane shared-memory table of the packed kernels.
//
// Band semantics of the reference (banded.rs:184-214): cells outside the band are sentinels
// (all states -inf, :474-484); cells inside the band but outside the sequences - row 0 and
// seq_j < 0 - keep the edge value (M = 0, gaps -inf, :462-472) and act as free starts.
// Columns where every row of every lane is an ordinary cell run an unmasked fast path;
// the ramps at both ends of a stripe run a masked path that forces these values.
//
// Recurrence and predicates are those of interseq_kernel.cuh (F = G - open, H = max3); the
// optimum is taken over the end candidates of banded.rs:598-638 by the traceback kernel,
// which needs H of the last row and of the last column (captured here) and the nibbles.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/b200align.h"
#include "common.cuh"
#include "pool.cuh"

namespace bt {

constexpr int BD_R = 16;
constexpr int BD_THREADS = 128;
constexpr int BD_MAX_LEN = 65535;
constexpr int BD_MAX_W = 4096;
constexpr int32_t BD_NEG = -(1 << 29);
constexpr int BD_STAGE = 8;          // columns of ring look-ahead staged in shared memory

__device__ __forceinline__ void bd_cp_async8(void *smem_dst, const void *gmem_src) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void bd_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bd_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct BandDesc {          // one group of 32 pairs = one warp
    int64_t dir_off;       // uint32 words
    int64_t ring_off;      // int2 entries ([slot][lane])
    int64_t rowcode_off;   // bytes
    int64_t colcode_off;   // bytes
    int64_t lastrow_off;   // int32 entries, [seq_j][lane]
    int64_t lastcol_off;   // int32 entries, [row][lane]
    int32_t nmax, mmax, stripes, cw, ring, wmax;
};

struct BandConst {
    int32_t gap_open, gap_ext, n_rows, n_cols, score_only, pad;
};

struct BandArgs {
    const BandDesc *desc;
    const int32_t *order;
    PairMap map;
    const int64_t *bands;
    const int32_t *matrix;
    const uint8_t *rowcodes, *colcodes;
    int32_t *ring, *lastrow, *lastcol;
    uint32_t *dirs;
    int32_t *score;
    int64_t n_groups;
};

// ---------------------------------------------------------------------------------------
// packer: lane-interleaved byte codes of each group
// ---------------------------------------------------------------------------------------
__global__ void banded_pack_kernel(const uint8_t *codes1, const uint8_t *codes2, const PairMap map,
                                   const int32_t *order, const BandDesc *desc,
                                   uint8_t *rowcodes, uint8_t *colcodes, int64_t n_groups) {
    const int64_t g = blockIdx.x;
    if (g >= n_groups) return;
    const BandDesc d = desc[g];
    const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5, nsub = blockDim.x >> 5;
    const int32_t p = order[g * 32 + lane];
    int64_t a = 0, b = 0;
    int n = 0, m = 0;
    if (p >= 0) { a = map.beg1(p); n = (int)map.len1(p); b = map.beg2(p); m = (int)map.len2(p); }
    const int rows = d.stripes * BD_R;
    for (int i = sub; i < rows; i += nsub) rowcodes[d.rowcode_off + (int64_t)i * 32 + lane] = i < n ? codes1[a + i] : 0;
    for (int j = sub; j < d.mmax; j += nsub) colcodes[d.colcode_off + (int64_t)j * 32 + lane] = j < m ? codes2[b + j] : 0;
}

#ifndef BD_PF
#define BD_PF 2
#endif
#ifndef BD_MIN_BLOCKS
#define BD_MIN_BLOCKS 3      // 168 registers: the unrolled look-ahead does not fit 128 without spills
#endif
// max with one predicate: adds `c` to `acc` where the FIRST operand attains the maximum
#define BD_MAX_PRED(out, a, b, acc, c)                                                    \
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, %3;\n\tmax.s32 %0, %2, %3;\n\t"         \
        "@p add.u32 %1, %1, %4;\n\t}"                                                     \
        : "=&r"(out), "+r"(acc)                                                            \
        : "r"(a), "r"(b), "r"(c))

template <bool TRACED>
__global__ void __launch_bounds__(BD_THREADS, BD_MIN_BLOCKS)
banded_fill_kernel(const BandConst C, const BandArgs A) {
    extern __shared__ int32_t btable[];  // [b][a][lane], one private copy per lane (no bank conflicts)
    const int lane = threadIdx.x & 31;
    {
        const int entries = C.n_rows * C.n_cols;
        for (int e = threadIdx.x >> 5; e < entries; e += BD_THREADS / 32) {
            const int a = e / C.n_cols, b = e % C.n_cols;
            btable[(b * C.n_rows + a) * 32 + lane] = A.matrix[e];
        }
    }
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * (BD_THREADS / 32) + (threadIdx.x >> 5);
    if (g >= A.n_groups) return;
    const BandDesc d = A.desc[g];
    const int32_t p = A.order[g * 32 + lane];
    // a padding lane behaves like a 1 x 1 pair with band (0, 0)
    int n = 1, m = 1, lower = 0, W = 1;
    if (p >= 0) {
        n = (int)A.map.len1(p); m = (int)A.map.len2(p);
        lower = (int)A.bands[2 * p]; W = (int)(A.bands[2 * p + 1] - A.bands[2 * p]) + 1;
    }
    const int32_t NEG = BD_NEG, open = C.gap_open, ext = C.gap_ext;
    const uint32_t col_stride = (uint32_t)C.n_rows * 128u;
    const char *tbase = (const char *)btable + lane * 4;
    const uint8_t *rc = A.rowcodes + d.rowcode_off + lane;
    const uint8_t *cc = A.colcodes + d.colcode_off + lane;
    const int ring = d.ring;
    int2 *const rg = (int2 *)A.ring + d.ring_off + lane;   // [slot][lane]: {H, t} of the stripe's last row
    // this lane's staging slots: [column & (BD_STAGE - 1)][lane] behind the score table
    int2 *const stage = (int2 *)(btable + C.n_rows * C.n_cols * 32) + (threadIdx.x >> 5) * (BD_STAGE * 32) + lane;
    const int32_t T_NEG = 2 * BD_NEG + 1;                  // t of a sentinel above: F2' = -inf, "M attains it"
    uint2 *dir = TRACED ? (uint2 *)(A.dirs + d.dir_off) + lane : nullptr;
    int32_t *lastrow = A.lastrow + d.lastrow_off + lane;
    int32_t *lastcol = A.lastcol + d.lastcol_off + lane;
    const int s_last = (n - 1) / BD_R, r_last = (n - 1) % BD_R;

    for (int s = 0; s < d.stripes; s++) {
        const int r0 = s * BD_R;
        const int base_j = r0 + lower;                      // seq_j of column c = 0
        // columns of this stripe that lie inside the lane's sequence 2, and the warp's union
        const int c_lo = max(0, -base_j), c_hi = min(BD_R + W - 2, m - 1 - base_j);
        const bool lane_has_rows = r0 < n;
        int c_first = __reduce_min_sync(0xffffffffu, lane_has_rows ? c_lo : 0x7fffffff);
        int c_last = __reduce_max_sync(0xffffffffu, lane_has_rows ? c_hi : -1);
        if (c_last < c_first) continue;
        // unmasked range: every row of every lane is an ordinary cell with ordinary neighbours
        const int f_lo = __reduce_max_sync(0xffffffffu, max(c_lo, BD_R - 1) + 1);   // +1: left neighbours ordinary too
        const int f_hi = __reduce_min_sync(0xffffffffu, min(c_hi, W - 1));
        const int c_m = m - 1 - base_j;                      // column holding seq_j = m - 1

        int32_t rp[BD_R], Ml[BD_R], F1l[BD_R], Hl[BD_R];
#pragma unroll
        for (int r = 0; r < BD_R; r++) {
            rp[r] = (int32_t)rc[(int64_t)(r0 + r) * 32] * 128;
            // state of column c_first - 1: left of every lane's sequence -> edge inside the band,
            // sentinel outside (banded.rs:184-190)
            const int cb = c_first - 1;
            const bool inb = cb >= r && cb - r <= W - 1;
            Ml[r] = inb ? 0 : NEG; F1l[r] = NEG; Hl[r] = inb ? 0 : NEG;
        }
        // row above the stripe at column c: in the band iff -1 <= c <= W - 2; inside the band it is
        // an edge cell (M = H = 0, gaps -inf: F2' = 0 opened from M) when it is table row 0 or lies
        // left of the sequence (banded.rs:184), else the previous stripe's last row (ring).
        // Out-of-band columns may never have been visited by the previous stripe, so they are
        // synthesised, not read.  Returns true when (uh, ut) are synthesised.
        auto synth = [&](int c, int32_t &uh, int32_t &ut) -> bool {
            const bool inb = c >= -1 && c <= W - 2;
            if (!inb) { uh = NEG; ut = T_NEG; return true; }
            if (s == 0 || base_j + c < 0) { uh = 0; ut = 1; return true; }
            return false;
        };
        int32_t hu_prev;
        {
            int32_t t0;
            if (!synth(c_first - 1, hu_prev, t0)) hu_prev = rg[(int64_t)((r0 + c_first - 1) % ring) * 32].x;
        }
        auto col_code = [&](int c) -> uint32_t {
            const int sjc = min(max(base_j + c, 0), d.mmax - 1);
            return cc[(int64_t)sjc * 32];
        };
        // ring slots of the column being written and of the column being staged
        int w_slot = (r0 + c_first) % ring, pf_slot = w_slot, pf_c = c_first;
        auto stage_next = [&]() {       // asynchronous copy of the ring entry of column pf_c; one group per column
            int32_t a, b;
            if (pf_c <= c_last && !synth(pf_c, a, b)) bd_cp_async8(stage + (pf_c & (BD_STAGE - 1)) * 32, rg + (int64_t)pf_slot * 32);
            bd_cp_async_commit();
            pf_c++;
            if (++pf_slot == ring) pf_slot = 0;
        };
#pragma unroll
        for (int k = 0; k < BD_STAGE - 1; k++) stage_next();
        // the column letters are fetched BD_PF columns ahead
        uint32_t q_code[BD_PF];
#pragma unroll
        for (int k = 0; k < BD_PF; k++) q_code[k] = c_first + k <= c_last ? col_code(c_first + k) : 0;
        for (int c0 = c_first; c0 <= c_last; c0 += BD_PF) {
#pragma unroll
        for (int u = 0; u < BD_PF; u++) {
            const int c = c0 + u;
            if (u > 0 && c > c_last) break;
            stage_next();                               // column c + BD_STAGE - 1
            bd_cp_async_wait<BD_STAGE - 1>();           // the group of column c has landed
            int32_t hu, ut;
            if (!synth(c, hu, ut)) { const int2 v = stage[(c & (BD_STAGE - 1)) * 32]; hu = v.x; ut = v.y; }
            int32_t up_m = 0, up_f2 = ut >> 1;          // F2 of the stripe's first row: done by the row above
            const uint32_t code = q_code[u];
            if (c + BD_PF <= c_last) q_code[u] = col_code(c + BD_PF);
            const int sj = base_j + c;
            const char *t_col = tbase + code * col_stride;
            uint32_t acc0 = 0, acc1 = 0;   // nibbles of rows 0-7 and 8-15
            int32_t sim[BD_R];
#pragma unroll
            for (int r = 0; r < BD_R; r++) sim[r] = *(const int32_t *)(t_col + rp[r]);
            const bool fast = c >= f_lo && c <= f_hi;
            if (fast) {
#define BD_BIT(r, b) ((1u << (b)) << (4 * ((r) & 7)))
                // identical to the packed kernel: in-place, rows software-pipelined
                {
                    const int32_t fe = F1l[0] + ext;
                    if (TRACED) { BD_MAX_PRED(F1l[0], Ml[0], fe, acc0, BD_BIT(0, 2)); } else F1l[0] = max(Ml[0], fe);
                    Ml[0] = hu_prev + sim[0];
                }
#pragma unroll
                for (int r = 0; r < BD_R; r++) {
                    if (r + 1 < BD_R) {
                        const int32_t fe = F1l[r + 1] + ext;
                        if (TRACED) {
                            if (r + 1 < 8) { BD_MAX_PRED(F1l[r + 1], Ml[r + 1], fe, acc0, BD_BIT(r + 1, 2)); }
                            else { BD_MAX_PRED(F1l[r + 1], Ml[r + 1], fe, acc1, BD_BIT(r + 1, 2)); }
                        } else F1l[r + 1] = max(Ml[r + 1], fe);
                        Ml[r + 1] = Hl[r] + sim[r + 1];
                    }
                    const int32_t fv = up_f2 + ext;
                    if (TRACED) {
                        int32_t t;
                        if (r == 0) {
                            acc0 += (uint32_t)(ut & 1) << 3;
                            BD_MAX_PRED(t, F1l[r], up_f2, acc0, BD_BIT(r, 1));
                            const int32_t to = t + open;
                            BD_MAX_PRED(Hl[r], Ml[r], to, acc0, BD_BIT(r, 0));
                        } else if (r < 8) {
                            BD_MAX_PRED(up_f2, up_m, fv, acc0, BD_BIT(r, 3));
                            BD_MAX_PRED(t, F1l[r], up_f2, acc0, BD_BIT(r, 1));
                            const int32_t to = t + open;
                            BD_MAX_PRED(Hl[r], Ml[r], to, acc0, BD_BIT(r, 0));
                        } else {
                            BD_MAX_PRED(up_f2, up_m, fv, acc1, BD_BIT(r, 3));
                            BD_MAX_PRED(t, F1l[r], up_f2, acc1, BD_BIT(r, 1));
                            const int32_t to = t + open;
                            BD_MAX_PRED(Hl[r], Ml[r], to, acc1, BD_BIT(r, 0));
                        }
                    } else {
                        if (r > 0) up_f2 = max(up_m, fv);
                        Hl[r] = __viaddmax_s32(max(F1l[r], up_f2), open, Ml[r]);
                    }
                    up_m = Ml[r];
                }
            } else {
                // masked path: compute, then force sentinel / edge values where the cell is not an
                // ordinary one; its stored state is what the neighbouring cells must read
                int32_t diag = hu_prev;
#pragma unroll
                for (int r = 0; r < BD_R; r++) {
                    const int32_t fe = F1l[r] + ext, fv = up_f2 + ext;
                    int32_t f1, f2, t, h;
                    int32_t mcell = diag + sim[r];
                    uint32_t bits = 0;
                    BD_MAX_PRED(f1, Ml[r], fe, bits, 4u);
                    if (r == 0) { f2 = up_f2; bits += (uint32_t)(ut & 1) << 3; }
                    else { BD_MAX_PRED(f2, up_m, fv, bits, 8u); }
                    BD_MAX_PRED(t, f1, f2, bits, 2u);
                    const int32_t to = t + open;
                    BD_MAX_PRED(h, mcell, to, bits, 1u);
                    const bool inb = c >= r && c - r <= W - 1;
                    if (!inb) { mcell = NEG; f1 = NEG; f2 = NEG; h = NEG; }
                    else if (sj < 0) { mcell = 0; f1 = NEG; f2 = NEG; h = 0; }
                    diag = Hl[r];
                    Hl[r] = h; Ml[r] = mcell; F1l[r] = f1;
                    up_m = mcell; up_f2 = f2;
                    if (TRACED) { if (r < 8) acc0 |= bits << (4 * r); else acc1 |= bits << (4 * (r - 8)); }
                }
            }
            hu_prev = hu;
            // the stripe's last row feeds the next stripe (forced values included)
            {
                const int32_t fvn = up_f2 + ext;
                const int32_t tn = 2 * max(up_m, fvn) + (up_m >= fvn ? 1 : 0);
                rg[(int64_t)w_slot * 32] = make_int2(Hl[BD_R - 1], tn);
                if (++w_slot == ring) w_slot = 0;
            }
            if (TRACED) dir[((int64_t)s * d.cw + c) * 32] = make_uint2(acc0, acc1);
            // end candidates (banded.rs:598-638): H along the last row and the last column
            if (s == s_last && sj >= 0 && sj < m) {
                int32_t h = 0;
#pragma unroll
                for (int r = 0; r < BD_R; r++) if (r == r_last) h = Hl[r];
                lastrow[(int64_t)sj * 32] = h;
            }
            if (c == c_m) {
#pragma unroll
                for (int r = 0; r < BD_R; r++) lastcol[(int64_t)(r0 + r) * 32] = Hl[r];
            }
        }
        }
    }
}

// ---------------------------------------------------------------------------------------
// optimum over the end candidates + first-priority traceback, one thread per pair
// ---------------------------------------------------------------------------------------
__global__ void banded_traceback_kernel(const BandConst C, const BandArgs A, int64_t *__restrict__ trace_len,
                                        int32_t *__restrict__ first, uint8_t *__restrict__ ops, int64_t n_slots) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const int32_t p = A.order[slot];
    if (p < 0) return;
    const int64_t g = slot >> 5;
    const int lane = (int)(slot & 31);
    const BandDesc d = A.desc[g];
    const int n = (int)A.map.len1(p), m = (int)A.map.len2(p);
    const int lower = (int)A.bands[2 * p], W = (int)(A.bands[2 * p + 1] - A.bands[2 * p]) + 1;
    const bool traced = !C.score_only;
    const uint32_t *base = A.dirs + d.dir_off + lane * 2;
    const int32_t *lastrow = A.lastrow + d.lastrow_off + lane;
    const int32_t *lastcol = A.lastcol + d.lastcol_off + lane;
    auto nibble = [&](int i, int sj) -> uint32_t {   // table row i >= 1, sequence column sj
        const int rg = i - 1, s = rg / BD_R, r = rg % BD_R;
        const int c = sj - (s * BD_R + lower);
        const uint32_t w = base[((int64_t)s * d.cw + c) * 64 + (r >> 3)];
        return (w >> (4 * (r & 7))) & 0xfu;
    };
    // candidate of band column jb (banded.rs:609-636)
    auto cand = [&](int jb, int &i, int &sj) {
        sj = jb + (n - 1) + lower - 1;
        if (sj < m) i = n;
        else { i = m - jb - lower + 1; sj = m - 1; }
    };
    auto cand_h = [&](int i, int sj) -> int32_t { return i == n ? lastrow[(int64_t)sj * 32] : lastcol[(int64_t)(i - 1) * 32]; };
    int32_t best = INT32_MIN;
    for (int jb = 1; jb <= W; jb++) { int i, sj; cand(jb, i, sj); best = max(best, cand_h(i, sj)); }
    A.score[p] = best;
    if (!traced) return;
    // start: states in the order Match, Gap1, Gap2, band columns ascending (banded.rs:582-590)
    int si = 0, ssj = 0, st = -1;
    for (int want = 0; want < 3 && st < 0; want++) {
        for (int jb = 1; jb <= W; jb++) {
            int i, sj; cand(jb, i, sj);
            if (cand_h(i, sj) != best) continue;
            const uint32_t nb = nibble(i, sj);
            const int top = (nb & 1u) ? ST_MATCH : ((nb & 2u) ? ST_GAP1 : ST_GAP2);
            // the state `want` attains the optimum iff it is the first-priority state of H, or
            // (for the gap states) ties with it: hM = 0 means H is a gap state; Gap1 ties Gap2
            // only when hG = 1 picked Gap1, so Gap2 is a start only if it is the top state
            bool hit = top == want;
            if (hit) { si = i; ssj = sj; st = want; break; }
        }
    }
    uint8_t *out = ops + A.map.ops(p);
    int i = si, sj = ssj, len = 0, fi = 0, fsj = 0;
    // step bytes go out four at a time once the position is word-aligned (see the packed kernels)
    uint32_t ew = 0;
    int enb = 0;
    const int ealign = (int)((uintptr_t)out & 3u);
    auto emit = [&](uint32_t op) {
        if (enb == 0 && ((ealign + len) & 3) != 0) out[len] = (uint8_t)op;
        else {
            ew |= op << (8 * enb);
            if (++enb == 4) { *(uint32_t *)(out + len - 3) = ew; ew = 0; enb = 0; }
        }
        len++;
    };
    while (i >= 1 && sj >= 0) {
        const int jb = sj - (i - 1) - lower + 1;
        if (jb < 1 || jb > W) break;
        if (st == ST_MATCH) {
            // A diagonal run stays on band column jb and visits (i-1, sj-1), (i-2, sj-2), ...:
            // their direction words are fetched together, so that a 10 kb trace pays one memory
            // latency per BD_AHEAD steps instead of one per step.
            constexpr int BD_AHEAD = 8;
            uint32_t ahead[BD_AHEAD];
#pragma unroll
            for (int t = 0; t < BD_AHEAD; t++) {
                const int ii = i - 1 - t, jj = sj - 1 - t;
                ahead[t] = 0;
                if (ii >= 1 && jj >= 0) {
                    const int rg = ii - 1, s = rg / BD_R;
                    ahead[t] = base[((int64_t)s * d.cw + (jj - (s * BD_R + lower))) * 64 + ((rg % BD_R) >> 3)];
                }
            }
#pragma unroll
            for (int t = 0; t < BD_AHEAD; t++) {
                fi = i; fsj = sj;
                emit(OP_DIAG); i--; sj--;
                if (!(i >= 1 && sj >= 0)) break;
                const uint32_t pn = (ahead[t] >> (4 * ((i - 1) & 7))) & 0xfu;
                st = (pn & 1u) ? ST_MATCH : ((pn & 2u) ? ST_GAP1 : ST_GAP2);
                if (st != ST_MATCH) break;
            }
        } else if (st == ST_GAP1) {
            const uint32_t nb = nibble(i, sj);
            fi = i; fsj = sj;
            emit(OP_TO1); sj--;
            st = (nb & 4u) ? ST_MATCH : ST_GAP1;
        } else {
            const uint32_t nb = nibble(i, sj);
            fi = i; fsj = sj;
            emit(OP_TO2); i--;
            st = (nb & 8u) ? ST_MATCH : ST_GAP2;
        }
    }
    for (int k = 0; k < enb; k++) out[len - enb + k] = (uint8_t)(ew >> (8 * k));
    trace_len[p] = len;
    first[2 * p] = fi;
    first[2 * p + 1] = len ? fsj - (fi - 1) - lower + 1 : 0;   // band column, as bt_expand_traces expects
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct BandWindow {
    int64_t group_begin, group_end;
    int64_t dir_words, ring_entries, rowcodes, colcodes, lastrow, lastcol;
};

struct BandedPlan {
    int64_t n_pairs = 0, n_groups = 0;
    BandConst C{};
    std::vector<int32_t> order;
    std::vector<BandDesc> desc;
    std::vector<BandWindow> windows;
    DevBuf d_order, d_desc, rowcodes, colcodes, dirs, ring, lastrow, lastcol;
};

struct BandedIO {
    const void *codes1, *codes2;
    PairMap map;
    const int64_t *bands;
    const int32_t *matrix;
    int32_t *score;
    int64_t *trace_len;
    int32_t *first;
    uint8_t *ops;
};

inline bool banded_eligible(const BtParams &P, const PairMap &hm,
                            const std::vector<int64_t> &bands, int64_t n_pairs, int32_t min_score,
                            int32_t max_score) {
    if (!P.banded || !P.affine || P.local || !P.use_matrix || P.code_bytes != 1) return false;
    if ((int64_t)P.n_rows * P.n_cols * 128 > 96 * 1024 || P.n_rows > 255 || P.n_cols > 255) return false;
    const char *forced = getenv("B200ALIGN_ROUTE");
    if (forced && strcmp(forced, "general") == 0) return false;
    double total = 0.0, max_pair = 0.0, max_waves = 0.0;
    int64_t span = 0;
    for (int64_t k = 0; k < n_pairs; k++) {
        const int64_t n = hm.len1(k), m = hm.len2(k);
        const int64_t w = bands[2 * k + 1] - bands[2 * k] + 1;
        if (n < 1 || m < 1 || n > BD_MAX_LEN || m > BD_MAX_LEN || w < 1 || w > BD_MAX_W) return false;
        if (bands[2 * k] < -(n - 1) || bands[2 * k + 1] > m - 1) return false;  // must be cropped
        total += (double)n * w; max_pair = std::max(max_pair, (double)n * w);
        max_waves = std::max(max_waves, (double)(2 * n + w));
        span = std::max(span, n + m);
    }
    // exact 32-bit range: no finite value may approach the -2^29 sentinel
    const int64_t step = std::max<int64_t>({std::abs((int64_t)min_score), std::abs((int64_t)max_score),
                                            std::abs((int64_t)P.gap_open), std::abs((int64_t)P.gap_ext)});
    if (step * (span + 4) >= (1 << 28)) return false;
    if (!(forced && strcmp(forced, "packed") == 0)) {
        // cost model as in interseq_eligible; measured on B200: ~8e8 cells/s for a lone warp,
        // ~5e11 cells/s for a full grid; general kernel ~7e10 cells/s on a full grid and
        // ~1.7 us per wavefront
        const double t_inter = std::max(32.0 * max_pair / 8.0e8, total / 5.0e11);
        const double t_general = std::max(total / 7.0e10, max_waves * 1.7e-6 * std::ceil((double)n_pairs / 1700.0));
        if (t_general < t_inter) return false;
    }
    return true;
}

inline int banded_prepare(BandedPlan &plan, const BtParams &P, const PairMap &hm,
                          const std::vector<int64_t> &bands, int64_t n_pairs,
                          int score_only, int64_t *cells_out, cudaStream_t stream) {
    plan.n_pairs = n_pairs;
    plan.C.gap_open = P.gap_open; plan.C.gap_ext = P.gap_ext;
    plan.C.n_rows = P.n_rows; plan.C.n_cols = P.n_cols; plan.C.score_only = score_only;
    // sort by length of the first sequence (rows), then of the second: std::sort on packed keys
    std::vector<uint64_t> keys((size_t)n_pairs);
    for (int64_t k = 0; k < n_pairs; k++)
        keys[(size_t)k] = ((uint64_t)hm.len1(k) << 48) | ((uint64_t)hm.len2(k) << 32) | (uint64_t)k;
    std::sort(keys.begin(), keys.end(), std::greater<uint64_t>());   // longest first
    plan.n_groups = (n_pairs + 31) / 32;
    plan.order.assign((size_t)plan.n_groups * 32, -1);
    for (int64_t q = 0; q < n_pairs; q++) plan.order[(size_t)q] = (int32_t)(keys[(size_t)q] & 0xffffffffu);
    plan.desc.resize((size_t)plan.n_groups);
    plan.windows.clear();
    const int64_t dir_budget_words = (int64_t)(88ll << 30) / 4;   // direction nibbles resident per window
    // A warp carries its 32 pairs through all stripes, so a window takes one "wave time" per
    // started wave of resident warps: cut the batch into the smallest number of windows that
    // fit the resident warps, equally large (3125 groups on 1776 slots: 2 x 1563, not 1776 + 1349
    // or - with the budget alone - 2000 + 1125, which needs three wave times).
    int64_t groups_per_window = plan.n_groups;
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t slots = (int64_t)sms * BD_MIN_BLOCKS * (BD_THREADS / 32);
        const int64_t n_win = std::max<int64_t>(1, (plan.n_groups + slots - 1) / slots);
        groups_per_window = (plan.n_groups + n_win - 1) / n_win;
    }
    BandWindow win{};
    // in-band, in-sequence cells of the batch (banded.rs:199-207), counted on a few threads
    int64_t cells = 0;
    {
        const int n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(interseq_host_threads(), 16u), n_pairs / 256));
        std::vector<int64_t> part((size_t)n_threads, 0);
        interseq_parallel_chunks(n_pairs, n_threads, [&](int t, int64_t b, int64_t e) {
            int64_t acc = 0;
            for (int64_t k = b; k < e; k++) {
                const int64_t n = hm.len1(k), m = hm.len2(k), lo = bands[2 * k], up = bands[2 * k + 1];
                for (int64_t i = 0; i < n; i++) {
                    const int64_t j0 = std::max<int64_t>(i + lo, 0), j1 = std::min<int64_t>(i + up + 1, m);
                    if (j1 > j0) acc += j1 - j0;
                }
            }
            part[(size_t)t] = acc;
        });
        for (int64_t v : part) cells += v;
    }
    int64_t max_dir = 0, max_ring = 0, max_rowc = 0, max_colc = 0, max_lr = 0, max_lc = 0;
    auto close = [&](int64_t g_end) {
        win.group_end = g_end;
        plan.windows.push_back(win);
        max_dir = std::max(max_dir, win.dir_words); max_ring = std::max(max_ring, win.ring_entries);
        max_rowc = std::max(max_rowc, win.rowcodes); max_colc = std::max(max_colc, win.colcodes);
        max_lr = std::max(max_lr, win.lastrow); max_lc = std::max(max_lc, win.lastcol);
        win = BandWindow{};
        win.group_begin = g_end;
    };
    for (int64_t g = 0; g < plan.n_groups; g++) {
        int nmax = 1, mmax = 1, wmax = 1;
        for (int q = 0; q < 32; q++) {
            const int32_t k = plan.order[(size_t)(g * 32 + q)];
            if (k < 0) continue;
            const int64_t n = hm.len1(k), m = hm.len2(k);
            const int64_t lo = bands[2 * k], up = bands[2 * k + 1];
            nmax = std::max<int>(nmax, (int)n); mmax = std::max<int>(mmax, (int)m);
            wmax = std::max<int>(wmax, (int)(up - lo + 1));
        }
        BandDesc &d = plan.desc[(size_t)g];
        d.nmax = nmax; d.mmax = mmax; d.wmax = wmax;
        d.stripes = (nmax + BD_R - 1) / BD_R;
        d.cw = BD_R + wmax - 1;
        const int ring = (BD_R + wmax + 2 + 7) / 8 * 8;   // live entries of a stripe seam, no power of two needed
        d.ring = ring;
        const int64_t dir_words = score_only ? 0 : (int64_t)d.stripes * d.cw * 64;
        if (win.group_begin < g && (win.dir_words + dir_words > dir_budget_words || g - win.group_begin >= groups_per_window)) close(g);
        d.dir_off = win.dir_words; d.ring_off = win.ring_entries;
        d.rowcode_off = win.rowcodes; d.colcode_off = win.colcodes;
        d.lastrow_off = win.lastrow; d.lastcol_off = win.lastcol;
        win.dir_words += dir_words;
        win.ring_entries += (int64_t)ring * 32;
        win.rowcodes += (int64_t)d.stripes * BD_R * 32;
        win.colcodes += (int64_t)mmax * 32;
        win.lastrow += (int64_t)mmax * 32;
        win.lastcol += (int64_t)d.stripes * BD_R * 32;
    }
    close(plan.n_groups);
    if (cells_out) *cells_out = cells;
    cudaError_t e;
    if ((e = plan.d_order.alloc(plan.order.size() * 4)) != cudaSuccess) goto fail;
    if ((e = plan.d_desc.alloc(plan.desc.size() * sizeof(BandDesc))) != cudaSuccess) goto fail;
    if ((e = plan.rowcodes.alloc((size_t)max_rowc)) != cudaSuccess) goto fail;
    if ((e = plan.colcodes.alloc((size_t)max_colc)) != cudaSuccess) goto fail;
    if ((e = plan.ring.alloc((size_t)max_ring * 8)) != cudaSuccess) goto fail;
    if ((e = plan.dirs.alloc((size_t)max_dir * 4)) != cudaSuccess) goto fail;
    if ((e = plan.lastrow.alloc((size_t)max_lr * 4)) != cudaSuccess) goto fail;
    if ((e = plan.lastcol.alloc((size_t)max_lc * 4)) != cudaSuccess) goto fail;
    if ((e = cudaMemcpyAsync(plan.d_order.ptr, plan.order.data(), plan.order.size() * 4, cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto fail;
    if ((e = cudaMemcpyAsync(plan.d_desc.ptr, plan.desc.data(), plan.desc.size() * sizeof(BandDesc), cudaMemcpyHostToDevice, stream)) != cudaSuccess) goto fail;
    return BT_OK;
fail:
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? BT_ERR_OOM : BT_ERR_CUDA;
}

inline BandArgs banded_args(BandedPlan &plan, const BandWindow &win, const BandedIO &io) {
    BandArgs A{};
    A.desc = plan.d_desc.as<BandDesc>() + win.group_begin;
    A.order = plan.d_order.as<int32_t>() + win.group_begin * 32;
    A.map = io.map; A.bands = io.bands; A.matrix = io.matrix;
    A.rowcodes = plan.rowcodes.as<uint8_t>(); A.colcodes = plan.colcodes.as<uint8_t>();
    A.ring = plan.ring.as<int32_t>(); A.lastrow = plan.lastrow.as<int32_t>(); A.lastcol = plan.lastcol.as<int32_t>();
    A.dirs = plan.dirs.as<uint32_t>();
    A.score = io.score;
    A.n_groups = win.group_end - win.group_begin;
    return A;
}

inline cudaError_t banded_fill_window(BandedPlan &plan, const BandWindow &win, const BandedIO &io,
                                      cudaStream_t stream, int64_t *launches) {
    const BandArgs A = banded_args(plan, win, io);
    banded_pack_kernel<<<(unsigned)A.n_groups, 256, 0, stream>>>(
        (const uint8_t *)io.codes1, (const uint8_t *)io.codes2, io.map, A.order, A.desc,
        plan.rowcodes.as<uint8_t>(), plan.colcodes.as<uint8_t>(), A.n_groups);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    (*launches)++;
    // score table (one copy per lane) + the ring staging slots of the block's warps
    const size_t smem = (size_t)plan.C.n_rows * plan.C.n_cols * 128 + (size_t)(BD_THREADS / 32) * BD_STAGE * 32 * 8;
    const unsigned grid = (unsigned)((A.n_groups + BD_THREADS / 32 - 1) / (BD_THREADS / 32));
    if (plan.C.score_only) {
        e = cudaFuncSetAttribute(banded_fill_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        banded_fill_kernel<false><<<grid, BD_THREADS, smem, stream>>>(plan.C, A);
    } else {
        e = cudaFuncSetAttribute(banded_fill_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        banded_fill_kernel<true><<<grid, BD_THREADS, smem, stream>>>(plan.C, A);
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) (*launches)++;
    return e;
}

inline cudaError_t banded_traceback_window(BandedPlan &plan, const BandWindow &win, const BandedIO &io,
                                           cudaStream_t stream, int64_t *launches) {
    const BandArgs A = banded_args(plan, win, io);
    const int64_t slots = A.n_groups * 32;
    banded_traceback_kernel<<<(unsigned)((slots + 127) / 128), 128, 0, stream>>>(plan.C, A, io.trace_len, io.first,
                                                                               io.ops, slots);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) (*launches)++;
    return e;
}

}  // namespace bt
