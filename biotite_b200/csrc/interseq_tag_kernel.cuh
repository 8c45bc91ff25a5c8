// Traced affine fill of the packed inter-sequence route with the traceback choices carried in
// the two low bits of every DP value ("tags") instead of in predicates.
//
// The first traced kernel (interseq_fill_kernel<.., TRACED = true>) takes its four choices per
// cell from the predicate outputs of VIMNMX.S16x2 and accumulates them with eight predicated
// adds per register of two cells: 15 of its 24 issue cycles.  Here every score v
// is kept as 4 v + tag, so that the maximum itself says which candidate won, ties included, and
// all four maxima fuse with their adds into VIADDMNMX.S16x2:
//
//     M1  = Hc(i-1,j-1) + P               P = 4 sim + 1 (profile byte), Hc = H' with the tag cleared
//     F1' = max(F1c(i,j-1) + 4 ext, M1(i,j-1))            tag 1 <=> o1 = [M >= F1 + ext]
//     F2' = max(F2c(i-1,j) + 4 ext, M1(i-1,j) + 1)        tag 2 <=> o2 = [M >= F2 + ext]
//     T'  = max(F1c + 1, F2c)                             tag 1 <=> hG = [F1 >= F2]
//     H'  = max(T' + 4 open, M1 + 1)                      tag 2 <=> hM = [M >= T + open]
//
// A candidate that must win ties (the reference's priorities, cell.rs:219-254, :331-394: Match
// from M before G1 before G2, a gap opened from M before it is extended) carries the larger tag
// and the other candidate a cleared one (one LOP3 per value: F1c, F2c, Hc), so equal scores
// resolve exactly like the predicates of the first kernel.  The nibble of a cell is
// (H' - Hc) + 4 ((F1' | F2') & 3): bits 0-1 = 2 (Match), 1 (Gap1) or 0 (Gap2) for the state
// holding H, bit 2 = o1, bit 3 = o2; four rows of both pairs make one 32-bit word (low pair in
// bits 0-15, high pair in bits 16-31) accumulated by multiply-adds on the FMA pipe.
//
// The row buffer between stripes has two planes instead of three: the F2' the next stripe's
// first row would compute (its tag is that row's o2 bit, made when the value is stored) and Hc.
//
// Exactness: scores scaled by 4 leave 13 bits, the route is taken only if every finite value
// provably stays inside +-8191 and the matrix entries inside [-32, 31] (profile byte 4 sim + 1);
// interseq_tag_eligible.  Everything else takes the first kernel.
#pragma once
// included by interseq_kernel.cuh after its kernels (structs, bt_pick16, PairScan) are defined

namespace bt {

constexpr uint32_t TG_CLR = 0xfffcfffcu;   // clears both tags
constexpr uint32_t TG_ONE = 0x00010001u;
constexpr uint32_t TG_TWO = 0x00020002u;
constexpr uint32_t TG_MASK = 0x00030003u;

inline bool interseq_tag_eligible(const BtParams &P, const PairScan &scan, int32_t min_score, int32_t max_score) {
    if (!P.affine || P.banded || !P.use_matrix) return false;
    if (min_score < -32 || max_score > 31) return false;
    if (const char *f = getenv("B200ALIGN_TAG")) { if (atoi(f) == 0) return false; }
    const int64_t open = P.gap_open, ext = P.gap_ext;
    const int64_t mn = std::min<int64_t>(min_score, 0), mx = std::max<int64_t>(max_score, 0);
    const int64_t neg = -8192 - open - ext - mn;
    const int64_t lowest = 3 * open + (scan.nmax + scan.mmax + 2 + IS_R) * ext + 2 * mn;
    const int64_t highest = std::min(scan.nmax, scan.mmax) * mx - open + mx;
    return lowest > neg + 1 && highest < 8190;
}

// cm = max(cm, v) per half; where v is strictly larger the row index r replaces ri_lo / ri_hi
// (ptxas fuses max + compares into one VIMNMX.S16x2 with two predicate outputs)
#define TG_MAX_ARG(cm, v, ri_lo, ri_hi, r)                                                \
    asm("{\n\t.reg .pred pl, ph;\n\t.reg .s16 x0, x1, y0, y1;\n\t.reg .b32 t;\n\t"        \
        "max.s16x2 t, %0, %3;\n\t"                                                       \
        "mov.b32 {x0, x1}, t;\n\tmov.b32 {y0, y1}, %0;\n\t"                              \
        "setp.eq.s16 pl, x0, y0;\n\tsetp.eq.s16 ph, x1, y1;\n\t"                         \
        "@!pl mov.u32 %1, %4;\n\t@!ph mov.u32 %2, %4;\n\t"                               \
        "mov.b32 %0, t;\n\t}"                                                            \
        : "+r"(cm), "+r"(ri_lo), "+r"(ri_hi)                                               \
        : "r"(v), "r"(r))

// decode one half of a tagged register
__device__ __forceinline__ int tg_lo(uint32_t x) { return (int)(int16_t)(x & 0xffffu) >> 2; }
__device__ __forceinline__ int tg_hi(uint32_t x) { return (int)(int16_t)(x >> 16) >> 2; }

// Persistent: one block per SM, `blockDim.x / 32` warps per block, every warp takes the next
// group of 64 pairs from a counter until none is left (no partial last wave, and the 1 KiB of
// shared memory the system reserves per block is paid once per SM instead of once per warp).
// The stripe profile of a warp holds the first `n_eff` column letters only: the launcher picks
// n_eff and the number of warps that then fit (interseq_launch_tag_fill); `ctl` carries the
// largest column letter the packer saw in this window, so that a variant built for fewer letters
// than the matrix has steps aside when a rarer letter turns up (the variant for all letters,
// launched right after it, then does the work).
template <int MODE, int PF, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
interseq_tag_fill_kernel(const InterseqConst C, const InterseqArgs A, const int n_eff, const int small_letters,
                         int *const ctl, const int counter_slot) {
    extern __shared__ uint4 smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        // ctl[0]: largest column letter of the window.  A variant whose profile is shorter than the
        // matrix runs only if every letter fits; the full variant only if the short one did not run.
        const int maxcol = ctl[0];
        if (n_eff < C.n_cols ? maxcol >= n_eff : (small_letters > 0 && maxcol < small_letters)) return;
    }
    uint4 *const smem_prof = smem_all + (size_t)warp * 2 * n_eff * 32;
    const uint32_t ext4 = ((uint32_t)(4 * C.gap_ext) & 0xffffu) * 0x10001u;
    const uint32_t open4 = ((uint32_t)(4 * C.gap_open) & 0xffffu) * 0x10001u;
    const uint32_t neg4 = ((uint32_t)(4 * C.neg) & 0xffffu) * 0x10001u;
    // constants the compiler must keep in registers (an immediate operand of a packed DPX
    // instruction is rebuilt with a PRMT for every use, see the first kernel)
    const uint32_t opaque = (uint32_t)C.n_cols >> 31;            // 0
    const uint32_t one2 = TG_ONE + opaque, clr = TG_CLR - opaque, msk = TG_MASK + opaque;
    const uint32_t pbase = (uint32_t)__cvta_generic_to_shared(smem_prof) + lane * 16;
    const uint32_t pbase_hi = pbase + (uint32_t)n_eff * 512u;
  for (;;) {
    int64_t g = 0;
    if (lane == 0) g = atomicAdd(ctl + counter_slot, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= A.n_groups) break;
    const WarpDesc d = A.desc[g];
    const int32_t p_lo = A.order[g * 64 + 2 * lane], p_hi = A.order[g * 64 + 2 * lane + 1];
    int n_lo = 0, n_hi = 0, m_lo = 0, m_hi = 0;
    if (p_lo >= 0) { n_lo = (int)A.map.len1(p_lo); m_lo = (int)A.map.len2(p_lo); }
    if (p_hi >= 0) { n_hi = (int)A.map.len1(p_hi); m_hi = (int)A.map.len2(p_hi); }

    const uint16_t *rc = A.rowcodes + d.rowcode_off + lane;
    const uint16_t *cc = A.colcodes + d.colcode_off + lane;
    // row buffer: [j][plane][lane], planes F2' of the next row (tagged) and Hc of the last row
    uint32_t *const rb = A.rowbuf + d.rowbuf_off * 2 + lane;
    uint4 *dir = (uint4 *)(A.dirs + d.dir_off) + lane;
    uint32_t *lastrow_lo = MODE == IS_SEMI ? A.lastrow + d.rowbuf_off * 2 + lane : nullptr;
    uint32_t *lastrow_hi = MODE == IS_SEMI ? lastrow_lo + (int64_t)d.mmax * 32 : nullptr;
    uint32_t *lastcol_lo = MODE == IS_SEMI ? A.lastcol + d.rowcode_off * 2 + lane : nullptr;
    uint32_t *lastcol_hi = MODE == IS_SEMI ? lastcol_lo + (int64_t)d.stripes * IS_R * 32 : nullptr;

    int32_t best_lo = 0, best_hi = 0, bidx_lo = 0, bidx_hi = 0, brow_lo = 0, brow_hi = 0;
    uint32_t best2 = 0x00010001u;      // local: 4 best + 1 of both pairs, the form the M1 registers have
    // table row 0 through the row buffer (pairwise.rs:747-780): M = G2 = -inf, so row 1 takes
    // F2 = -inf "from M" (the first kernel's o2 = [neg >= neg + ext]); H = open + (j-1) ext or 0
    {
        uint32_t h_row0 = MODE == IS_GLOBAL ? open4 : 0u;
        uint32_t *w = rb;
        for (int j = 0; j < d.mmax; j++, w += 64) {
            w[0] = neg4 + TG_TWO; w[32] = h_row0;
            if (MODE == IS_GLOBAL) h_row0 = __vadd2(h_row0, ext4);
        }
    }
    for (int s = 0; s < d.stripes; s++) {
        const int r0 = s * IS_R;
        const int j_lo = (n_lo - 1) / IS_R == s ? m_lo - 1 : -1;
        const int j_hi = (n_hi - 1) / IS_R == s ? m_hi - 1 : -1;
        const int rl_lo = (n_lo - 1) / IS_R == s ? (n_lo - 1) % IS_R : -1;
        const int rl_hi = (n_hi - 1) / IS_R == s ? (n_hi - 1) % IS_R : -1;
        const bool rows_valid = r0 + IS_R <= n_lo && r0 + IS_R <= n_hi;
        uint32_t M1l[IS_R], F1cl[IS_R], Hcl[IS_R];
        {
            // stripe profile: byte r of prof[half][b][lane] = 4 matrix[row letter r][b] + 1
            uint32_t a_lo[IS_R], a_hi[IS_R];
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                const uint32_t v = rc[(r0 + r) * 32];
                a_lo[r] = (v & 0xffu) * (uint32_t)C.n_cols;
                a_hi[r] = (v >> 8) * (uint32_t)C.n_cols;
            }
            __syncwarp();
            for (int b = 0; b < n_eff; b++) {
                const int32_t *col = A.matrix + b;
                uint32_t w_lo[4], w_hi[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    w_lo[q] = w_hi[q] = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        w_lo[q] |= ((uint32_t)(4 * __ldg(col + a_lo[4 * q + k]) + 1) & 0xffu) << (8 * k);
                        w_hi[q] |= ((uint32_t)(4 * __ldg(col + a_hi[4 * q + k]) + 1) & 0xffu) << (8 * k);
                    }
                }
                smem_prof[b * 32 + lane] = make_uint4(w_lo[0], w_lo[1], w_lo[2], w_lo[3]);
                smem_prof[(n_eff + b) * 32 + lane] = make_uint4(w_hi[0], w_hi[1], w_hi[2], w_hi[3]);
            }
            __syncwarp();
        }
#pragma unroll
        for (int r = 0; r < IS_R; r++) {
            // column 0 (pairwise.rs:707-745): M = G1 = -inf, H = G2 = open + (i-1) ext (or 0)
            M1l[r] = neg4 + TG_ONE;
            F1cl[r] = neg4;
            Hcl[r] = MODE == IS_GLOBAL ? ((uint32_t)(4 * (C.gap_open + (r0 + r) * C.gap_ext)) & 0xffffu) * 0x10001u : 0u;
        }
        uint32_t hu_prev = (s == 0 || MODE != IS_GLOBAL)
                               ? 0u : ((uint32_t)(4 * (C.gap_open + (r0 - 1) * C.gap_ext)) & 0xffffu) * 0x10001u;
        const uint16_t *ccp = cc;
        uint32_t *rbp = rb;
        uint4 *dirp = dir + (int64_t)s * d.mmax * 32;
        uint32_t q_c[PF], q_f[PF], q_h[PF];
#pragma unroll
        for (int k = 0; k < PF; k++) { q_c[k] = ccp[k * 32]; q_f[k] = rbp[k * 64]; q_h[k] = rbp[k * 64 + 32]; }
        // One column of the stripe.  `u` (compile-time after unrolling) names the look-ahead
        // registers of this column; CAPTURE = false leaves out everything that depends on where
        // the pairs end, so that PF columns form one basic block and ptxas overlaps the tail of
        // a column (the vertical F2 -> T -> H chain) with the head of the next one.
        auto column = [&](auto capture_tag, const int u, const int j) {
            constexpr bool CAPTURE = decltype(capture_tag)::value;
            const uint32_t v = q_c[u];
            const uint32_t up_f2 = q_f[u];     // F2' of this stripe's first row (tagged)
            const uint32_t hu = q_h[u];        // Hc of the row above
            q_c[u] = ccp[PF * 32];
            q_f[u] = rbp[PF * 64]; q_h[u] = rbp[PF * 64 + 32];
            if (IS_PREFETCH > 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + (PF + IS_PREFETCH) * 64));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rbp + (PF + IS_PREFETCH) * 64 + 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ccp + (PF + IS_PREFETCH) * 32));
            }
            uint32_t pl[4], ph[4];
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(pl[0]), "=r"(pl[1]), "=r"(pl[2]), "=r"(pl[3]) : "r"(pbase + (v & 0xffu) * 512u));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(ph[0]), "=r"(ph[1]), "=r"(ph[2]), "=r"(ph[3]) : "r"(pbase_hi + (v >> 8) * 512u));
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            uint32_t sim[IS_R], f1t[IS_R];
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                const uint32_t sel = (uint32_t)(r & 3) * 0x1111u + 0xC480u;
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(sim[r]) : "r"(pl[r >> 2]), "r"(ph[r >> 2]), "r"(sel));
            }
            // rows are software-pipelined as in the first kernel: F1' and M1 of row r + 1 are
            // formed from the previous column's state before Hc of row r is replaced
#define TG_ROW_F1_M(r, diag_hc)                                                              \
    {                                                                                        \
        f1t[r] = __viaddmax_s16x2(F1cl[r], ext4, M1l[r]);                                    \
        F1cl[r] = f1t[r] & clr;                                                              \
        M1l[r] = MODE == IS_LOCAL ? __viaddmax_s16x2(diag_hc, sim[r], one2) : __vadd2(diag_hc, sim[r]); \
    }
            TG_ROW_F1_M(0, hu_prev);
            hu_prev = hu;
            uint32_t f2t = up_f2, up_m2 = 0u, f2c = 0u;
#pragma unroll
            for (int r = 0; r < IS_R; r++) {
                if (r + 1 < IS_R) TG_ROW_F1_M(r + 1, Hcl[r]);
                if (r > 0) f2t = __viaddmax_s16x2(f2c, ext4, up_m2);
                f2c = f2t & clr;
                const uint32_t t = __viaddmax_s16x2(F1cl[r], one2, f2c);
                up_m2 = __vadd2(M1l[r], one2);
                const uint32_t h = __viaddmax_s16x2(t, open4, up_m2);
                const uint32_t hc = h & clr;
                Hcl[r] = hc;
                const uint32_t tf = (f1t[r] | f2t) & msk;
                const uint32_t nib = (h - hc) + tf * 4u;
                acc[r >> 2] = nib * (1u << (4 * (r & 3))) + acc[r >> 2];
            }
            // the F2' of the row below the stripe, and Hc of the stripe's last row
            rbp[0] = __viaddmax_s16x2(f2c, ext4, up_m2);
            rbp[32] = Hcl[IS_R - 1];
            *dirp = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            dirp += 32;
            ccp += 32; rbp += 64;
            if (!CAPTURE) {
                if (MODE == IS_LOCAL) {
                    // Local, stripes whose 16 rows are table rows of every pair of the warp: the first
                    // maximum of M in row-major order (pairwise.rs:849-865) without leaving the column's
                    // basic block - a packed arg-max over the rows, then one predicated scalar update per
                    // pair.  (The capture path below, which also knows about rows past the end of a pair,
                    // ran on every column before: homologous pairs raise their maximum at almost every
                    // column, so some lane always took its 16-row scalar scan and the traced local fill
                    // ran at half the rate of the global one.)
                    // Columns in which no pair of the warp reaches its running maximum (most of them: the
                    // pairs of a group are alike, their maxima grow in the same few columns of a stripe)
                    // cost 15 maxima and a vote; the others the arg-max with the rows.
                    uint32_t cm = M1l[0];
#pragma unroll
                    for (int r = 1; r < IS_R; r++) cm = __vmaxs2(cm, M1l[r]);
                    bool ge_hi, ge_lo;
                    (void)__vibmax_s16x2(cm, best2, &ge_hi, &ge_lo);              // cm >= best, per half
                    const bool need = (ge_lo & (j < m_lo) & ((cm & 0xffffu) > 1u)) | (ge_hi & (j < m_hi) & ((cm >> 16) > 1u));
                    if (__any_sync(0xffffffffu, need)) {
                        uint32_t cm2 = M1l[0];
                        int ri_lo = 0, ri_hi = 0;
#pragma unroll
                        for (int r = 1; r < IS_R; r++) TG_MAX_ARG(cm2, M1l[r], ri_lo, ri_hi, r);
                        const int c_lo = tg_lo(cm2), c_hi = tg_hi(cm2);
                        const int row_lo = r0 + ri_lo + 1, row_hi = r0 + ri_hi + 1;
                        // a later cell with the same score comes first in row-major order only in a higher row
                        const bool up_lo = (j < m_lo) & ((c_lo > best_lo) | ((c_lo == best_lo) & (c_lo > 0) & (row_lo < brow_lo)));
                        const bool up_hi = (j < m_hi) & ((c_hi > best_hi) | ((c_hi == best_hi) & (c_hi > 0) & (row_hi < brow_hi)));
                        best_lo = up_lo ? c_lo : best_lo;
                        brow_lo = up_lo ? row_lo : brow_lo;
                        bidx_lo = up_lo ? row_lo * (m_lo + 1) + j + 1 : bidx_lo;
                        best_hi = up_hi ? c_hi : best_hi;
                        brow_hi = up_hi ? row_hi : brow_hi;
                        bidx_hi = up_hi ? row_hi * (m_hi + 1) + j + 1 : bidx_hi;
                        best2 = ((uint32_t)(4 * best_lo + 1) & 0xffffu) | ((uint32_t)(4 * best_hi + 1) << 16);
                    }
                }
                return;
            }

            if (MODE == IS_GLOBAL) {
                if (j == j_lo || j == j_hi) {
                    // optimum: H(n, m) (pairwise.rs:866-879); Hc holds 4 H
                    if (j == j_lo) best_lo = tg_lo(bt_pick16(Hcl, rl_lo));
                    if (j == j_hi) best_hi = tg_hi(bt_pick16(Hcl, rl_hi));
                }
            } else if (MODE == IS_SEMI) {
                if (rl_lo >= 0 || rl_hi >= 0) {
                    uint32_t w_lo = 0, w_hi = 0;
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        if (r == rl_lo) w_lo = M1l[r];
                        if (r == rl_hi) w_hi = M1l[r];
                    }
                    if (rl_lo >= 0) lastrow_lo[(int64_t)j * 32] = w_lo;
                    if (rl_hi >= 0) lastrow_hi[(int64_t)j * 32] = w_hi;
                }
                if (j == m_lo - 1) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) lastcol_lo[(int64_t)(r0 + r) * 32] = M1l[r];
                }
                if (j == m_hi - 1) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) lastcol_hi[(int64_t)(r0 + r) * 32] = M1l[r];
                }
            } else {
                // local: first maximum of M in row-major order (pairwise.rs:849-865); M1 = 4 M + 1
                // orders like M, the scalar path decodes it
                uint32_t cm = M1l[0];
#pragma unroll
                for (int r = 1; r < IS_R; r++) cm = __vmaxs2(cm, M1l[r]);
                const bool inside = rows_valid && j < m_lo && j < m_hi;
                bool slow = !inside;
                if (inside) {
                    const int c_lo = tg_lo(cm), c_hi = tg_hi(cm);
                    slow = c_lo > best_lo || c_hi > best_hi ||
                           (c_lo == best_lo && c_lo > 0 && brow_lo > r0 + 1) ||
                           (c_hi == best_hi && c_hi > 0 && brow_hi > r0 + 1);
                }
                if (slow) {
#pragma unroll
                    for (int r = 0; r < IS_R; r++) {
                        const int row = r0 + r + 1, col = j + 1;
                        const int v_lo = tg_lo(M1l[r]), v_hi = tg_hi(M1l[r]);
                        if (row <= n_lo && col <= m_lo) {
                            const int idx = row * (m_lo + 1) + col;
                            if (v_lo > best_lo || (v_lo == best_lo && v_lo > 0 && idx < bidx_lo)) { best_lo = v_lo; bidx_lo = idx; brow_lo = row; }
                        }
                        if (row <= n_hi && col <= m_hi) {
                            const int idx = row * (m_hi + 1) + col;
                            if (v_hi > best_hi || (v_hi == best_hi && v_hi > 0 && idx < bidx_hi)) { best_hi = v_hi; bidx_hi = idx; brow_hi = row; }
                        }
                    }
                }
            }
        };
        // columns below `jcut` need no capture in any lane of the warp
        int jcut = 0;
        if (MODE == IS_GLOBAL) {
            int mine = d.mmax;
            if (j_lo >= 0) mine = min(mine, j_lo);
            if (j_hi >= 0) mine = min(mine, j_hi);
            jcut = __reduce_min_sync(0xffffffffu, mine);
        } else if (MODE == IS_SEMI) {
            int mine = (rl_lo >= 0 || rl_hi >= 0) ? 0 : d.mmax;
            if (p_lo >= 0) mine = min(mine, m_lo - 1);
            if (p_hi >= 0) mine = min(mine, m_hi - 1);
            jcut = __reduce_min_sync(0xffffffffu, mine);
        } else {
            // local: the capture-free columns track the optimum themselves when every lane's rows are real
            jcut = __all_sync(0xffffffffu, rows_valid) ? d.mmax : 0;
            best2 = ((uint32_t)(4 * best_lo + 1) & 0xffffu) | ((uint32_t)(4 * best_hi + 1) << 16);
        }
        int j0 = 0;
        for (; j0 + PF <= jcut; j0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; u++) column(std::false_type{}, u, j0 + u);
        }
        for (; j0 < d.mmax; j0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; u++) {
                if (u > 0 && j0 + u >= d.mmax) break;
                column(std::true_type{}, u, j0 + u);
            }
        }
    }
    if (MODE != IS_SEMI) {
        if (p_lo >= 0) A.score[p_lo] = best_lo;
        if (p_hi >= 0) A.score[p_hi] = best_hi;
    }
    if (MODE == IS_LOCAL) {
        if (p_lo >= 0) { A.start[2 * p_lo] = bidx_lo / (m_lo + 1); A.start[2 * p_lo + 1] = bidx_lo % (m_lo + 1); }
        if (p_hi >= 0) { A.start[2 * p_hi] = bidx_hi / (m_hi + 1); A.start[2 * p_hi + 1] = bidx_hi % (m_hi + 1); }
    }
    __syncwarp();   // the next group's profile overwrites this one's
  }
}

constexpr int TG_SMEM_BYTES = 226 * 1024;   // per SM: 227 KiB less the block's reserved KiB
constexpr int TG_SMALL_LETTERS = 20;        // the standard amino acids: 11 warps per SM instead of 9

// warps per SM whose stripe profiles of n_eff letters fit
inline int interseq_tag_warps(int n_eff) { return std::max(1, std::min(16, TG_SMEM_BYTES / (2 * n_eff * 512))); }

template <int MODE>
inline cudaError_t interseq_launch_tag_variant(const InterseqConst &C, const InterseqArgs &A, int n_eff,
                                               int small_letters, int *ctl, int counter_slot, int sms,
                                               cudaStream_t stream) {
    const int warps = interseq_tag_warps(n_eff);
    const size_t smem = (size_t)warps * 2 * n_eff * 512;
    const unsigned grid = (unsigned)std::min<int64_t>(sms, (A.n_groups + warps - 1) / warps);
    cudaError_t e;
    if (warps > 11) {   // small alphabets: up to 16 warps, 128 registers each
        auto kernel = interseq_tag_fill_kernel<MODE, BT_PF_TRACED, 16>;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kernel<<<grid, 32 * warps, smem, stream>>>(C, A, n_eff, small_letters, ctl, counter_slot);
    } else {
        auto kernel = interseq_tag_fill_kernel<MODE, BT_PF_TRACED, 11>;
        if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        kernel<<<grid, 32 * warps, smem, stream>>>(C, A, n_eff, small_letters, ctl, counter_slot);
    }
    return cudaGetLastError();
}

// ctl: three ints of the window's slot - [0] largest column letter (written by the packer),
// [1], [2] group counters of the two variants (zeroed with [0] before the packer runs).
template <int MODE>
inline cudaError_t interseq_launch_tag_fill(const InterseqConst &C, const InterseqArgs &A, int *ctl, int sms,
                                            cudaStream_t stream, int64_t *launches) {
    cudaError_t e = cudaSuccess;
    int small_letters = 0;
    if (C.n_cols > TG_SMALL_LETTERS && interseq_tag_warps(TG_SMALL_LETTERS) > interseq_tag_warps(C.n_cols)) {
        // matrices with more letters than the 20 standard amino acids (B, Z, X, * of the protein
        // alphabet): windows that use only those get the larger number of warps
        small_letters = TG_SMALL_LETTERS;
        e = interseq_launch_tag_variant<MODE>(C, A, small_letters, small_letters, ctl, 1, sms, stream);
        if (e != cudaSuccess) return e;
        (*launches)++;
    }
    e = interseq_launch_tag_variant<MODE>(C, A, C.n_cols, small_letters, ctl, 2, sms, stream);
    if (e == cudaSuccess) (*launches)++;
    return e;
}

}  // namespace bt
