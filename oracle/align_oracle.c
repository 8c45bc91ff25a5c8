/*
 * align_oracle.c - CPU restatement of the reference's pairwise-alignment DP.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (biotite_b200/) may
 * import, link or call this file; it is the checker for the CUDA kernels
 * (tests/, __graft_entry__.smoke()) and the timed CPU baseline of bench.py.
 *
 * It restates, in plain scalar C with int32 scores and one table cell per DP
 * position (the reference's memory layout), the algorithm of
 *   src/rust/sequence/align/pairwise.rs:161-267   (align_optimal_generic)
 *   src/rust/sequence/align/pairwise.rs:378-415   (linear recurrence)
 *   src/rust/sequence/align/pairwise.rs:430-467   (linear first row/column)
 *   src/rust/sequence/align/pairwise.rs:532-559   (linear traceback starts)
 *   src/rust/sequence/align/pairwise.rs:585-681   (affine recurrence, neg_inf)
 *   src/rust/sequence/align/pairwise.rs:695-780   (affine first row/column)
 *   src/rust/sequence/align/pairwise.rs:845-881   (affine traceback starts)
 *   src/rust/sequence/align/banded.rs:164-255     (align_banded_generic)
 *   src/rust/sequence/align/banded.rs:310-414     (banded linear)
 *   src/rust/sequence/align/banded.rs:436-593     (banded affine)
 *   src/rust/sequence/align/banded.rs:598-638     (get_potential_starts)
 *   src/rust/sequence/align/cell.rs:125-138,219-254 (traceback step order)
 *   src/rust/sequence/align/cell.rs:291-394       (max + tie sets)
 *   src/rust/sequence/align/trace.rs:46-142       (follow_trace, path_to_trace)
 *   src/rust/sequence/align/scoring.rs:62-64,101-115 (matrix / match scoring)
 * The reference itself is Rust/PyO3 and cannot be built in this image (no
 * cargo/rustc), so parity is pinned against the reference's golden vectors
 * (tests/golden/, see tests/golden/make_golden.py), not against a reference
 * binary.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t score_t;

/* i32 arithmetic wraps silently in the reference (Cargo.toml: overflow-checks = false) */
static inline score_t wadd(score_t a, score_t b) {
    return (score_t)((uint32_t)a + (uint32_t)b);
}
static inline score_t wsub(score_t a, score_t b) {
    return (score_t)((uint32_t)a - (uint32_t)b);
}
static inline score_t smax(score_t a, score_t b) { return a > b ? a : b; }

/* direction bit sets, cell.rs:44-80 */
enum { L_MATCH = 1, L_GAP1 = 2, L_GAP2 = 4 };
enum {
    A_MM = 1, A_G1M = 2, A_G2M = 4, A_MG1 = 8, A_G1G1 = 16, A_MG2 = 32, A_G2G2 = 64
};
enum { ST_MATCH = 0, ST_GAP1 = 1, ST_GAP2 = 2 };
enum { DIR_DIAG = 0, DIR_TO1 = 1, DIR_TO2 = 2 };

typedef struct {
    score_t m, g1, g2; /* linear mode uses only m */
    uint8_t dir;
} cell_t;

typedef struct {
    const int64_t *code1, *code2;
    int64_t n, m;
    const int32_t *matrix; /* NULL -> match/mismatch */
    int64_t n_cols;
    score_t match, mismatch;
    int affine;
    score_t gap_open, gap_ext; /* linear: gap_open is the penalty */
    int local, terminal, score_only;
    score_t neg_inf;
} params_t;

static inline score_t similarity(const params_t *p, int64_t s1, int64_t s2) {
    if (p->matrix) return p->matrix[s1 * p->n_cols + s2]; /* scoring.rs:62-64 */
    return s1 == s2 ? p->match : p->mismatch;             /* scoring.rs:101-108 */
}

static score_t scheme_min(const int32_t *matrix, int64_t n_rows, int64_t n_cols,
                          score_t match, score_t mismatch) {
    if (!matrix) return match < mismatch ? match : mismatch;
    int64_t total = n_rows * n_cols;
    if (total == 0) return 0;
    score_t mn = matrix[0];
    for (int64_t k = 1; k < total; k++)
        if (matrix[k] < mn) mn = matrix[k];
    return mn;
}

/* cell.rs:291-323 */
static inline score_t trace_linear(score_t a, score_t b, score_t c, uint8_t *dir) {
    if (a > b) {
        if (a > c) { *dir = L_MATCH; return a; }
        if (a == c) { *dir = L_MATCH | L_GAP2; return a; }
        *dir = L_GAP2; return c;
    } else if (a == b) {
        if (a > c) { *dir = L_MATCH | L_GAP1; return a; }
        if (a == c) { *dir = L_MATCH | L_GAP1 | L_GAP2; return a; }
        *dir = L_GAP2; return c;
    } else if (b > c) { *dir = L_GAP1; return b; }
    else if (b == c) { *dir = L_GAP1 | L_GAP2; return b; }
    *dir = L_GAP2; return c;
}

/* pairwise.rs:378-415 and banded.rs:348-375 */
static inline cell_t fill_linear(const params_t *p, int64_t s1, int64_t s2,
                                 const cell_t *fm, const cell_t *f1, const cell_t *f2,
                                 int waive1, int waive2) {
    cell_t out = {0, 0, 0, 0};
    score_t a = wadd(fm->m, similarity(p, s1, s2));
    score_t b = waive1 ? f1->m : wadd(f1->m, p->gap_open);
    score_t c = waive2 ? f2->m : wadd(f2->m, p->gap_open);
    if (p->score_only) {
        out.m = smax(smax(a, b), c);
    } else {
        out.m = trace_linear(a, b, c, &out.dir);
    }
    if (p->local && out.m <= 0) { out.m = 0; out.dir = 0; }
    return out;
}

/* pairwise.rs:608-681, banded.rs:486-548, cell.rs:331-394 */
static inline cell_t fill_affine(const params_t *p, int64_t s1, int64_t s2,
                                 const cell_t *fm, const cell_t *f1, const cell_t *f2,
                                 int waive1, int waive2) {
    cell_t out;
    score_t sim = similarity(p, s1, s2);
    score_t mm = wadd(fm->m, sim), g1m = wadd(fm->g1, sim), g2m = wadd(fm->g2, sim);
    score_t mg1, g1g1, mg2, g2g2;
    if (waive1) { mg1 = f1->m; g1g1 = f1->g1; }
    else { mg1 = wadd(f1->m, p->gap_open); g1g1 = wadd(f1->g1, p->gap_ext); }
    if (waive2) { mg2 = f2->m; g2g2 = f2->g2; }
    else { mg2 = wadd(f2->m, p->gap_open); g2g2 = wadd(f2->g2, p->gap_ext); }
    uint8_t dir = 0;
    if (p->score_only) {
        out.m = smax(smax(mm, g1m), g2m);
        out.g1 = smax(mg1, g1g1);
        out.g2 = smax(mg2, g2g2);
    } else {
        uint8_t d3;
        out.m = trace_linear(mm, g1m, g2m, &d3); /* same tie logic, bits 1,2,4 */
        dir = d3;
        if (mg1 > g1g1) { dir |= A_MG1; out.g1 = mg1; }
        else if (mg1 < g1g1) { dir |= A_G1G1; out.g1 = g1g1; }
        else { dir |= A_MG1 | A_G1G1; out.g1 = mg1; }
        if (mg2 > g2g2) { dir |= A_MG2; out.g2 = mg2; }
        else if (mg2 < g2g2) { dir |= A_G2G2; out.g2 = g2g2; }
        else { dir |= A_MG2 | A_G2G2; out.g2 = mg2; }
    }
    if (p->local) {
        if (out.m <= 0) { dir &= (uint8_t)~(A_MM | A_G1M | A_G2M); out.m = 0; }
        if (out.g1 <= 0) { dir &= (uint8_t)~(A_MG1 | A_G1G1); out.g1 = p->neg_inf; }
        if (out.g2 <= 0) { dir &= (uint8_t)~(A_MG2 | A_G2G2); out.g2 = p->neg_inf; }
    }
    out.dir = dir;
    return out;
}

/* ---- traceback ---------------------------------------------------------------- */

typedef struct { int dir; int state; } step_t;

/* cell.rs:125-138 (linear) and :219-254 (affine): ordered list of steps */
static int cell_steps(int affine, uint8_t d, int state, step_t steps[3]) {
    int k = 0;
    if (!affine) {
        if (d & L_MATCH) { steps[k].dir = DIR_DIAG; steps[k++].state = 0; }
        if (d & L_GAP1) { steps[k].dir = DIR_TO1; steps[k++].state = 0; }
        if (d & L_GAP2) { steps[k].dir = DIR_TO2; steps[k++].state = 0; }
        return k;
    }
    switch (state) {
    case ST_MATCH:
        if (d & A_MM) { steps[k].dir = DIR_DIAG; steps[k++].state = ST_MATCH; }
        if (d & A_G1M) { steps[k].dir = DIR_DIAG; steps[k++].state = ST_GAP1; }
        if (d & A_G2M) { steps[k].dir = DIR_DIAG; steps[k++].state = ST_GAP2; }
        break;
    case ST_GAP1:
        if (d & A_MG1) { steps[k].dir = DIR_TO1; steps[k++].state = ST_MATCH; }
        if (d & A_G1G1) { steps[k].dir = DIR_TO1; steps[k++].state = ST_GAP1; }
        break;
    default:
        if (d & A_MG2) { steps[k].dir = DIR_TO2; steps[k++].state = ST_MATCH; }
        if (d & A_G2G2) { steps[k].dir = DIR_TO2; steps[k++].state = ST_GAP2; }
        break;
    }
    return k;
}

typedef struct {
    int64_t *idx;
    int64_t len, cap;
} path_t;

typedef struct {
    path_t *paths;
    int64_t n, cap;
} pathlist_t;

static void path_push(path_t *p, int64_t v) {
    if (p->len == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 64;
        p->idx = (int64_t *)realloc(p->idx, (size_t)p->cap * sizeof(int64_t));
    }
    p->idx[p->len++] = v;
}

static path_t path_clone(const path_t *p) {
    path_t c;
    c.len = p->len;
    c.cap = p->len > 64 ? p->len * 2 : 64;
    c.idx = (int64_t *)malloc((size_t)c.cap * sizeof(int64_t));
    if (p->len) memcpy(c.idx, p->idx, (size_t)p->len * sizeof(int64_t));
    return c;
}

static void list_push(pathlist_t *l, path_t p) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 16;
        l->paths = (path_t *)realloc(l->paths, (size_t)l->cap * sizeof(path_t));
    }
    l->paths[l->n++] = p;
}

/* trace.rs:46-92: branches are explored (and emitted) before the main path */
static void follow_trace(const cell_t *table, int affine, int64_t start, int state,
                         const int64_t offs[3], int64_t *count, int64_t max_count,
                         pathlist_t *out, path_t path) {
    int64_t index = start;
    for (;;) {
        step_t steps[3];
        int ns = cell_steps(affine, table[index].dir, state, steps);
        if (ns == 0) break;
        path_push(&path, index);
        for (int k = 1; k < ns; k++) {
            if (*count < max_count) {
                *count += 1;
                follow_trace(table, affine, index + offs[steps[k].dir], steps[k].state,
                             offs, count, max_count, out, path_clone(&path));
            }
        }
        index += offs[steps[0].dir];
        state = steps[0].state;
    }
    list_push(out, path);
}

/* ---- result object ------------------------------------------------------------ */

typedef struct {
    score_t score;
    int64_t n_traces;
    int64_t *lengths;
    int64_t **traces; /* each lengths[k] x 2, row-major */
} result_t;

/* trace.rs:103-142: reverse, map to sequence positions, repeated value -> -1.
 * Positions are non-decreasing along a path, so "seen before" == "equals the
 * previous value of that column" (the very first -1 of a column stays -1). */
static void emit_trace(result_t *res, int64_t k, const path_t *path, int64_t stride,
                       int banded, int64_t lower) {
    int64_t len = path->len;
    int64_t *rows = (int64_t *)malloc((size_t)(len ? len : 1) * 2 * sizeof(int64_t));
    int64_t out = 0;
    int64_t prev0 = INT64_MIN, prev1 = INT64_MIN;
    for (int64_t r = len - 1; r >= 0; r--) {
        int64_t i = path->idx[r] / stride, j = path->idx[r] % stride;
        int64_t s0 = i - 1;
        int64_t s1 = banded ? j + s0 + lower - 1 : j - 1;
        int64_t v0 = (s0 == prev0) ? -1 : s0;
        int64_t v1 = (s1 == prev1) ? -1 : s1;
        prev0 = s0;
        prev1 = s1;
        if (v0 == -1 && v1 == -1) continue;
        rows[2 * out] = v0;
        rows[2 * out + 1] = v1;
        out++;
    }
    res->lengths[k] = out;
    res->traces[k] = rows;
}

static result_t *finish(const cell_t *table, const params_t *p, score_t score,
                        const int64_t *starts, const int *start_states, int64_t n_starts,
                        const int64_t offs[3], int64_t stride, int64_t max_number,
                        int banded, int64_t lower) {
    result_t *res = (result_t *)calloc(1, sizeof(result_t));
    res->score = score;
    if (p->score_only) return res;
    pathlist_t paths = {0, 0, 0};
    for (int64_t s = 0; s < n_starts; s++) {
        int64_t count = 1; /* pairwise.rs:241-254: counter reset per start */
        path_t empty = {0, 0, 0};
        follow_trace(table, p->affine, starts[s], start_states[s], offs, &count,
                     max_number, &paths, empty);
        /* the reference enumerates every start before truncating; stop early once
         * enough paths exist - later starts cannot change the first max_number */
        if (paths.n >= max_number) break;
    }
    int64_t keep = paths.n < max_number ? paths.n : max_number;
    if (keep < 0) keep = 0;
    res->n_traces = keep;
    res->lengths = (int64_t *)calloc((size_t)(keep ? keep : 1), sizeof(int64_t));
    res->traces = (int64_t **)calloc((size_t)(keep ? keep : 1), sizeof(int64_t *));
    for (int64_t k = 0; k < keep; k++)
        emit_trace(res, k, &paths.paths[k], stride, banded, lower);
    for (int64_t k = 0; k < paths.n; k++) free(paths.paths[k].idx);
    free(paths.paths);
    return res;
}

/* ---- align_optimal -------------------------------------------------------------- */

static result_t *run_optimal(const params_t *p, int64_t max_number) {
    const int64_t n = p->n, m = p->m, stride = m + 1;
    cell_t *table = (cell_t *)calloc((size_t)((n + 1) * (m + 1)), sizeof(cell_t));
    const score_t NI = p->neg_inf;
    const int semi = !p->terminal && !p->local;

    /* origin + first column + first row: pairwise.rs:188-203, :430-467, :695-780 */
    if (p->affine) {
        table[0].m = 0; table[0].g1 = NI; table[0].g2 = NI; table[0].dir = 0;
        for (int64_t i = 1; i <= n; i++) {
            cell_t *c = &table[i * stride];
            const cell_t *pr = &table[(i - 1) * stride];
            if (p->local) { c->m = NI; c->g1 = NI; c->g2 = 0; c->dir = 0; continue; }
            score_t open = p->terminal ? wadd(pr->m, p->gap_open) : pr->m;
            score_t ext = p->terminal ? wadd(pr->g2, p->gap_ext) : pr->g2;
            c->m = NI; c->g1 = NI;
            if (open >= ext) { c->g2 = open; c->dir = A_MG2; }
            else { c->g2 = ext; c->dir = A_G2G2; }
            if (p->score_only) c->dir = 0;
        }
        for (int64_t j = 1; j <= m; j++) {
            cell_t *c = &table[j];
            const cell_t *pr = &table[j - 1];
            if (p->local) { c->m = NI; c->g1 = 0; c->g2 = NI; c->dir = 0; continue; }
            score_t open = p->terminal ? wadd(pr->m, p->gap_open) : pr->m;
            score_t ext = p->terminal ? wadd(pr->g1, p->gap_ext) : pr->g1;
            c->m = NI; c->g2 = NI;
            if (open >= ext) { c->g1 = open; c->dir = A_MG1; }
            else { c->g1 = ext; c->dir = A_G1G1; }
            if (p->score_only) c->dir = 0;
        }
    } else {
        for (int64_t i = 1; i <= n; i++) {
            cell_t *c = &table[i * stride];
            if (p->local) continue;
            c->m = p->terminal ? wadd(table[(i - 1) * stride].m, p->gap_open) : 0;
            c->dir = p->score_only ? 0 : L_GAP2;
        }
        for (int64_t j = 1; j <= m; j++) {
            cell_t *c = &table[j];
            if (p->local) continue;
            c->m = p->terminal ? wadd(table[j - 1].m, p->gap_open) : 0;
            c->dir = p->score_only ? 0 : L_GAP1;
        }
    }

    /* hot loop: pairwise.rs:205-227 */
    for (int64_t i = 1; i <= n; i++) {
        const int64_t s1 = p->code1[i - 1];
        const int last_row = (i == n);
        cell_t *row = &table[i * stride];
        const cell_t *up = &table[(i - 1) * stride];
        for (int64_t j = 1; j <= m; j++) {
            const int last_col = (j == m);
            /* waive1: trailing gap in seq 1 (last row); waive2: last column */
            const int w1 = semi && last_row, w2 = semi && last_col;
            row[j] = p->affine
                ? fill_affine(p, s1, p->code2[j - 1], &up[j - 1], &row[j - 1], &up[j], w1, w2)
                : fill_linear(p, s1, p->code2[j - 1], &up[j - 1], &row[j - 1], &up[j], w1, w2);
        }
    }

    /* traceback starts: pairwise.rs:532-559, :845-881 */
    score_t score;
    int64_t n_starts = 0, cap = 0;
    int64_t *starts = NULL;
    int *states = NULL;
    const int64_t total = (n + 1) * (m + 1);
    if (p->local) {
        score = INT32_MIN;
        for (int64_t k = 0; k < total; k++) {
            score_t v = table[k].m;
            if (v > score) { score = v; n_starts = 0; }
            if (v == score) {
                if (p->score_only) { n_starts = 1; continue; }
                if (n_starts == cap) {
                    cap = cap ? cap * 2 : 16;
                    starts = (int64_t *)realloc(starts, (size_t)cap * sizeof(int64_t));
                    states = (int *)realloc(states, (size_t)cap * sizeof(int));
                }
                /* more starts than max_number can never all be used */
                if (n_starts < max_number) {
                    starts[n_starts] = k; states[n_starts] = ST_MATCH; n_starts++;
                }
            }
        }
    } else {
        starts = (int64_t *)malloc(3 * sizeof(int64_t));
        states = (int *)malloc(3 * sizeof(int));
        const cell_t *c = &table[total - 1];
        if (p->affine) {
            score = smax(smax(c->m, c->g1), c->g2);
            if (c->m == score) { starts[n_starts] = total - 1; states[n_starts++] = ST_MATCH; }
            if (c->g1 == score) { starts[n_starts] = total - 1; states[n_starts++] = ST_GAP1; }
            if (c->g2 == score) { starts[n_starts] = total - 1; states[n_starts++] = ST_GAP2; }
        } else {
            score = c->m;
            starts[0] = total - 1; states[0] = 0; n_starts = 1;
        }
    }
    const int64_t offs[3] = {-(stride + 1), -1, -stride};
    result_t *res = finish(table, p, score, starts, states, n_starts, offs, stride,
                           max_number, 0, 0);
    free(starts); free(states); free(table);
    return res;
}

/* ---- align_banded ---------------------------------------------------------------- */

static result_t *run_banded(const params_t *p, int64_t lower, int64_t upper,
                            int64_t max_number) {
    const int64_t n = p->n, m = p->m;
    const int64_t W = upper - lower + 1, stride = W + 2;
    const score_t NI = p->neg_inf;
    const int64_t total = (n + 1) * stride;
    cell_t *table = (cell_t *)malloc((size_t)total * sizeof(cell_t));
    /* banded.rs:184-190: edges everywhere, sentinels in columns 0 and W+1 */
    cell_t edge = {0, 0, 0, 0}, sent = {NI, 0, 0, 0};
    if (p->affine) { edge.g1 = NI; edge.g2 = NI; sent.g1 = NI; sent.g2 = NI; }
    for (int64_t k = 0; k < total; k++) table[k] = edge;
    for (int64_t i = 0; i <= n; i++) { table[i * stride] = sent; table[i * stride + W + 1] = sent; }

    /* banded.rs:199-214 */
    for (int64_t si = 0; si < n; si++) {
        const int64_t i = si + 1;
        int64_t j0 = si + lower; if (j0 < 0) j0 = 0;
        int64_t j1 = si + upper + 1; if (j1 > m) j1 = m; if (j1 < 0) j1 = 0;
        cell_t *row = &table[i * stride];
        const cell_t *up = &table[(i - 1) * stride];
        const int64_t s1 = p->code1[si];
        for (int64_t sj = j0; sj < j1; sj++) {
            const int64_t c = sj + 1 - si - lower;
            row[c] = p->affine
                ? fill_affine(p, s1, p->code2[sj], &up[c], &row[c - 1], &up[c + 1], 0, 0)
                : fill_linear(p, s1, p->code2[sj], &up[c], &row[c - 1], &up[c + 1], 0, 0);
        }
    }

    score_t score;
    int64_t n_starts = 0, cap = 0;
    int64_t *starts = NULL;
    int *states = NULL;
    if (p->local) {
        score = INT32_MIN;
        for (int64_t k = 0; k < total; k++) {
            score_t v = table[k].m;
            if (v > score) { score = v; n_starts = 0; }
            if (v == score) {
                if (p->score_only) { n_starts = 1; continue; }
                if (n_starts == cap) {
                    cap = cap ? cap * 2 : 16;
                    starts = (int64_t *)realloc(starts, (size_t)cap * sizeof(int64_t));
                    states = (int *)realloc(states, (size_t)cap * sizeof(int));
                }
                if (n_starts < max_number) {
                    starts[n_starts] = k; states[n_starts] = ST_MATCH; n_starts++;
                }
            }
        }
    } else {
        /* banded.rs:598-638 get_potential_starts, then :400-413 / :575-592 */
        int64_t *cand = (int64_t *)malloc((size_t)W * sizeof(int64_t));
        for (int64_t j = 1; j <= W; j++) {
            int64_t sj = j + (n - 1) + lower - 1;
            int64_t i = sj < m ? n : m - j - lower + 1;
            cand[j - 1] = i * stride + j;
        }
        score = INT32_MIN;
        for (int64_t k = 0; k < W; k++) {
            const cell_t *c = &table[cand[k]];
            score = smax(score, c->m);
            if (p->affine) score = smax(smax(score, c->g1), c->g2);
        }
        starts = (int64_t *)malloc((size_t)(3 * W) * sizeof(int64_t));
        states = (int *)malloc((size_t)(3 * W) * sizeof(int));
        const int n_states = p->affine ? 3 : 1;
        for (int st = 0; st < n_states; st++) {
            for (int64_t k = 0; k < W; k++) {
                const cell_t *c = &table[cand[k]];
                score_t v = st == ST_MATCH ? c->m : (st == ST_GAP1 ? c->g1 : c->g2);
                if (v == score) { starts[n_starts] = cand[k]; states[n_starts++] = st; }
            }
        }
        free(cand);
    }
    /* banded.rs:194-196: diag=(-1,0), gap1=(0,-1), gap2=(-1,+1) */
    const int64_t offs[3] = {-stride, -1, -stride + 1};
    result_t *res = finish(table, p, score, starts, states, n_starts, offs, stride,
                           max_number, 1, lower);
    free(starts); free(states); free(table);
    return res;
}

/* ---- align_region: local X-drop extension (localgapped.rs) ------------------------------ */

/* localgapped.rs:182-310 (fill_xdrop), :446-489 (linear rule), :558-641 (affine rule).
 * A score of 0 marks an unreached cell; the (0,0) anchor holds threshold + 1; the table is
 * filled antidiagonal by antidiagonal, every candidate is pruned against the running
 * maximum.  `status` = 1 when the reference's doubling table (100 x 100 initially,
 * localgapped.rs:45, :232-256) would exceed max_table_size. */
static result_t *run_region(const params_t *p, score_t threshold, int64_t max_number,
                            int64_t max_table_size, int *status) {
    const int64_t n = p->n, m = p->m, stride = m + 1;
    const score_t init = wadd(threshold, 1);
    cell_t *table = (cell_t *)calloc((size_t)((n + 1) * stride), sizeof(cell_t));
    table[0].m = init;
    score_t max_score = init;
    int64_t rows = n + 1 < 100 ? n + 1 : 100, cols = m + 1 < 100 ? m + 1 : 100;
    int64_t i_min_k0 = 0, i_max_k0 = 0, i_min_k1 = 0, i_max_k1 = 0, i_min_k2, i_max_k2;
    *status = 0;
    for (int64_t k = 1; k <= n + m; k++) {
        i_min_k2 = i_min_k1; i_max_k2 = i_max_k1;
        i_min_k1 = i_min_k0; i_max_k1 = i_max_k0;
        i_min_k0 = k; i_max_k0 = 0;
        int64_t i_min = i_min_k1 < i_min_k2 + 1 ? i_min_k1 : i_min_k2 + 1;
        const int64_t lo = k > m ? k - m : 0;
        if (i_min < lo) i_min = lo;
        int64_t i_max = i_max_k1 + 1 > i_max_k2 + 1 ? i_max_k1 + 1 : i_max_k2 + 1;
        if (i_max > n) i_max = n;
        if (i_min > i_max) break;
        const int64_t j_max = k - i_min;
        if (i_max >= rows) {
            if (rows * 2 * cols > max_table_size) { *status = 1; break; }
            rows *= 2;
        }
        if (j_max >= cols) {
            if (rows * cols * 2 > max_table_size) { *status = 1; break; }
            cols *= 2;
        }
        for (int64_t i = i_min; i <= i_max; i++) {
            const int64_t j = k - i;
            const cell_t zero = {0, 0, 0, 0};
            const cell_t *fd = (i > 0 && j > 0) ? &table[(i - 1) * stride + j - 1] : NULL;
            const cell_t f1 = j > 0 ? table[i * stride + j - 1] : zero;
            const cell_t f2 = i > 0 ? table[(i - 1) * stride + j] : zero;
            const score_t sim = fd ? similarity(p, p->code1[i - 1], p->code2[j - 1]) : 0;
            cell_t out = {0, 0, 0, 0};
            int valid = 0;
            if (!p->affine) {
                const score_t a = (fd && fd->m != 0) ? wadd(fd->m, sim) : 0;
                const score_t b = wadd(f1.m, p->gap_open), c = wadd(f2.m, p->gap_open);
                uint8_t dir = 0;
                const score_t sc = p->score_only ? smax(smax(a, b), c) : trace_linear(a, b, c, &dir);
                if (sc >= wsub(max_score, threshold)) {
                    if (sc > max_score) max_score = sc;
                    out.m = sc; out.dir = dir; valid = 1;
                }
            } else {
                const score_t mm = (fd && fd->m != 0) ? wadd(fd->m, sim) : 0;
                const score_t g1m = (fd && fd->g1 != 0) ? wadd(fd->g1, sim) : 0;
                const score_t g2m = (fd && fd->g2 != 0) ? wadd(fd->g2, sim) : 0;
                const score_t mg1 = wadd(f1.m, p->gap_open), g1g1 = wadd(f1.g1, p->gap_ext);
                const score_t mg2 = wadd(f2.m, p->gap_open), g2g2 = wadd(f2.g2, p->gap_ext);
                score_t ms, g1s, g2s;
                uint8_t dir = 0;
                if (p->score_only) {
                    ms = smax(smax(mm, g1m), g2m); g1s = smax(mg1, g1g1); g2s = smax(mg2, g2g2);
                } else {
                    ms = trace_linear(mm, g1m, g2m, &dir);
                    if (mg1 > g1g1) { dir |= A_MG1; g1s = mg1; }
                    else if (mg1 < g1g1) { dir |= A_G1G1; g1s = g1g1; }
                    else { dir |= A_MG1 | A_G1G1; g1s = mg1; }
                    if (mg2 > g2g2) { dir |= A_MG2; g2s = mg2; }
                    else if (mg2 < g2g2) { dir |= A_G2G2; g2s = g2g2; }
                    else { dir |= A_MG2 | A_G2G2; g2s = mg2; }
                }
                score_t req = wsub(max_score, threshold);
                if (ms >= req) {
                    out.m = ms; valid = 1;
                    if (ms > max_score) { max_score = ms; req = wsub(max_score, threshold); }
                }
                if (g1s >= req) {
                    out.g1 = g1s; valid = 1;
                    if (g1s > max_score) { max_score = g1s; req = wsub(max_score, threshold); }
                }
                if (g2s >= req) {
                    out.g2 = g2s; valid = 1;
                    if (g2s > max_score) max_score = g2s;
                }
                out.dir = dir;
            }
            if (valid) {
                if (i_min_k0 == k) i_min_k0 = i;
                i_max_k0 = i;
                table[i * stride + j] = out;
            }
        }
    }
    result_t *res;
    if (*status) {
        res = (result_t *)calloc(1, sizeof(result_t));
    } else {
        /* localgapped.rs:491-511, :621-641: every cell holding the maximum (of M), row-major */
        int64_t n_starts = 0, cap = 16;
        int64_t *starts = (int64_t *)malloc((size_t)cap * sizeof(int64_t));
        int *states = NULL;
        if (!p->score_only) {
            score_t best = INT32_MIN;
            for (int64_t idx = 0; idx < (n + 1) * stride; idx++) {
                const score_t v = table[idx].m;
                if (v > best) { best = v; n_starts = 0; }
                if (v == best) {
                    if (n_starts == cap) { cap *= 2; starts = (int64_t *)realloc(starts, (size_t)cap * sizeof(int64_t)); }
                    starts[n_starts++] = idx;
                }
            }
            states = (int *)calloc((size_t)(n_starts ? n_starts : 1), sizeof(int));
        }
        const int64_t offs[3] = {-stride - 1, -1, -stride};
        res = finish(table, p, wsub(max_score, init), starts, states, n_starts, offs, stride, max_number, 0, 0);
        free(starts); free(states);
    }
    free(table);
    return res;
}

/* ---- exported C interface (ctypes) ------------------------------------------------- */

static void setup(params_t *p, const int64_t *code1, int64_t n, const int64_t *code2,
                  int64_t m, const int32_t *matrix, int64_t n_rows, int64_t n_cols,
                  int32_t match, int32_t mismatch, int affine, int32_t gap_open,
                  int32_t gap_ext, int local, int terminal, int score_only, int banded) {
    p->code1 = code1; p->n = n; p->code2 = code2; p->m = m;
    p->matrix = matrix; p->n_cols = n_cols; p->match = match; p->mismatch = mismatch;
    p->affine = affine; p->gap_open = gap_open; p->gap_ext = gap_ext;
    p->local = local; p->terminal = terminal; p->score_only = score_only;
    score_t mn = scheme_min(matrix, n_rows, n_cols, match, mismatch);
    /* pairwise.rs:585-602, banded.rs:310-314, :437-441 */
    score_t ni = INT32_MIN;
    if (affine) { ni = wsub(wsub(ni, gap_open), gap_ext); if (mn < 0) ni = wsub(ni, mn); }
    else if (banded) { ni = wsub(ni, gap_open); if (mn < 0) ni = wsub(ni, mn); }
    p->neg_inf = ni;
}

void *oracle_align_optimal(const int64_t *code1, int64_t n, const int64_t *code2, int64_t m,
                           const int32_t *matrix, int64_t n_rows, int64_t n_cols,
                           int32_t match, int32_t mismatch, int affine, int32_t gap_open,
                           int32_t gap_ext, int terminal, int local, int score_only,
                           int64_t max_number) {
    params_t p;
    setup(&p, code1, n, code2, m, matrix, n_rows, n_cols, match, mismatch, affine,
          gap_open, gap_ext, local, terminal, score_only, 0);
    return run_optimal(&p, max_number);
}

void *oracle_align_banded(const int64_t *code1, int64_t n, const int64_t *code2, int64_t m,
                          const int32_t *matrix, int64_t n_rows, int64_t n_cols,
                          int32_t match, int32_t mismatch, int affine, int32_t gap_open,
                          int32_t gap_ext, int64_t lower, int64_t upper, int local,
                          int score_only, int64_t max_number) {
    params_t p;
    setup(&p, code1, n, code2, m, matrix, n_rows, n_cols, match, mismatch, affine,
          gap_open, gap_ext, local, 1, score_only, 1);
    return run_banded(&p, lower, upper, max_number);
}

/* localgapped.rs:57-70 `align_region`; *status = 1: "Maximum table size exceeded" */
void *oracle_align_region(const int64_t *code1, int64_t n, const int64_t *code2, int64_t m,
                          const int32_t *matrix, int64_t n_rows, int64_t n_cols,
                          int32_t match, int32_t mismatch, int affine, int32_t gap_open,
                          int32_t gap_ext, int32_t threshold, int score_only, int64_t max_number,
                          int64_t max_table_size, int *status) {
    params_t p;
    setup(&p, code1, n, code2, m, matrix, n_rows, n_cols, match, mismatch, affine,
          gap_open, gap_ext, 0, 1, score_only, 0);
    return run_region(&p, threshold, max_number, max_table_size, status);
}

int32_t oracle_result_score(const void *r) { return ((const result_t *)r)->score; }
int64_t oracle_result_count(const void *r) { return ((const result_t *)r)->n_traces; }
int64_t oracle_result_length(const void *r, int64_t k) { return ((const result_t *)r)->lengths[k]; }
void oracle_result_copy(const void *r, int64_t k, int64_t *out) {
    const result_t *res = (const result_t *)r;
    memcpy(out, res->traces[k], (size_t)res->lengths[k] * 2 * sizeof(int64_t));
}
void oracle_result_free(void *r) {
    result_t *res = (result_t *)r;
    if (!res) return;
    for (int64_t k = 0; k < res->n_traces; k++) free(res->traces[k]);
    free(res->traces); free(res->lengths); free(res);
}

/*
 * Batch driver (max_number = 1 semantics, what every in-tree caller of the
 * reference uses): pairs are independent, so they are spread over worker
 * threads the way the reference's own examples spread them over processes
 * (doc/examples/scripts/sequence/sequencing/genome_assembly.py:329-357).
 * codes: concatenated int64 codes; off1/off2: n_pairs+1 offsets.
 * band: NULL for align_optimal, else 2*n_pairs (lower, upper) already cropped.
 * trace_out: NULL or buffer with room for (len1+len2) rows per pair at
 * trace_off[k] (in rows); trace_len receives L per pair.
 */
typedef struct {
    const int64_t *codes1, *off1, *codes2, *off2;
    int64_t n_pairs;
    const int32_t *matrix;
    int64_t n_rows, n_cols;
    int32_t match, mismatch;
    int affine;
    int32_t gap_open, gap_ext;
    int terminal, local, score_only;
    const int64_t *band;
    int32_t *scores_out;
    int64_t *trace_out;
    const int64_t *trace_off;
    int64_t *trace_len;
    int64_t next; /* shared work counter */
    pthread_mutex_t lock;
} batch_t;

static void *batch_worker(void *arg) {
    batch_t *b = (batch_t *)arg;
    for (;;) {
        pthread_mutex_lock(&b->lock);
        int64_t k0 = b->next;
        b->next += 8;
        pthread_mutex_unlock(&b->lock);
        if (k0 >= b->n_pairs) break;
        int64_t k1 = k0 + 8 < b->n_pairs ? k0 + 8 : b->n_pairs;
        for (int64_t k = k0; k < k1; k++) {
            params_t p;
            setup(&p, b->codes1 + b->off1[k], b->off1[k + 1] - b->off1[k],
                  b->codes2 + b->off2[k], b->off2[k + 1] - b->off2[k], b->matrix,
                  b->n_rows, b->n_cols, b->match, b->mismatch, b->affine, b->gap_open,
                  b->gap_ext, b->local, b->band ? 1 : b->terminal, b->score_only,
                  b->band != NULL);
            result_t *res = b->band ? run_banded(&p, b->band[2 * k], b->band[2 * k + 1], 1)
                                    : run_optimal(&p, 1);
            b->scores_out[k] = res->score;
            if (!b->score_only && b->trace_out) {
                b->trace_len[k] = res->lengths[0];
                memcpy(b->trace_out + 2 * b->trace_off[k], res->traces[0],
                       (size_t)res->lengths[0] * 2 * sizeof(int64_t));
            }
            oracle_result_free(res);
        }
    }
    return NULL;
}

int oracle_batch(const int64_t *codes1, const int64_t *off1, const int64_t *codes2,
                 const int64_t *off2, int64_t n_pairs, const int32_t *matrix,
                 int64_t n_rows, int64_t n_cols, int32_t match, int32_t mismatch,
                 int affine, int32_t gap_open, int32_t gap_ext, int terminal, int local,
                 int score_only, const int64_t *band, int32_t *scores_out,
                 int64_t *trace_out, const int64_t *trace_off, int64_t *trace_len,
                 int n_threads) {
    batch_t b = {codes1, off1, codes2, off2, n_pairs, matrix, n_rows, n_cols, match,
                 mismatch, affine, gap_open, gap_ext, terminal, local, score_only, band,
                 scores_out, trace_out, trace_off, trace_len, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_t *threads = (pthread_t *)malloc((size_t)n_threads * sizeof(pthread_t));
    int started = 0;
    for (int t = 0; t < n_threads - 1; t++)
        if (pthread_create(&threads[started], NULL, batch_worker, &b) == 0) started++;
    batch_worker(&b);
    for (int t = 0; t < started; t++) pthread_join(threads[t], NULL);
    free(threads);
    return 0;
}
