// Host-side trace utilities of libb200align: expansion of compact step strings
// into the reference's (L,2) trace arrays and co-optimal enumeration over a
// direction table copied back from the GPU.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace bt {

// One enumerated path = table indices of its cells, end-to-start (trace.rs:21-92).
using Path = std::vector<int64_t>;

struct Start {
    int64_t idx;
    int state;
};

struct Step {
    int dir, state;
};

// cell.rs:125-138 (linear) / :219-254 (affine): ordered steps of a cell in a state.
inline int cell_steps(bool affine, uint8_t d, int state, Step out[3]) {
    int k = 0;
    if (!affine) {
        if (d & L_MATCH) out[k++] = {OP_DIAG, 0};
        if (d & L_GAP1) out[k++] = {OP_TO1, 0};
        if (d & L_GAP2) out[k++] = {OP_TO2, 0};
        return k;
    }
    if (state == ST_MATCH) {
        if (d & A_MM) out[k++] = {OP_DIAG, ST_MATCH};
        if (d & A_G1M) out[k++] = {OP_DIAG, ST_GAP1};
        if (d & A_G2M) out[k++] = {OP_DIAG, ST_GAP2};
    } else if (state == ST_GAP1) {
        if (d & A_MG1) out[k++] = {OP_TO1, ST_MATCH};
        if (d & A_G1G1) out[k++] = {OP_TO1, ST_GAP1};
    } else {
        if (d & A_MG2) out[k++] = {OP_TO2, ST_MATCH};
        if (d & A_G2G2) out[k++] = {OP_TO2, ST_GAP2};
    }
    return k;
}

// Emits paths in the order of the reference's recursion (every extra step of a
// tie spawns a branch that is explored, and pushed, before the main path
// continues; trace.rs:71-91) using an explicit stack, so max_number does not
// bound the native stack depth.  Stops once `limit` paths exist: later paths
// would be cut by the reference's final truncate (pairwise.rs:255).
inline void enumerate_from(const uint8_t *dir, bool affine, const int64_t off[3], Start start,
                           int64_t max_number, int64_t limit, std::vector<Path> &out) {
    struct Frame {
        int64_t idx;
        int state;
        Path path;
        Step steps[3];
        int n_steps = 0, next = 0;
        bool at_cell = false;  // steps of the current cell are loaded, branches pending
    };
    int64_t count = 1;  // reset per start (pairwise.rs:241-254)
    std::vector<Frame> stack;
    stack.emplace_back();
    stack.back().idx = start.idx;
    stack.back().state = start.state;
    while (!stack.empty()) {
        if ((int64_t)out.size() >= limit) return;
        Frame &f = stack.back();
        if (!f.at_cell) {
            f.n_steps = cell_steps(affine, (uint8_t)(dir[f.idx] & 0x7f), f.state, f.steps);
            if (f.n_steps == 0) {
                out.push_back(std::move(f.path));
                stack.pop_back();
                continue;
            }
            f.path.push_back(f.idx);
            f.next = 1;
            f.at_cell = true;
        }
        bool spawned = false;
        while (f.next < f.n_steps) {
            const Step s = f.steps[f.next++];
            if (count < max_number) {
                count++;
                Frame child;
                child.idx = f.idx + off[s.dir];
                child.state = s.state;
                child.path = f.path;
                stack.push_back(std::move(child));  // invalidates f
                spawned = true;
                break;
            }
        }
        if (spawned) continue;
        f.idx += off[f.steps[0].dir];
        f.state = f.steps[0].state;
        f.at_cell = false;
    }
}

// trace.rs:103-142.  `rs` = table row stride; banded tables un-straighten the
// column with `lower` (banded.rs:245-251).
inline void path_to_trace(const Path &path, int64_t rs, bool banded, int64_t lower,
                          std::vector<int64_t> &rows) {
    rows.clear();
    int64_t prev0 = INT64_MIN, prev1 = INT64_MIN;
    for (int64_t r = (int64_t)path.size() - 1; r >= 0; r--) {
        const int64_t i = path[r] / rs, j = path[r] % rs;
        const int64_t s0 = i - 1;
        const int64_t s1 = banded ? j + s0 + lower - 1 : j - 1;
        const int64_t v0 = (s0 == prev0) ? -1 : s0;
        const int64_t v1 = (s1 == prev1) ? -1 : s1;
        prev0 = s0;
        prev1 = s1;
        if (v0 == -1 && v1 == -1) continue;
        rows.push_back(v0);
        rows.push_back(v1);
    }
}

// Compact trace (first cell + end-to-start step bytes) -> (L,2) rows.
// Returns the number of rows written (== len unless a degenerate row is dropped).
inline int64_t expand_ops(int64_t len, int32_t first_i, int32_t first_j, const uint8_t *ops,
                          bool banded, int64_t lower, int64_t *out) {
    if (len <= 0) return 0;
    int64_t s0 = (int64_t)first_i - 1;
    int64_t s1 = banded ? (int64_t)first_j + s0 + lower - 1 : (int64_t)first_j - 1;
    int64_t rows = 0;
    // first cell: no predecessor in the path, values are taken as they are
    if (!(s0 == -1 && s1 == -1)) { out[0] = s0; out[1] = s1; rows = 1; }
    for (int64_t k = len - 2; k >= 0; k--) {
        const uint8_t op = ops[k];
        if (op == OP_DIAG) { s0++; s1++; out[2 * rows] = s0; out[2 * rows + 1] = s1; }
        else if (op == OP_TO1) { s1++; out[2 * rows] = -1; out[2 * rows + 1] = s1; }
        else { s0++; out[2 * rows] = s0; out[2 * rows + 1] = -1; }
        rows++;
    }
    return rows;
}

}  // namespace bt
