"""An independent pin of the oracle's *scores* (the reference's golden files carry alignments, not
scores: SURVEY.md 8c): for tiny sequences every alignment is enumerated as a path of columns and
scored column by column from the definition - substitution scores of the aligned pairs, gap open
for the first column of a gap run and gap extension for the following ones, nothing for terminal
gaps when they are free - and the best of them must be the oracle's score.

The affine model of the reference has three states with G1 <- {M, G1} and G2 <- {M, G2}
(pairwise.rs:608-681): a gap in one sequence directly followed by a gap in the other is not a path,
so the enumeration leaves those out for affine penalties; the linear model has one state and
admits them.  The co-optimal sets of the oracle are checked against the enumeration as well."""
import itertools

import numpy as np
import pytest

from oracle import oracle

DIAG, GAP1, GAP2 = 0, 1, 2      # column kinds: both symbols / gap in sequence 1 / gap in sequence 2


def _paths(n, m, affine):
    """Every alignment of sequences of lengths n and m as a tuple of column kinds."""
    out = []

    def walk(i, j, path):
        if i == n and j == m:
            out.append(tuple(path))
            return
        last = path[-1] if path else None
        if i < n and j < m:
            walk(i + 1, j + 1, path + [DIAG])
        if j < m and not (affine and last == GAP2):
            walk(i, j + 1, path + [GAP1])
        if i < n and not (affine and last == GAP1):
            walk(i + 1, j, path + [GAP2])

    walk(0, 0, [])
    return out


def _column_score(path, a, b, matrix, gap_open, gap_ext, terminal_penalty):
    """Score of one alignment from its columns."""
    has1 = [k != GAP1 for k in path]
    has2 = [k != GAP2 for k in path]
    if terminal_penalty:
        start, stop = 0, len(path)
    else:
        def first(flags):
            return flags.index(True) if True in flags else len(flags)

        def last(flags):
            return len(flags) - flags[::-1].index(True) if True in flags else 0

        start, stop = max(first(has1), first(has2)), min(last(has1), last(has2))
    total, i, j = 0, 0, 0
    for col, kind in enumerate(path):
        if kind == DIAG:
            total += int(matrix[a[i], b[j]])
        elif start <= col < stop:
            total += gap_ext if col > 0 and path[col - 1] == kind else gap_open
        i += kind != GAP1
        j += kind != GAP2
    return total


def _trace_of(path):
    rows, i, j = [], 0, 0
    for kind in path:
        rows.append((i if kind != GAP1 else -1, j if kind != GAP2 else -1))
        i += kind != GAP1
        j += kind != GAP2
    return np.array(rows, dtype=np.int64).reshape(-1, 2)


def _cases(seed, count, longest):
    rng = np.random.default_rng(seed)
    for _ in range(count):
        matrix = rng.integers(-4, 6, (3, 3)).astype(np.int32)
        a = rng.integers(0, 3, rng.integers(0, longest + 1))
        b = rng.integers(0, 3, rng.integers(0, longest + 1))
        yield matrix, a, b


@pytest.mark.parametrize("gap", [-3, -1, 0, (-4, -1), (-2, -2), (-1, -3), (-3, 0), (0, 0)])
@pytest.mark.parametrize("terminal_penalty", [True, False])
def test_global_scores_are_the_best_of_all_alignments(gap, terminal_penalty):
    affine = isinstance(gap, tuple)
    gap_open, gap_ext = gap if affine else (gap, gap)
    for matrix, a, b in _cases(hash((gap, terminal_penalty)) % 2**32, 40, 5):
        paths = _paths(len(a), len(b), affine)
        scores = [_column_score(p, a, b, matrix, gap_open, gap_ext, terminal_penalty) for p in paths]
        best = max(scores)
        traces, score = oracle.align_optimal(a, b, matrix, gap, terminal_penalty, False, False, 10_000)
        assert score == best, (a, b, matrix, gap)
        # the co-optimal set of the oracle is exactly the set of best alignments ...
        expected = {p for p, s in zip(paths, scores) if s == best}
        got = {tuple(map(tuple, t.tolist())) for t in traces}
        if len(a) and len(b):
            assert got == {tuple(map(tuple, _trace_of(p).tolist())) for p in expected}
        # ... and the alignment of max_number = 1 (first-priority step at every tie) is one of them;
        # with one traceback start (linear gaps) it is the one the full enumeration emits last
        # (trace.rs:71-91: branches finish before the main path)
        first, _ = oracle.align_optimal(a, b, matrix, gap, terminal_penalty, False, False, 1)
        assert any(np.array_equal(first[0], t) for t in traces)
        if not affine:
            assert np.array_equal(first[0], traces[-1])


@pytest.mark.parametrize("gap", [-3, -1, (-4, -1), (-2, -2), (-3, 0)])
def test_local_scores_are_the_best_global_score_of_any_two_substrings(gap):
    affine = isinstance(gap, tuple)
    gap_open, gap_ext = gap if affine else (gap, gap)
    for matrix, a, b in _cases(hash(gap) % 2**32, 25, 4):
        best = 0
        for i0, i1 in itertools.combinations(range(len(a) + 1), 2):
            for j0, j1 in itertools.combinations(range(len(b) + 1), 2):
                sa, sb = a[i0:i1], b[j0:j1]
                for p in _paths(len(sa), len(sb), affine):
                    best = max(best, _column_score(p, sa, sb, matrix, gap_open, gap_ext, True))
        _, score = oracle.align_optimal(a, b, matrix, gap, True, True, True, 1)
        assert score == best, (a, b, matrix, gap)


# ---- align_banded: the same enumeration restricted to the band ---------------------------------------
# Path vertices (i, j) = symbols consumed of either sequence; banded.rs:199-214 fills only the cells
# with lower <= j - i <= upper, the cells on the upper and left border keep the start value 0
# (fill_edges, :334-337 / :462-472) and the end may be any cell of the last row or column inside the
# band (:598-638): a semi-global alignment with every gap between start and end charged.

def _band_paths(n, m, lower, upper, affine, local):
    """(start vertex, column kinds) of every alignment inside the band."""
    inside = lambda i, j: 0 <= i <= n and 0 <= j <= m and lower <= j - i <= upper   # noqa: E731
    out = []

    def walk(i, j, start, path):
        if path and (local or i == n or j == m):
            out.append((start, tuple(path)))
        last = path[-1] if path else None
        if inside(i + 1, j + 1):
            walk(i + 1, j + 1, start, path + [DIAG])
        if inside(i, j + 1) and not (affine and last == GAP2):
            walk(i, j + 1, start, path + [GAP1])
        if inside(i + 1, j) and not (affine and last == GAP1):
            walk(i + 1, j, start, path + [GAP2])

    for i in range(n + 1):
        for j in range(m + 1):
            if inside(i, j) and (local or i == 0 or j == 0):
                walk(i, j, (i, j), [])
    return out


def _shown_trace(start, path):
    """
    The trace array the reference prints for a path (trace.rs:103-142): every cell the path enters
    becomes the row (i - 1, j - 1) and a value seen before in its column becomes a gap.  The start
    cell is not part of the path, so a path whose first step is a gap shows that column as the pair
    of the two symbols in front of the cell it enters - a quirk of the reference that a drop-in has
    to keep (it needs a band whose border start makes a leading gap cheaper than a mismatch).
    """
    (i, j), rows = start, []
    seen = (set(), set())
    for kind in path:
        i += kind != GAP1
        j += kind != GAP2
        row = []
        for value, column in zip((i - 1, j - 1), seen):
            row.append(-1 if value in column else value)
            column.add(value)
        if row != [-1, -1]:
            rows.append(tuple(row))
    return tuple(rows)


@pytest.mark.parametrize("gap", [-3, -1, (-4, -1), (-2, -2), (-3, 0)])
@pytest.mark.parametrize("local", [False, True])
def test_banded_scores_are_the_best_of_all_alignments_inside_the_band(gap, local):
    affine = isinstance(gap, tuple)
    gap_open, gap_ext = gap if affine else (gap, gap)
    rng = np.random.default_rng(hash((gap, local)) % 2**32)
    done = 0
    while done < 40:
        matrix = rng.integers(-4, 6, (3, 3)).astype(np.int32)
        n = int(rng.integers(1, 5))
        m = int(rng.integers(n, 6))
        a, b = rng.integers(0, 3, n), rng.integers(0, 3, m)
        band = tuple(sorted(rng.integers(-n - 1, m + 2, 2).tolist()))
        try:
            swapped, lower, upper = oracle.prepare_band(n, m, band)
        except ValueError:
            continue
        assert not swapped
        done += 1
        best, best_set = (0, set()) if local else (None, set())
        for (i0, j0), path in _band_paths(n, m, lower, upper, affine, local):
            sa, sb = a[i0:], b[j0:]
            s = _column_score(path, sa, sb, matrix, gap_open, gap_ext, True)
            if best is None or s > best:
                best, best_set = s, set()
            if s == best:
                best_set.add(_shown_trace((i0, j0), path))
        traces, score = oracle.align_banded(a, b, matrix, gap, lower, upper, local, False, 10_000)
        assert score == best, (a, b, matrix, gap, lower, upper)
        for t in traces:
            if len(t):
                assert tuple(map(tuple, t.tolist())) in best_set, (a, b, matrix, gap, lower, upper)
