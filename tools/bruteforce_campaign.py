"""Time-bounded campaign of tests/test_oracle_bruteforce.py's enumeration against the oracle (CPU only):
random tiny sequences, 4-letter random matrices, gap penalties incl. 0 and |open| < |extension|.

    python tools/bruteforce_campaign.py [seconds per part]

Prints the number of cases per part; stops at the first disagreement."""
import itertools
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import tests.test_oracle_bruteforce as B  # noqa: E402
from oracle import oracle  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(12345)
GAPS = [-3, -1, 0, -5, (-4, -1), (-2, -2), (-1, -3), (-3, 0), (0, 0), (-5, -2), (0, -2)]


def draw():
    gap = GAPS[rng.integers(len(GAPS))]
    affine = isinstance(gap, tuple)
    return rng.integers(-5, 7, (4, 4)).astype(np.int32), gap, affine, (gap if affine else (gap, gap))


def key(trace):
    return tuple(map(tuple, trace.tolist()))


t0, n = time.time(), 0
while time.time() - t0 < budget:                     # align_optimal, global and semi-global: score and co-optimal set
    matrix, gap, affine, (g_open, g_ext) = draw()
    a, b = rng.integers(0, 4, rng.integers(0, 6)), rng.integers(0, 4, rng.integers(0, 6))
    term = bool(rng.integers(2))
    paths = B._paths(len(a), len(b), affine)
    scores = [B._column_score(p, a, b, matrix, g_open, g_ext, term) for p in paths]
    traces, score = oracle.align_optimal(a, b, matrix, gap, term, False, False, 100_000)
    assert score == max(scores), (a, b, matrix, gap, term)
    if len(a) and len(b):
        assert {key(t) for t in traces} == {key(B._trace_of(p)) for p, s in zip(paths, scores) if s == score}
    n += 1
print("align_optimal global / semi-global:", n, "cases")

t0, n = time.time(), 0
while time.time() - t0 < budget:                     # align_banded, semi-global and local: score, printed traces
    matrix, gap, affine, (g_open, g_ext) = draw()
    nn = int(rng.integers(1, 5))
    mm = int(rng.integers(nn, 6))
    a, b = rng.integers(0, 4, nn), rng.integers(0, 4, mm)
    local = bool(rng.integers(2))
    try:
        _, lower, upper = oracle.prepare_band(nn, mm, tuple(sorted(rng.integers(-nn - 1, mm + 2, 2).tolist())))
    except ValueError:
        continue
    best, shown = (0, set()) if local else (None, set())
    for (i0, j0), p in B._band_paths(nn, mm, lower, upper, affine, local):
        s = B._column_score(p, a[i0:], b[j0:], matrix, g_open, g_ext, True)
        if best is None or s > best:
            best, shown = s, set()
        if s == best:
            shown.add(B._shown_trace((i0, j0), p))
    traces, score = oracle.align_banded(a, b, matrix, gap, lower, upper, local, False, 100_000)
    assert score == best, (a, b, matrix, gap, lower, upper, local)
    assert all(key(t) in shown for t in traces if len(t)), (a, b, matrix, gap, lower, upper, local)
    n += 1
print("align_banded:", n, "cases")

t0, n = time.time(), 0
while time.time() - t0 < budget:                     # local: best global score of any two substrings
    matrix, gap, affine, (g_open, g_ext) = draw()
    a, b = rng.integers(0, 4, rng.integers(0, 5)), rng.integers(0, 4, rng.integers(0, 5))
    best = 0
    for i0, i1 in itertools.combinations(range(len(a) + 1), 2):
        for j0, j1 in itertools.combinations(range(len(b) + 1), 2):
            for p in B._paths(i1 - i0, j1 - j0, affine):
                best = max(best, B._column_score(p, a[i0:i1], b[j0:j1], matrix, g_open, g_ext, True))
    assert oracle.align_optimal(a, b, matrix, gap, True, True, True, 1)[1] == best, (a, b, matrix, gap)
    n += 1
print("align_optimal local:", n, "cases")
