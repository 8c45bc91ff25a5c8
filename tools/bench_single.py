"""Latency of the single-pair drop-in API on the reference's own benchmark shapes
(benchmarks/sequence/align/benchmark_pairwise.py:22-48: two Cas9-sized proteins, align_optimal
traced and score-only, align_banded +-50) and on BASELINE config 1, plus the device throughput of
256 Cas9-sized pairs through the batch API (general route).  Synthetic sequences of the
Cas9 lengths (1368 and 1345 aa): the timing depends on the lengths only."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402
from tests.util import golden  # noqa: E402


def timed(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return float(np.median(t)) * 1e3, min(t) * 1e3


rng = np.random.default_rng(0)
matrix = bt.SubstitutionMatrix.std_protein_matrix()
a, b = bt.ProteinSequence(), bt.ProteinSequence()
a.code = bench.draw_residues(rng, 1368)
b.code = a.code.copy()
mut = rng.random(1368) < 0.4
b.code[mut] = bench.draw_residues(rng, int(mut.sum()))
b.code = np.delete(b.code, rng.choice(1368, 23, replace=False))
top, bottom = golden()["readme_avidin_streptavidin"]["gapped"]
s1, s2 = bt.ProteinSequence(top.replace("-", "")), bt.ProteinSequence(bottom.replace("-", ""))

cases = [
    ("C1 avidin x streptavidin, max_number=1000 (README call)", lambda: bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=False), 100),
    ("C1 avidin x streptavidin, max_number=1", lambda: bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=False, max_number=1), 100),
    ("C1 avidin x streptavidin, score_only", lambda: bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=False, score_only=True), 100),
    ("Cas9-sized 1368 x 1345, align_optimal affine, max_number=1", lambda: bt.align_optimal(a, b, matrix, (-10, -1), max_number=1), 30),
    ("Cas9-sized 1368 x 1345, align_optimal affine, score_only", lambda: bt.align_optimal(a, b, matrix, (-10, -1), score_only=True), 30),
    ("Cas9-sized 1368 x 1345, align_optimal linear, max_number=1", lambda: bt.align_optimal(a, b, matrix, -10, max_number=1), 30),
    ("Cas9-sized 1368 x 1345, align_optimal local affine, max_number=1", lambda: bt.align_optimal(a, b, matrix, (-10, -1), local=True, max_number=1), 30),
    ("Cas9-sized 1368 x 1345, align_banded +-50 affine, max_number=1", lambda: bt.align_banded(a, b, matrix, (-50, 50), (-10, -1), max_number=1), 30),
    ("Cas9-sized 1368 x 1345, align_banded +-50 affine, score_only", lambda: bt.align_banded(a, b, matrix, (-50, 50), (-10, -1), score_only=True), 30),
]
for name, fn, reps in cases:
    med, best = timed(fn, reps)
    print(json.dumps({"config": name, "gpu_ms_median": med, "gpu_ms_min": best}), flush=True)

# kernel time alone (CUDA events) of one pair through a resident batch
def one_pair(x, y, **kw):
    o1 = np.array([0, len(x.code)], dtype=np.int64)
    o2 = np.array([0, len(y.code)], dtype=np.int64)
    with bt.DeviceBatch(x.code, o1, y.code, o2, matrix.score_matrix(), (-10, -1), kw.get("term", True), False,
                        score_only=kw.get("score_only", False)) as batch:
        batch.run()
        ms = min(batch.run() for _ in range(5))
        fill, back = batch.timings()
        return {"device_ms": ms, "fill_ms": fill, "traceback_ms": back, "kernel": batch.kernel,
                "gcups": batch.cells / ms * 1e-6}


print(json.dumps({"config": "C1 pair, device time, traced", **one_pair(s1, s2, term=False)}), flush=True)
print(json.dumps({"config": "C1 pair, device time, score only", **one_pair(s1, s2, term=False, score_only=True)}), flush=True)
print(json.dumps({"config": "Cas9-sized pair, device time, traced", **one_pair(a, b)}), flush=True)
print(json.dumps({"config": "Cas9-sized pair, device time, score only", **one_pair(a, b, score_only=True)}), flush=True)

# 256 Cas9-sized pairs, batch API (cost model sends them to the general route)
n_pairs = 256
len1 = rng.integers(1300, 1400, n_pairs)
len2 = rng.integers(1300, 1400, n_pairs)
off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
codes1 = bench.draw_residues(rng, int(off1[-1]))
codes2 = bench.draw_residues(rng, int(off2[-1]))
for score_only in (False, True):
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix.score_matrix(), (-10, -1), True, False,
                        score_only=score_only) as batch:
        batch.run()
        ms = min(batch.run() for _ in range(3))
        print(json.dumps({"config": f"256 Cas9-sized pairs, global affine, score_only={score_only}",
                          "kernel": batch.kernel, "device_ms": ms, "gcups": batch.cells / ms * 1e-6}), flush=True)
