"""BASELINE config 5 at full size: all-vs-all scoring of 20 000 synthetic proteins (log-normal
lengths, median 270, clipped [30, 5000]), semi-global affine BLOSUM62 (-10,-1): the guide-tree score
matrix of align_multiple (2.0001e8 pairs) as ONE generated batch - the GPU enumerates the triangle,
sorts, groups and packs it (bt_batch_create_generated).  Device-timed and wall-clock GCUPS after a
small warm-up call (CUDA context, module load), 200 sampled pairs against the oracle."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench, biotite_b200 as bt
from biotite_b200.batch import PackedSequences
from oracle import oracle
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
lengths = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
pool = PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)
matrix = bt.SubstitutionMatrix.std_protein_matrix()
small = PackedSequences(pool.codes[:off[200]], off[:201], pool.alphabet)
bt.all_vs_all_scores(small, matrix, (-10, -1), terminal_penalty=False)          # context, module load
timings = {}
t0 = time.perf_counter()
scores = bt.all_vs_all_scores(pool, matrix, (-10, -1), terminal_penalty=False, timings=timings)
wall = time.perf_counter() - t0
pick = np.random.default_rng(1)
ok = bool(np.array_equal(scores, scores.T))
for _ in range(200):
    i, j = int(pick.integers(0, n)), int(pick.integers(0, n))
    a, b = pool.codes[off[i]:off[i + 1]], pool.codes[off[j]:off[j + 1]]
    _, s = oracle.align_optimal(a, b, matrix.score_matrix(), (-10, -1), False, False, True, 1)
    ok &= bool(s == scores[i, j])
out = {"config": f"C5 at full size: all-vs-all of {n} proteins (log-normal lengths, median 270, clipped [30, 5000]), semi-global affine BLOSUM62, one generated batch",
       "pairs": n * (n + 1) // 2, "cells": timings["cells"], "device_s": timings["device_ms"] * 1e-3, "wall_s": wall,
       "gcups_device": timings["cells"] / timings["device_ms"] * 1e-6, "gcups_wall": timings["cells"] / wall * 1e-9,
       "kernel": timings["kernel"], "parity_checked": 200, "parity_ok": ok}
print(json.dumps(out))
open('gpurun_out/r2_c5_full.json', 'w').write(json.dumps(out) + "\n")
