"""BASELINE config 5 at full size: all-vs-all scoring of 20 000 synthetic proteins (log-normal
lengths, median 270, clipped [30, 5000]), semi-global affine BLOSUM62 (-10,-1), the guide-tree
score matrix of align_multiple (2.0001e8 pairs).  Device-timed and wall-clock GCUPS."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402
from biotite_b200.batch import PackedSequences  # noqa: E402

rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
lengths = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
pool = PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)
timings = {}
t0 = time.perf_counter()
scores = bt.all_vs_all_scores(pool, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1), terminal_penalty=False,
                              timings=timings)
wall = time.perf_counter() - t0
print(json.dumps({"config": f"C5: all-vs-all of {n} proteins, semi-global affine, scores", "pairs": n * (n + 1) // 2,
                  "cells": timings["cells"], "device_s": timings["device_ms"] * 1e-3,
                  "gcups_device": timings["cells"] / (timings["device_ms"] * 1e-3) * 1e-9,
                  "wall_s": wall, "gcups_wall": timings["cells"] / wall * 1e-9,
                  "longest": int(lengths.max()), "checksum": int(np.triu(scores).astype(np.int64).sum())}))
