"""BASELINE config 3 through the database-search API: 1000 queries against a database sample
(log-normal lengths, median 270, clipped [30, 5000]), local affine BLOSUM62, score only.
Reports device-timed GCUPS (sum of the batches' CUDA-event times) and wall-clock GCUPS (host
planning of chunk k+1 overlaps the GPU work of chunk k)."""
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
n_q = 1000
n_db = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000


def pool(lengths):
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    return PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)


queries = pool(np.clip(np.rint(rng.normal(300, 100, n_q)), 50, 600).astype(np.int64))
database = pool(np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n_db)), 30, 5000).astype(np.int64))
matrix = bt.SubstitutionMatrix.std_protein_matrix()
timings = {}
t0 = time.perf_counter()
scores = bt.search_scores(queries, database, matrix, (-10, -1), chunk_size=chunk, timings=timings)
wall = time.perf_counter() - t0
print(json.dumps({"config": f"C3: {n_q} queries x {n_db} database sequences (chunks of {chunk}), local affine, score only",
                  "pairs": n_q * n_db, "cells": timings["cells"], "device_s": timings["device_ms"] * 1e-3,
                  "gcups_device": timings["cells"] / (timings["device_ms"] * 1e-3) * 1e-9,
                  "wall_s": wall, "gcups_wall": timings["cells"] / wall * 1e-9,
                  "checksum": int(scores.astype(np.int64).sum())}))
