"""BASELINE configs 3 and 5 at reduced scale through the indexed batch API (device-timed):
C3: queries x database, local affine, score only; C5: all-vs-all, semi-global affine, scores."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402

rng = np.random.default_rng(0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()


def pool(lengths):
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    return bench.draw_residues(rng, int(off[-1])), off


def db_lengths(count, cap):
    return np.clip(np.rint(rng.lognormal(np.log(270), 0.7, count)), 30, cap).astype(np.int64)


# C3: 1000 queries (300 +- 100) x database sample, local, score only
n_q, n_db = 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 20000
q_codes, q_off = pool(np.clip(np.rint(rng.normal(300, 100, n_q)), 50, 600).astype(np.int64))
d_codes, d_off = pool(db_lengths(n_db, 5000))
idx1 = np.repeat(np.arange(n_q, dtype=np.int32), n_db)
idx2 = np.tile(np.arange(n_db, dtype=np.int32), n_q)
t0 = time.perf_counter()
with bt.DeviceBatch(q_codes, q_off, d_codes, d_off, matrix, (-10, -1), True, True, None, True, idx1=idx1, idx2=idx2) as b:
    t1 = time.perf_counter()
    b.run()
    ms = b.run()
    print(json.dumps({"config": f"C3 shape: {n_q} queries x {n_db} db sequences, local affine, score only",
                      "pairs": len(idx1), "kernel": b.kernel, "ms": ms, "gcups": b.cells / (ms * 1e-3) * 1e-9,
                      "create_s": t1 - t0}))
# C5: all-vs-all of N proteins, semi-global, scores
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
codes, off = pool(db_lengths(n, 2900))   # longer pairs leave the 16-bit range: see all_vs_all_scores
i_idx, j_idx = np.triu_indices(n)
i_idx, j_idx = i_idx.astype(np.int32), j_idx.astype(np.int32)
t0 = time.perf_counter()
with bt.DeviceBatch(codes, off, codes, off, matrix, (-10, -1), False, False, None, True, idx1=i_idx, idx2=j_idx) as b:
    t1 = time.perf_counter()
    b.run()
    ms = b.run()
    print(json.dumps({"config": f"C5 shape: all-vs-all of {n} proteins, semi-global affine, scores",
                      "pairs": len(i_idx), "kernel": b.kernel, "ms": ms, "gcups": b.cells / (ms * 1e-3) * 1e-9,
                      "create_s": t1 - t0}))
