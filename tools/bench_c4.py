"""BASELINE config 4 shape (banded 10 kb nucleotide pairs, band +-128, NUC, affine) and a long-pair
protein batch through whatever kernels they are routed to; device-timed GCUPS."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import biotite_b200 as bt  # noqa: E402

rng = np.random.default_rng(0)
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
L = 10_000
c1, c2, o1, o2 = [], [], [0], [0]
for k in range(n_pairs):
    a = rng.integers(0, 4, L).astype(np.uint8)
    b = a.copy()
    sub = rng.random(L) < 0.05
    b[sub] = rng.integers(0, 4, int(sub.sum()))
    keep = rng.random(L) >= 0.005
    b = b[keep]
    ins = np.sort(rng.choice(len(b), int(0.005 * L), replace=False))
    b = np.insert(b, ins, rng.integers(0, 4, len(ins)).astype(np.uint8))
    if len(b) < len(a):
        a, b = b, a
    c1.append(a); c2.append(b); o1.append(o1[-1] + len(a)); o2.append(o2[-1] + len(b))
codes1, codes2 = np.concatenate(c1), np.concatenate(c2)
off1, off2 = np.array(o1, dtype=np.int64), np.array(o2, dtype=np.int64)
bands = np.array([bt.banded.crop_band(off1[k + 1] - off1[k], off2[k + 1] - off2[k], (-128, 128))
                  for k in range(n_pairs)], dtype=np.int64)
matrix = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()
with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, bands) as b:
    b.run()
    ms = min(b.run() for _ in range(2))
    print(json.dumps({"config": "C4 shape: banded 10 kb, +-128, NUC affine, score+trace", "pairs": n_pairs,
                      "kernel": b.kernel, "ms": ms, "gcups": b.cells / (ms * 1e-3) * 1e-9}))
# long protein pairs (Cas9-sized), global affine
n_long = 256
len1 = rng.integers(1000, 1500, n_long); len2 = rng.integers(1000, 1500, n_long)
off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
codes1 = rng.integers(0, 20, off1[-1]).astype(np.uint8); codes2 = rng.integers(0, 20, off2[-1]).astype(np.uint8)
pm = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
for n_use in (1, 32, 256):
    with bt.DeviceBatch(codes1[:off1[n_use]], off1[:n_use + 1], codes2[:off2[n_use]], off2[:n_use + 1], pm,
                        (-10, -1), True, False) as b:
        b.run()
        ms = min(b.run() for _ in range(2))
        print(json.dumps({"config": "Cas9-sized pairs, global affine, score+trace", "pairs": n_use,
                          "kernel": b.kernel, "ms": ms, "gcups": b.cells / (ms * 1e-3) * 1e-9}))
