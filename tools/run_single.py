"""One pair through a resident batch (general route): device / fill / traceback ms.  Driver for
ncu captures of the wavefront kernel.  usage: run_single.py n m [score_only] [band]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402

n, m = int(sys.argv[1]), int(sys.argv[2])
score_only = len(sys.argv) > 3 and sys.argv[3] == "1"
band = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(0)
c1 = bench.draw_residues(rng, n)
c2 = bench.draw_residues(rng, m)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
o1 = np.array([0, n], dtype=np.int64)
o2 = np.array([0, m], dtype=np.int64)
bands = None if not band else np.array([[-band, band]], dtype=np.int64)
with bt.DeviceBatch(c1, o1, c2, o2, matrix, (-10, -1), True, False, bands=bands, score_only=score_only) as b:
    for k in range(4):
        ms = b.run()
        fill, back = b.timings()
        print(f"run {k}: {ms:.4f} ms fill {fill:.4f} traceback {back:.4f} kernel {b.kernel} {b.cells / ms * 1e-6:.3f} GCUPS", flush=True)
