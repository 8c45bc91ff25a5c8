"""Device-timed GCUPS of the packed kernels for the other alignment modes on the C2-shaped
workload (parity-test configs of BASELINE.json, not bench lines)."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
codes1, off1, codes2, off2 = bench.make_c2(n_pairs, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
rows = []
GAPS = {"affine": (-10, -1), "linear": -10}
for name, term, local, score_only in [
    ("global linear, score+trace", True, False, False),
    ("global linear, score only", True, False, True),
    ("local linear, score only", True, True, True),
    ("global affine, score+trace (C2)", True, False, False),
    ("global affine, score only", True, False, True),
    ("semi-global affine, score+trace (C5 shape)", False, False, False),
    ("semi-global affine, score only", False, False, True),
    ("local affine, score only (C3 shape)", True, True, True),
    ("local affine, score+trace", True, True, False),
]:
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, GAPS[name.split()[1].rstrip(',')], term, local, None, score_only) as b:
        for _ in range(2):
            b.run()
        ms = [b.run() for _ in range(3)]
        fill, back = b.timings()
        rows.append({"mode": name, "kernel": b.kernel, "gcups": b.cells / (min(ms) * 1e-3) * 1e-9,
                     "ms": min(ms), "fill_ms": fill, "traceback_ms": back})
        print(json.dumps(rows[-1]))
