"""One packed mode on the C2-shaped workload through a resident batch (driver for ncu captures).
usage: run_mode.py n_pairs local(0/1) terminal(0/1) score_only(0/1)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402

n_pairs, local, term, score_only = int(sys.argv[1]), sys.argv[2] == "1", sys.argv[3] == "1", sys.argv[4] == "1"
codes1, off1, codes2, off2 = bench.make_c2(n_pairs, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), term, local, None, score_only) as b:
    for k in range(3):
        ms = b.run()
        fill, back = b.timings()
        print(f"run {k}: {ms:.3f} ms fill {fill:.3f} traceback {back:.3f} {b.cells / ms * 1e-6:.1f} GCUPS {b.kernel}", flush=True)
