"""C2 workload (bench.make_c2) through a resident batch: prints device / fill / traceback ms of
`runs` runs.  Small driver for ncu captures and A/B builds (B200ALIGN_LIB)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
runs = int(sys.argv[2]) if len(sys.argv) > 2 else 5
codes1, off1, codes2, off2 = bench.make_c2(n_pairs, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False) as b:
    for k in range(runs):
        ms = b.run()
        fill, back = b.timings()
        print(f"run {k}: {ms:.3f} ms  fill {fill:.3f}  traceback {back:.3f}  "
              f"{b.cells / ms * 1e-6:.1f} GCUPS  kernel {b.kernel}", flush=True)
