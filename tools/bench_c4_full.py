"""BASELINE config 4 at full size: 10^5 pairs of 10 kb nucleotide sequences (5 % substitutions, one\nindel block), band +-128, NUC matrix, affine (-10,-1), semi-global, score + trace; device-timed."""
import sys, json, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import biotite_b200 as bt
rng = np.random.default_rng(0)
n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
L = 10_000
a = rng.integers(0, 4, (n_pairs, L), dtype=np.uint8)
b = a.copy()
sub = rng.random((n_pairs, L)) < 0.05
b[sub] = rng.integers(0, 4, int(sub.sum()), dtype=np.uint8)
# one deletion block and one insertion block per pair keep lengths equal and the optimum in band
codes1 = a.reshape(-1); 
for k in range(0, n_pairs):
    p = int(rng.integers(1000, 9000)); w = int(rng.integers(1, 40))
    b[k, p:L - w] = b[k, p + w:].copy(); b[k, L - w:] = rng.integers(0, 4, w, dtype=np.uint8)
codes2 = b.reshape(-1)
off = (np.arange(n_pairs + 1, dtype=np.int64) * L)
bands = np.tile(np.array(bt.banded.crop_band(L, L, (-128, 128)), dtype=np.int64), (n_pairs, 1))
matrix = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()
with bt.DeviceBatch(codes1, off, codes2, off, matrix, (-10, -1), True, False, bands) as bb:
    bb.run()
    ms = min(bb.run() for _ in range(2)); f, t = bb.timings()
    print(json.dumps({"config": "C4 shape", "pairs": n_pairs, "kernel": bb.kernel, "ms": ms, "fill_ms": f, "tb_ms": t, "gcups": bb.cells / (ms * 1e-3) * 1e-9}))
