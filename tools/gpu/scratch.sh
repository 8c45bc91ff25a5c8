#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tools')
import numpy as np
import bench, biotite_b200 as bt
from biotite_b200.batch import DeviceBatch, PackedSequences
rng = np.random.default_rng(0)
n = 6000
lengths = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
pool = PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)
m = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
for rep in range(3):
    t0 = time.perf_counter()
    b = DeviceBatch.generated(pool.codes, pool.offsets, pool.codes, pool.offsets, m, (-10, -1), False, False, kind="triangle")
    t1 = time.perf_counter()
    ms = b.run()
    t2 = time.perf_counter()
    out = b.score_matrix()
    t3 = time.perf_counter()
    b.close()
    t4 = time.perf_counter()
    print(f"rep {rep}: create {t1-t0:.3f} run {t2-t1:.3f} (device {ms*1e-3:.3f}) fetch {t3-t2:.3f} close {t4-t3:.3f} total {t4-t0:.3f}", flush=True)
PY
exit 0
