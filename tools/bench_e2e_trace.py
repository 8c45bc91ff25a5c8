"""Timeline of the one-shot path on config 2 (10^6 pairs): run with B200ALIGN_TRACE=1 to get the
per-window start / input-ready / kernels-done / results-returned times of bt_align_batch_compact."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402
from biotite_b200 import _cabi  # noqa: E402

import os  # noqa: E402
_cabi.check(_cabi.load().bt_set_device(int(os.environ.get("LOCAL_RANK", "0"))))   # one process per GPU under torchrun
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
compact = (sys.argv[2] != "classic") if len(sys.argv) > 2 else True
codes1, off1, codes2, off2 = bench.make_c2(n, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
pin = {}
for name, a in (("c1", codes1), ("c2", codes2), ("o1", off1), ("o2", off2)):
    pin[name] = _cabi.pinned_empty(a.shape, a.dtype)
    pin[name][:] = a
if compact:
    out = bt.compact_out(off1, off2)
else:
    out = (_cabi.pinned_empty((n,), np.int32), _cabi.pinned_empty((n,), np.int64), _cabi.pinned_empty((n, 2), np.int32),
           _cabi.pinned_empty((int(off1[-1] + off2[-1]),), np.uint8))
for i in range(5):
    t0 = time.perf_counter()
    bt.align_batch_arrays(pin["c1"], pin["o1"], pin["c2"], pin["o2"], matrix, (-10, -1), True, False, out=out,
                          compact=compact)
    print("e2e ms", round((time.perf_counter() - t0) * 1e3, 2), flush=True)
