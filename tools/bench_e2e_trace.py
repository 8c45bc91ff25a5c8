"""Timeline of the one-shot path on config 2 (10^6 pairs): run with B200ALIGN_TRACE=1 to get the\nper-window start / input-ready / kernels-done / results-returned times of bt_align_batch."""
import sys, time, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import bench, biotite_b200 as bt
from biotite_b200 import _cabi
n=1000000
codes1, off1, codes2, off2 = bench.make_c2(n, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
p1=_cabi.pinned_empty(codes1.shape,np.uint8); p1[:]=codes1
p2=_cabi.pinned_empty(codes2.shape,np.uint8); p2[:]=codes2
out=(_cabi.pinned_empty((n,),np.int32),_cabi.pinned_empty((n,),np.int64),_cabi.pinned_empty((n,2),np.int32),_cabi.pinned_empty((int(off1[-1]+off2[-1]),),np.uint8))
for i in range(5):
    t0=time.perf_counter()
    bt.align_batch_arrays(p1, off1, p2, off2, matrix, (-10,-1), True, False, out=out)
    print("e2e ms", round((time.perf_counter()-t0)*1e3,2), flush=True)
