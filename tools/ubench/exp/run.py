import sys, numpy as np
sys.path.insert(0,'.')
import bench, biotite_b200 as bt
n=300000
codes1, off1, codes2, off2 = bench.make_c2(n, 0)
matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
for so in (False, True):
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10,-1), True, False, None, so) as b:
        b.run(); b.run(); ms=b.run(); f,t=b.timings()
        print("score_only" if so else "traced", round(f,2), "ms fill", round(b.cells/(f*1e-3)*1e-9), "GCUPS fill", flush=True)
