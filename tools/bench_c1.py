"""BASELINE config 1: avidin x streptavidin through the single-pair drop-in API (latency per call).
The CPU figure quoted beside it in DESIGN.md (0.33 ms for the port of the reference algorithm)
comes from profiles/r01_configs_c1_c3_c4_c5.jsonl; only tests/, smoke() and bench.py may run the
oracle, so this script times the GPU path alone."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import biotite_b200 as bt  # noqa: E402
from tests.util import golden  # noqa: E402

top, bottom = golden()["readme_avidin_streptavidin"]["gapped"]
s1, s2 = bt.ProteinSequence(top.replace("-", "")), bt.ProteinSequence(bottom.replace("-", ""))
matrix = bt.SubstitutionMatrix.std_protein_matrix()
for name, kwargs in (("max_number=1000 (README call)", {}), ("max_number=1", {"max_number": 1}),
                     ("score_only", {"score_only": True})):
    bt.align_optimal(s1, s2, matrix, gap_penalty=(-10, -1), terminal_penalty=False, **kwargs)
    t = []
    for _ in range(50):
        t0 = time.perf_counter()
        res = bt.align_optimal(s1, s2, matrix, gap_penalty=(-10, -1), terminal_penalty=False, **kwargs)
        t.append(time.perf_counter() - t0)
    print(json.dumps({"config": "C1 avidin x streptavidin, " + name, "gpu_ms_median": float(np.median(t)) * 1e3,
                      "gpu_ms_min": min(t) * 1e3,
                      "n_alignments": res if isinstance(res, int) else len(res)}))
