#!/bin/bash
# round 2, call v: modes of the packed kernels, differential fuzzing on the current build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/bench_modes.py > gpurun_out/r2v_modes.jsonl 2>&1; cat gpurun_out/r2v_modes.jsonl
timeout 400 python tools/fuzz_parity.py 240 7000 > gpurun_out/r2v_fuzz.jsonl 2> gpurun_out/r2v_fuzz.err; echo "fuzz rc=$?"
python - <<'PY'
import json, collections
rows = [json.loads(l) for l in open('gpurun_out/r2v_fuzz.jsonl') if l.startswith('{')]
errs = [r for r in rows if 'error' in r]
k = collections.Counter(r.get('kernel', '?') for r in rows)
print(len(rows), 'iterations', len(errs), 'errors', dict(k))
print(rows[0])
for e in errs[:5]: print(e)
PY
exit 0
