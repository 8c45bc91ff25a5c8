#!/bin/bash
# round 2, call B: compact results, hot/cold tag kernel, look-ahead A/B, ncu of the tag kernel
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
timeout 300 python tools/run_c2.py 1000000 4 > gpurun_out/r2b_run_pf3.log 2>&1
B200ALIGN_LIB=$PWD/build/libb200align_pf2.so timeout 300 python tools/run_c2.py 1000000 4 > gpurun_out/r2b_run_pf2.log 2>&1
B200ALIGN_LIB=$PWD/build/libb200align_pf4.so timeout 300 python tools/run_c2.py 1000000 4 > gpurun_out/r2b_run_pf4.log 2>&1
tail -2 gpurun_out/r2b_run_pf*.log
B200ALIGN_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
cat gpurun_out/r2b_bench.json
grep -E "window|alloc|e2e step" gpurun_out/r2b_bench.err | tail -20
timeout 200 python tools/fuzz_parity.py 60 2000 > gpurun_out/r2b_fuzz.jsonl 2> gpurun_out/r2b_fuzz.err; echo "fuzz rc=$?"
grep -c error gpurun_out/r2b_fuzz.jsonl
# launch list with the SM clock of every launch, then one full capture of the tag fill kernel
timeout 300 python tools/run_c2.py 1000000 3 > gpurun_out/r2b_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2b_launches.csv python tools/run_c2.py 1000000 3 > gpurun_out/r2b_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:interseq_tag_fill -s 3 -c 1 -o gpurun_out/r2b_prof_tag python tools/run_c2.py 1000000 3 > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out/
