#!/usr/bin/env python
"""
bench.py - the hot-path benchmark (BASELINE.json `metric`: GCUPS, affine BLOSUM62).

Workload at every N: BASELINE config 2 per GPU - 10^6 synthetic protein pairs of
~300 aa, global alignment, affine gaps (-10,-1), BLOSUM62, terminal_penalty=True,
score + first-optimal trace (weak scaling: every rank aligns its own 10^6 pairs,
no data-path collective; the only exchange is the max-over-ranks of the timing).

A "step" is one pass of the hot path (DP fill + traceback) over the batch.
  value : whole-job GCUPS with inputs resident in HBM, timed with CUDA events on
          the stream the kernels are launched on (bt_batch_run), max over ranks.
  e2e   : the same metric through the public batch API with pinned HOST buffers:
          H2D of the codes, kernels, D2H of scores + compact traces, every step.
  roofline    : integer-ALU (DPX) roofline of the dominant kernel (the DP fill),
                denominator measured live by bt_microbench on the same device.
  cpu_baseline: the CPU oracle (a port of the reference's algorithm; the reference
                itself is Rust and cannot be built here) on all host threads over
                a bounded sample of the same workload.
`--impl reference` times that CPU path alone, same metric/config.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# background frequencies of the 20 standard amino acids in alphabet order
# ACDEFGHIKLMNPQRSTVWY (values as in the reference's tests/sequence/align/test_statistics.py:11-43)
AA_FREQ = np.array([
    35155, 8669, 24161, 28354, 17367, 33229, 9906, 23161, 25872, 40625,
    10101, 20212, 23435, 19208, 23105, 32070, 26311, 29012, 5990, 14488], dtype=np.float64)
AA_COUNTS = AA_FREQ.astype(np.int64)
AA_TABLE = np.repeat(np.arange(20, dtype=np.uint8), AA_COUNTS)  # 450431-entry inverse CDF
AA_FREQ /= AA_FREQ.sum()


# one metric name for both arms (the reference arm is wall-clock timed on the host: there is no device)
METRIC = "GCUPS (affine BLOSUM62, device-timed)"

def draw_residues(rng, count):
    return AA_TABLE[rng.integers(0, len(AA_TABLE), count, dtype=np.int32)]

GAP = (-10, -1)
ALGO_LANE_INSTR_PER_CELL = 3.0  # affine 3-state recurrence, s16x2 DPX: 6 instr / 2 cells
DIR_BYTES_PER_CELL = 0.5        # 4-bit first-optimal directions
# dram__bytes_read.sum + dram__bytes_write.sum per cell of the dominant fill kernel comes from the
# ncu --set full capture named in this file (written by tools/ncu_traffic.py from the .ncu-rep)
TRAFFIC_FILE = ROOT / "profiles" / "fill_traffic.json"

def make_c2(n_pairs, seed):
    """BASELINE config 2: lengths N(300,30^2) clipped to [200,400]; the second sequence is a
    mutated copy (10 % substitutions, ~2 % indels) for half the pairs, independent otherwise."""
    rng = np.random.default_rng(seed)
    len1 = np.clip(np.rint(rng.normal(300, 30, n_pairs)), 200, 400).astype(np.int64)
    off1 = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(len1, out=off1[1:])
    codes1 = draw_residues(rng, int(off1[-1]))

    homolog = np.zeros(n_pairs, dtype=bool)
    homolog[::2] = True
    # independent partners
    len2 = np.clip(np.rint(rng.normal(300, 30, n_pairs)), 200, 400).astype(np.int64)
    # homologous partners: substitute 10 %, delete 1 %, insert 1 %
    pair_of = np.repeat(np.arange(n_pairs, dtype=np.int32), len1)
    hom_res = homolog[pair_of]
    src = codes1[hom_res]
    src_pair = pair_of[hom_res]
    sub = rng.random(len(src), dtype=np.float32) < 0.10
    src = src.copy()
    src[sub] = draw_residues(rng, int(sub.sum()))
    keep = rng.random(len(src), dtype=np.float32) >= 0.01
    src, src_pair = src[keep], src_pair[keep]
    ins = rng.random(len(src), dtype=np.float32) < 0.01
    n_ins = int(ins.sum())
    ins_codes = draw_residues(rng, n_ins)
    pos = np.nonzero(ins)[0]
    hom_codes = np.insert(src, pos, ins_codes)
    hom_pair = np.insert(src_pair, pos, src_pair[pos])
    hom_len = np.bincount(hom_pair, minlength=n_pairs)
    len2[homolog] = hom_len[homolog]
    off2 = np.zeros(n_pairs + 1, dtype=np.int64)
    np.cumsum(len2, out=off2[1:])
    codes2 = np.empty(int(off2[-1]), dtype=np.uint8)
    pair2 = np.repeat(np.arange(n_pairs, dtype=np.int32), len2)
    hom2 = homolog[pair2]
    codes2[hom2] = hom_codes
    n_ind = int((~hom2).sum())
    codes2[~hom2] = draw_residues(rng, n_ind)
    return codes1, off1, codes2, off2


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, sm_max, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clock, cmax, pw = float(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            inside = t_begin - 0.05 <= t <= t_end + 0.15
            if inside:
                sm.append(clock)
                power.append(pw)
                for name, flag in zip(names, parts[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            sm_max.append(cmax)
        return {
            "sm_mhz": statistics.median(sm) if sm else None,
            "sm_max_mhz": max(sm_max) if sm_max else None,
            "power_w_max": max(power) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu_oracle(codes1, off1, codes2, off2, matrix, target_seconds, threads):
    """Oracle (CPU port of the reference algorithm) on `threads` host threads over a bounded
    prefix of the workload; returns (gcups, n_pairs_used, seconds)."""
    from oracle import oracle
    n_total = len(off1) - 1
    probe = min(n_total, max(64, 8 * threads))

    def run(n):
        a, b = off1[:n + 1], off2[:n + 1]
        t0 = time.perf_counter()
        oracle.batch(codes1[:a[-1]], a, codes2[:b[-1]], b, matrix, GAP, True, False,
                     score_only=False, n_threads=threads)
        dt = time.perf_counter() - t0
        cells = float(np.sum(np.diff(a) * np.diff(b)))
        return cells, dt

    cells, dt = run(probe)
    rate = cells / dt
    n = int(min(n_total, max(probe, target_seconds * rate / (cells / probe))))
    cells, dt = run(n)
    return cells / dt * 1e-9, n, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU (config 2: 10^6)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="all", choices=["all", "none"],
                    help="also run C1 / Cas9 latencies, C3, C4, C5 (JSON key `configs`)")
    ap.add_argument("--c3-db", type=int, default=200_000, help="database sequences of the C3 entry")
    ap.add_argument("--c4-pairs", type=int, default=100_000, help="pairs of the C4 entry (full size: 10^5)")
    ap.add_argument("--c5-seqs", type=int, default=6000, help="proteins of the C5 entry")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    # version banner on file descriptor 1), so everything but the final line goes to stderr.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    config = {
        "workload": "C2: synthetic protein pairs ~300 aa (N(300,30) clipped [200,400], half homologous), "
                    "global affine (-10,-1) BLOSUM62, terminal_penalty=True, score + first-optimal trace",
        "pairs_per_gpu": args.pairs,
        "global_pairs": args.pairs * world,
        "parallelism": f"pairs sharded over {world} GPU(s), no data-path collective",
        "l2": "inputs + direction store far larger than the 126 MB L2, no explicit flush needed",
    }

    import biotite_b200 as bt
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = cpu_threads()
        sample_pairs = min(args.pairs, 200_000)
        codes1, off1, codes2, off2 = make_c2(sample_pairs, seed=0)
        per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
        results = []
        n_used = 0
        for step in range(args.warmup + args.steps):
            gcups, n_used, dt = time_cpu_oracle(codes1, off1, codes2, off2, matrix, per_step, threads)
            if step >= args.warmup:
                results.append((gcups, dt))
        value = statistics.mean(g for g, _ in results)
        ms = statistics.mean(d for _, d in results) * 1e3
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "GCUPS", "cores": threads, "kind": "port",
                             "sample": f"first {n_used} pairs of the C2 workload per step, "
                                       f"oracle/align_oracle.c on {threads} host threads"},
            "e2e": {"value": value, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    from biotite_b200 import _cabi
    numa_cpus = None
    if world > 1:
        # one process per GPU: stay on the GPU's NUMA node (pinned staging, planning threads)
        from biotite_b200.distributed import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local_rank)
    lib = _cabi.load()
    if lib.bt_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    _cabi.check(lib.bt_set_device(local_rank))

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    codes1, off1, codes2, off2 = make_c2(args.pairs, seed=rank)
    # pinned host staging for the end-to-end path
    p_codes1 = _cabi.pinned_empty(codes1.shape, np.uint8); p_codes1[:] = codes1
    p_codes2 = _cabi.pinned_empty(codes2.shape, np.uint8); p_codes2[:] = codes2
    p_off1 = _cabi.pinned_empty(off1.shape, np.int64); p_off1[:] = off1
    p_off2 = _cabi.pinned_empty(off2.shape, np.int64); p_off2[:] = off2
    n = args.pairs
    out = (_cabi.pinned_empty((n,), np.int32), _cabi.pinned_empty((n,), np.int64),
           _cabi.pinned_empty((n, 2), np.int32),
           _cabi.pinned_empty((int(off1[-1] + off2[-1]),), np.uint8))

    # ---- device-resident measurement (`value`) --------------------------------------------
    batch = bt.DeviceBatch(p_codes1, off1, p_codes2, off2, matrix, GAP, True, False)
    cells = batch.cells
    for _ in range(args.warmup):
        batch.run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_begin = time.perf_counter()
    device_ms, fill_ms, back_ms, launches = 0.0, 0.0, 0.0, 0
    for _ in range(args.steps):
        device_ms += batch.run()
        f, b = batch.timings()
        fill_ms += f
        back_ms += b
        launches += batch.launches
    n_windows = batch.windows
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end)
    kernel = batch.kernel
    result = batch.fetch(out)
    checksum = int(result.scores.astype(np.int64).sum())
    batch.close()
    device_s = max_over_ranks(device_ms * 1e-3)
    total_cells = sum_over_ranks(float(cells))
    value = total_cells * args.steps / device_s * 1e-9

    # ---- end to end through the public batch API (host buffers in, host buffers out) ---------
    # results in the compact form (int32 lengths, 2-bit steps): what the reference's FFI returns,
    # the (L, 2) traces of trace.rs:103-130, is rebuilt from it on the host (bt_expand_traces2)
    out_c = bt.compact_out(off1, off2)
    e2e_last = [None]

    def e2e_step():
        # the public one-shot batch API: host arrays in, host arrays out
        t_step = time.perf_counter()
        res = bt.align_batch_arrays(p_codes1, p_off1, p_codes2, p_off2, matrix, GAP, True, False,
                                    out=out_c, compact=True)
        e2e_last[0] = res
        if os.environ.get("B200ALIGN_TRACE"):
            print(f"[bench] rank {rank}: e2e step {1e3 * (time.perf_counter() - t_step):.2f} ms", file=sys.stderr)
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    # every step returns with its results in host memory: the rank's clock stops here, the
    # closing barrier (an NCCL launch that can take tens of ms after idle time) is not work
    e2e_local = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(e2e_local)
    e2e_value = total_cells * args.steps / e2e_s * 1e-9
    h2d = int(sum_over_ranks(float(p_codes1.nbytes + p_codes2.nbytes + off1.nbytes + off2.nbytes + matrix.nbytes)))
    # bytes that actually crossed the link: scores, lengths, first cells and the step words in use
    # (every window copies exactly its compacted words: ceil(L / 16) per pair)
    step_words = int(((out_c[1].astype(np.int64) + 15) // 16).sum())
    d2h = int(sum_over_ranks(float(out_c[0].nbytes + out_c[1].nbytes + out_c[2].nbytes + 4 * step_words)))

    # the same call followed by the expansion to the (L, 2) int64 traces the reference's FFI returns
    # (trace.rs:103-130; bt_expand_traces2 on the host threads): `e2e_expanded`
    barrier()
    t0 = time.perf_counter()
    exp_steps = max(1, min(args.steps, 3))
    for _ in range(exp_steps):
        e2e_step()
        flat, rows_off = e2e_last[0].traces_flat()
    exp_local = time.perf_counter() - t0
    expanded_bytes = int(flat.nbytes)
    del flat
    barrier()
    exp_s = max_over_ranks(exp_local)
    # self-check of the timed workload: sampled pairs of the last end-to-end result against the oracle
    parity_n, parity_ok = 0, True
    if rank == 0:
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_configs
        parity_n, parity_ok = bench_configs.parity_c2(e2e_last[0], codes1, off1, codes2, off2, matrix, cpu_threads())

    # ---- roofline of the dominant kernel (DP fill) ---------------------------------------------
    peak = np.zeros(3)
    for kind in range(3):
        v = _cabi.ctypes.c_double(0.0)
        _cabi.check(lib.bt_microbench(kind, _cabi.ctypes.byref(v)))
        peak[kind] = v.value
    fill_s = fill_ms * 1e-3 / args.steps          # all fill launches of one step (rank 0)
    n_fill_launches = max(1, n_windows)           # one fill launch per window (bt_batch_windows)
    traffic, traffic_src = None, None
    if TRAFFIC_FILE.exists():
        import hashlib
        tf = json.loads(TRAFFIC_FILE.read_text())
        src = ROOT / tf["source"]
        traffic = tf["dram_bytes_per_cell"] * cells / n_fill_launches
        traffic_src = {"file": tf["source"], "kernel": tf["kernel"], "dram_bytes_per_cell": tf["dram_bytes_per_cell"],
                       "sha256": hashlib.sha256(src.read_bytes()).hexdigest() if src.exists() else None}
    achieved = cells * ALGO_LANE_INSTR_PER_CELL / fill_s
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    hbm_peak, hbm_src = 6650.0, "fallback"
    if peaks_file.exists():
        hbm_peak, hbm_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured"
    roofline = {
        "bound": "alu", "kernel": kernel + " fill",
        "achieved": achieved * 1e-12, "peak": peak[0] * 1e-12, "unit": "Tlane-instr/s",
        "frac": achieved / peak[0],
        "peak_source": "bt_microbench kind 0 (VIADDMNMX.S16x2 + VIMNMX3.S16x2, register-only) measured in this run",
        "algorithmic_lane_instr_per_cell": ALGO_LANE_INSTR_PER_CELL,
        "fill_launches_per_step": n_fill_launches, "cells_per_launch": cells / n_fill_launches,
        "ms_per_launch": fill_s * 1e3 / n_fill_launches, "fill_ms_per_step": fill_s * 1e3,
        "traceback_ms_per_step": back_ms / args.steps,
        "fill_share_of_step": fill_ms / device_ms if device_ms else None,
        "peak_s32_dpx": peak[1] * 1e-12, "peak_dual_issue": peak[2] * 1e-12,
        "traffic": traffic,
        "traffic_source": traffic_src,
        "traffic_note": "dram bytes per fill launch = bytes per cell of the ncu --set full capture named in "
                        "traffic_source x the cells of one launch of this run",
        "hbm": {"bound": "hbm", "achieved": cells * DIR_BYTES_PER_CELL / fill_s * 1e-9,
                "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src,
                "frac": cells * DIR_BYTES_PER_CELL / fill_s * 1e-9 / hbm_peak,
                "algorithmic_bytes_per_cell": DIR_BYTES_PER_CELL},
    }

    line = {
        "metric": METRIC, "value": value, "unit": "GCUPS",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": device_s * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s16x2" if "s16" in kernel else "int32", "data": "synthetic",
        "config": config, "kernel": kernel, "pairs_per_s": args.pairs * world * args.steps / device_s,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "pairs_per_s": args.pairs * world * args.steps / e2e_s,
                "result": "scores + first-optimal traces in compact form (int32 lengths, first cell, 2-bit steps)"},
        "e2e_expanded": {"value": total_cells * exp_steps / exp_s * 1e-9, "unit": "GCUPS",
                         "ms_per_step": exp_s * 1e3 / exp_steps, "steps": exp_steps,
                         "expanded_bytes_per_step_rank0": expanded_bytes,
                         "result": "the same call + bt_expand_traces2: (L, 2) int64 traces as the reference's FFI "
                                   "returns them (trace.rs:103-130), in pageable host memory"},
        "parity_checked": parity_n, "parity_ok": parity_ok,
        "gpu_launches": launches, "roofline": roofline, "score_checksum_rank0": checksum,
    }
    if numa_cpus is not None:
        line["config"]["host_binding"] = f"rank 0 bound to the {len(numa_cpus)} CPUs of its GPU's NUMA node"
    if rank == 0 and not args.no_cpu_baseline:
        threads = cpu_threads()
        gcups, n_used, dt = time_cpu_oracle(codes1, off1, codes2, off2, matrix, args.cpu_seconds, threads)
        line["cpu_baseline"] = {
            "value": gcups, "unit": "GCUPS", "cores": threads, "kind": "port",
            "sample": f"first {n_used} pairs of rank 0's C2 workload ({dt:.1f} s), "
                      f"oracle/align_oracle.c on {threads} host threads"}
    if args.configs == "all":
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_configs
        out = out_c = batch = None
        lib.bt_pool_trim()
        configs = {}
        threads = cpu_threads()
        def guarded(name, fn):
            # a failing side configuration must not cost the headline line
            try:
                return fn()
            except Exception as exc:   # noqa: BLE001
                print(f"[bench] config {name} failed: {type(exc).__name__}: {exc}", file=sys.stderr)
                return {"error": f"{type(exc).__name__}: {exc}"}

        # C3 is collective (every rank scores its database slice): an exception on one rank would leave
        # the others in a collective, so it is not guarded per rank
        entry = bench_configs.config_c3(args.c3_db, 100_000, peak, threads, dist, rank, world)
        if rank == 0:
            configs["c3"] = entry
            configs["latency"] = guarded("latency", lambda: bench_configs.single_pair_latencies(threads))
            configs["c4"] = guarded("c4", lambda: bench_configs.config_c4(args.c4_pairs, peak, threads))
            lib.bt_pool_trim()
            configs["c5"] = guarded("c5", lambda: bench_configs.config_c5(args.c5_seqs, peak, threads))
            configs["note"] = ("C3 runs on all ranks (database-sharded), the other entries on rank 0's GPU; "
                               "roofline_frac = lane-instructions of the recurrence per cell x device-timed cells/s / "
                               "the DPX issue rate bt_microbench measured in this run")
            line["configs"] = configs
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
