"""
The other BASELINE configurations as bench lines (bench.py adds them to its JSON line under
``configs``): C1 and the Cas9-sized pair of the reference's own benchmark
(benchmarks/sequence/align/benchmark_pairwise.py:22-48) as single-call latencies, C3 (database
search, local affine, score only; the database is sharded over the ranks at N > 1 and the gather of
the score columns is inside the timed region), C4 (banded 10 kb nucleotide pairs, +-128) and C5
(all-vs-all guide-tree score matrix).  Every entry carries device-timed and wall-clock GCUPS, the
integer-ALU roofline fraction of its fill kernel, the CPU oracle's rate on a sample of the same
workload, and `parity_checked`: the number of sampled pairs whose GPU result was compared with
the oracle after the timing (`parity_ok` says whether all of them agreed).
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402
from biotite_b200.batch import PackedSequences  # noqa: E402

GAP = (-10, -1)

# lane-instructions per cell of the recurrences: the packed kernels update two
# 16-bit cells per instruction, the 32-bit kernels one
INSTR = {("interseq_s16x2", True): 3.0, ("interseq_s16x2", False): 1.5,
         ("banded_i32", True): 6.0, ("banded_i32", False): 3.0,
         ("general_i32", True): 6.0, ("general_i32", False): 3.0}


def roofline_frac(kernel, affine, cells_per_s, peaks):
    """Fraction of the measured DPX issue rate (bt_microbench: [s16x2, s32, dual issue])."""
    if "+" in kernel:   # several routes in one job: the packed kernel carries nearly all cells
        names = kernel.split("+")
        kernel = next((k for k in names if "s16" in k), names[0])
    per_cell = INSTR.get((kernel, affine))
    if per_cell is None or not cells_per_s:
        return None
    peak = peaks[0] if "s16" in kernel else peaks[1]
    return cells_per_s * per_cell / peak


def oracle_rate(codes1, off1, codes2, off2, matrix, gap, terminal, local, score_only, bands, threads, seconds=2.0):
    """GCUPS of the CPU oracle on a prefix of the pairs that takes about `seconds`."""
    from oracle import oracle
    n_total = len(off1) - 1
    n = min(n_total, max(threads, 8))
    rate = None
    for _ in range(2):
        a, b = off1[:n + 1], off2[:n + 1]
        bd = None if bands is None else bands[:n]
        t0 = time.perf_counter()
        oracle.batch(codes1[:a[-1]], a, codes2[:b[-1]], b, matrix, gap, terminal, local, score_only=score_only,
                     bands=bd, n_threads=threads)
        dt = time.perf_counter() - t0
        if bands is None:
            cells = float(np.sum(np.diff(a) * np.diff(b)))
        else:
            cells = float(np.sum(np.diff(a) * (bd[:, 1] - bd[:, 0] + 1)))
        rate = cells / dt
        n_next = int(min(n_total, n * max(1.0, seconds / max(dt, 1e-4))))
        if n_next <= n:
            break
        n = n_next
    return rate * 1e-9, n


def _pool(rng, lengths):
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    return PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)


def _gather_pairs(pool1, idx1, pool2, idx2):
    """Concatenated codes + offsets of the listed sequences of two pools (oracle input)."""
    def take(pool, idx):
        lens = pool.lengths[idx]
        off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
        codes = np.concatenate([pool.codes[pool.offsets[i]:pool.offsets[i + 1]] for i in idx]) if len(idx) else np.zeros(0, np.uint8)
        return codes, off
    c1, o1 = take(pool1, idx1)
    c2, o2 = take(pool2, idx2)
    return c1, o1, c2, o2


# ---------------------------------------------------------------------------------------------
def single_pair_latencies(threads):
    """C1 and the Cas9-sized pair through the drop-in single-pair API (wall clock of the call)."""
    from oracle import oracle
    from tests.util import golden
    rng = np.random.default_rng(0)
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    a, b = bt.ProteinSequence(), bt.ProteinSequence()
    a.code = bench.draw_residues(rng, 1368)
    b.code = a.code.copy()
    mut = rng.random(1368) < 0.4
    b.code[mut] = bench.draw_residues(rng, int(mut.sum()))
    b.code = np.delete(b.code, rng.choice(1368, 23, replace=False))
    top, bottom = golden()["readme_avidin_streptavidin"]["gapped"]
    s1, s2 = bt.ProteinSequence(top.replace("-", "")), bt.ProteinSequence(bottom.replace("-", ""))
    m = matrix.score_matrix()

    def timed(fn, reps):
        fn()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            t.append(time.perf_counter() - t0)
        return float(np.median(t)) * 1e3

    out = {}
    cases = [
        ("c1_traced", s1, s2, dict(terminal_penalty=False, max_number=1), None, 100),
        ("c1_score_only", s1, s2, dict(terminal_penalty=False, score_only=True), None, 100),
        ("cas9_traced", a, b, dict(max_number=1), None, 30),
        ("cas9_score_only", a, b, dict(score_only=True), None, 30),
        ("cas9_banded50_traced", a, b, dict(max_number=1), (-50, 50), 30),
    ]
    for name, x, y, kw, band, reps in cases:
        if band is None:
            call = lambda: bt.align_optimal(x, y, matrix, GAP, **kw)  # noqa: E731
            ref = lambda: oracle.align_optimal(x.code, y.code, m, GAP, kw.get("terminal_penalty", True), False,  # noqa: E731
                                               kw.get("score_only", False), 1)
        else:
            call = lambda: bt.align_banded(x, y, matrix, band, GAP, **kw)  # noqa: E731
            ref = lambda: oracle.banded_wrapper(x.code, y.code, m, band, GAP, False, False, 1)  # noqa: E731
        gpu_ms = timed(call, reps)
        cpu_ms = timed(ref, max(3, reps // 10))
        got = call()
        ref_traces, ref_score = ref()
        if kw.get("score_only"):
            ok = int(got) == int(ref_score)
        else:
            ok = got[0].score == ref_score and np.array_equal(got[0].trace, ref_traces[0])
        cells = len(x) * len(y) if band is None else len(x) * 101
        out[name] = {"gpu_ms": gpu_ms, "cpu_port_ms": cpu_ms, "cells": cells,
                     "gcups_wall": cells / gpu_ms * 1e-6, "parity_checked": 1, "parity_ok": bool(ok)}
    # 256 Cas9-sized pairs through the batch API: too few for the inter-sequence kernels' groups of 64
    # pairs per warp, so the intra-pair kernels run them (one block per pair)
    n_pairs = 256
    len1, len2 = rng.integers(1300, 1400, n_pairs), rng.integers(1300, 1400, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1, codes2 = bench.draw_residues(rng, int(off1[-1])), bench.draw_residues(rng, int(off2[-1]))
    for score_only in (False, True):
        with bt.DeviceBatch(codes1, off1, codes2, off2, m, GAP, True, False, score_only=score_only) as batch:
            batch.run()
            ms = min(batch.run() for _ in range(3))
            res = batch.fetch()
            cells, kernel = batch.cells, batch.kernel
        ref_scores, ref_traces = oracle.batch(codes1[:off1[16]], off1[:17], codes2[:off2[16]], off2[:17], m, GAP, True, False,
                                              score_only=score_only, n_threads=threads)
        ok = bool(np.array_equal(res.scores[:16], ref_scores))
        if not score_only:
            ok &= all(np.array_equal(res.trace(k), ref_traces[k]) for k in range(16))
        out["cas9_256_pairs_" + ("score_only" if score_only else "traced")] = {
            "device_ms": ms, "cells": cells, "kernel": kernel, "gcups_device": cells / ms * 1e-6,
            "parity_checked": 16, "parity_ok": ok}
    out["workload"] = ("C1: avidin x streptavidin (README call, semi-global affine BLOSUM62); Cas9-sized 1368 x 1345 "
                       "synthetic pair (benchmark_pairwise.py shapes), global affine, and align_banded +-50; "
                       "median wall clock of the single-pair API call, CPU port on one thread beside it")
    return out


def config_c3(n_db, chunk, peaks, threads, dist=None, rank=0, world=1):
    """1000 queries against a database sample; database-sharded over the ranks at N > 1."""
    rng = np.random.default_rng(0)          # every rank draws the same database
    n_q = 1000
    queries = _pool(rng, np.clip(np.rint(rng.normal(300, 100, n_q)), 50, 600).astype(np.int64))
    database = _pool(rng, np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n_db)), 30, 5000).astype(np.int64))
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    timings = {}
    if world > 1:
        # first use of the gather's point-to-point connections (NCCL sets them up lazily)
        from biotite_b200.distributed import gather_columns
        gather_columns(np.zeros((4, 4), dtype=np.int32))
    # warm the buffer pool / page-locked staging with one small search (a cold first call pays
    # cudaMalloc / cudaHostAlloc of gigabytes once per process, not per search)
    bt.search_scores(queries, database, matrix, GAP, chunk_size=chunk, db_range=(0, min(n_db, 20_000)))
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    if world > 1:
        from biotite_b200.search import search_scores_sharded
        scores = search_scores_sharded(queries, database, matrix, GAP, chunk_size=chunk, timings=timings)
    else:
        scores = bt.search_scores(queries, database, matrix, GAP, chunk_size=chunk, timings=timings)
    wall = time.perf_counter() - t0          # rank 0: includes the gather of all score columns
    cells = float(queries.lengths.sum()) * float(database.lengths.sum())
    device_s = timings["device_ms"] * 1e-3
    if dist is not None:
        import torch
        t = torch.tensor([device_s, wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        device_s, wall = float(t[0]), float(t[1])
    entry = None
    if rank == 0:
        from oracle import oracle
        pick = np.random.default_rng(1)
        qi, di = pick.integers(0, n_q, 1000), pick.integers(0, n_db, 1000)
        c1, o1, c2, o2 = _gather_pairs(queries, qi, database, di)
        ref, _ = oracle.batch(c1, o1, c2, o2, matrix.score_matrix(), GAP, True, True, score_only=True, n_threads=threads)
        ok = bool(np.array_equal(scores[qi, di], ref))
        cpu, n_cpu = oracle_rate(c1, o1, c2, o2, matrix.score_matrix(), GAP, True, True, True, None, threads)
        kernel = timings.get("kernel", "interseq_s16x2")
        entry = {"workload": f"C3: {n_q} queries x {n_db} database sequences (log-normal lengths, median 270, clipped "
                             f"[30, 5000]; chunks of {chunk}), local affine BLOSUM62, score only"
                             + (f"; database sharded over {world} ranks, gather of the score columns timed" if world > 1 else ""),
                 "pairs": n_q * n_db, "cells": cells, "kernel": kernel,
                 "gcups_device": cells / device_s * 1e-9, "gcups_wall": cells / wall * 1e-9, "wall_s": wall,
                 "roofline_frac": roofline_frac(kernel, True, cells / device_s / world, peaks),
                 "cpu_port_gcups": cpu, "cpu_port_sample_pairs": n_cpu, "parity_checked": 1000, "parity_ok": ok}
    return entry


def make_c4(n_pairs, seed):
    """10 kb nucleotide pairs: 5 % substitutions and one indel block of up to 40 columns."""
    rng = np.random.default_rng(seed)
    L = 10_000
    a = np.empty((n_pairs, L), dtype=np.uint8)
    b = np.empty((n_pairs, L), dtype=np.uint8)
    for lo in range(0, n_pairs, 8192):
        hi = min(n_pairs, lo + 8192)
        blk = rng.integers(0, 4, (hi - lo, L), dtype=np.uint8)
        a[lo:hi] = blk
        sub = rng.random((hi - lo, L), dtype=np.float32) < 0.05
        blk = blk.copy()
        blk[sub] = rng.integers(0, 4, int(sub.sum()), dtype=np.uint8)
        b[lo:hi] = blk
    pos = rng.integers(1000, 9000, n_pairs)
    wid = rng.integers(1, 40, n_pairs)
    for k in range(n_pairs):
        p, w = int(pos[k]), int(wid[k])
        b[k, p:L - w] = b[k, p + w:].copy()
    off = np.arange(n_pairs + 1, dtype=np.int64) * L
    bands = np.tile(np.array(bt.banded.crop_band(L, L, (-128, 128)), dtype=np.int64), (n_pairs, 1))
    return a.reshape(-1), off, b.reshape(-1), off, bands


def config_c4(n_pairs, peaks, threads):
    codes1, off1, codes2, off2, bands = make_c4(n_pairs, 0)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()
    t0 = time.perf_counter()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, GAP, True, False, bands) as batch:
        batch.run()
        ms = min(batch.run() for _ in range(2))
        fill, back = batch.timings()
        kernel, cells = batch.kernel, batch.cells
        t1 = time.perf_counter()
        batch.run()
        res = batch.fetch()
        wall = time.perf_counter() - t1
    from oracle import oracle
    pick = np.sort(np.random.default_rng(1).choice(n_pairs, min(n_pairs, 1000), replace=False))
    L = 10_000
    sel = (pick[:, None] * L + np.arange(L)[None, :]).reshape(-1)
    o = np.arange(len(pick) + 1, dtype=np.int64) * L
    ref_scores, ref_traces = oracle.batch(codes1[sel], o, codes2[sel], o, matrix, GAP, True, False, bands=bands[pick],
                                          n_threads=threads)
    ok = bool(np.array_equal(res.scores[pick], ref_scores))
    for q, k in enumerate(pick[:200]):
        ok &= bool(np.array_equal(res.trace(int(k)), ref_traces[q]))
    cpu, n_cpu = oracle_rate(codes1[sel], o, codes2[sel], o, matrix, GAP, True, False, False, bands[pick], threads)
    return {"workload": f"C4: {n_pairs} pairs of 10 kb nucleotide sequences (5 % substitutions, one indel block), "
                        "align_banded +-128, NUC matrix, affine (-10,-1), score + first-optimal trace",
            "pairs": n_pairs, "cells": cells, "kernel": kernel, "gcups_device": cells / ms * 1e-6,
            "fill_ms": fill, "traceback_ms": back,
            "gcups_wall": cells / wall * 1e-9, "wall_s": wall, "setup_s": t1 - t0,
            "wall_note": "run + fetch of scores and step bytes of the resident batch (inputs already uploaded)",
            "roofline_frac": roofline_frac(kernel, True, cells / (fill * 1e-3), peaks),
            "cpu_port_gcups": cpu, "cpu_port_sample_pairs": n_cpu,
            "parity_checked": int(len(pick)), "parity_traces_checked": int(min(len(pick), 200)), "parity_ok": ok}


def config_c5(n_seq, peaks, threads):
    rng = np.random.default_rng(0)
    pool = _pool(rng, np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n_seq)), 30, 5000).astype(np.int64))
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    # two calls: the first pays the device allocations of a cold buffer pool (reported as wall_s_cold)
    t0 = time.perf_counter()
    bt.all_vs_all_scores(pool, matrix, GAP, terminal_penalty=False)
    wall_cold = time.perf_counter() - t0
    timings = {}
    t0 = time.perf_counter()
    scores = bt.all_vs_all_scores(pool, matrix, GAP, terminal_penalty=False, timings=timings)
    wall = time.perf_counter() - t0
    cells = float(timings["cells"])
    device_s = timings["device_ms"] * 1e-3
    from oracle import oracle
    pick = np.random.default_rng(1)
    i, j = pick.integers(0, n_seq, 1000), pick.integers(0, n_seq, 1000)
    c1, o1, c2, o2 = _gather_pairs(pool, i, pool, j)
    ref, _ = oracle.batch(c1, o1, c2, o2, matrix.score_matrix(), GAP, False, False, score_only=True, n_threads=threads)
    ok = bool(np.array_equal(scores[i, j], ref)) and bool(np.array_equal(scores, scores.T))
    cpu, n_cpu = oracle_rate(c1, o1, c2, o2, matrix.score_matrix(), GAP, False, False, True, None, threads)
    kernel = timings.get("kernel", "interseq_s16x2")
    return {"workload": f"C5: all-vs-all of {n_seq} proteins (log-normal lengths, median 270, clipped [30, 5000]), "
                        "semi-global affine BLOSUM62, the guide-tree score matrix of align_multiple",
            "pairs": n_seq * (n_seq + 1) // 2, "cells": cells, "kernel": kernel,
            "gcups_device": cells / device_s * 1e-9, "gcups_wall": cells / wall * 1e-9, "wall_s": wall,
            "wall_s_cold": wall_cold,
            "roofline_frac": roofline_frac(kernel, True, cells / device_s, peaks),
            "cpu_port_gcups": cpu, "cpu_port_sample_pairs": n_cpu, "parity_checked": 1000, "parity_ok": ok}


def parity_c2(result, codes1, off1, codes2, off2, matrix, threads, count=1000):
    """Scores and traces of `count` sampled C2 pairs of an end-to-end result against the oracle."""
    from oracle import oracle
    n = len(off1) - 1
    pick = np.sort(np.random.default_rng(1).choice(n, min(n, count), replace=False))
    l1, l2 = np.diff(off1)[pick], np.diff(off2)[pick]
    o1 = np.concatenate(([0], np.cumsum(l1))).astype(np.int64)
    o2 = np.concatenate(([0], np.cumsum(l2))).astype(np.int64)
    c1 = np.concatenate([codes1[off1[k]:off1[k + 1]] for k in pick])
    c2 = np.concatenate([codes2[off2[k]:off2[k + 1]] for k in pick])
    ref_scores, ref_traces = oracle.batch(c1, o1, c2, o2, matrix, GAP, True, False, n_threads=threads)
    ok = bool(np.array_equal(result.scores[pick], ref_scores))
    for q, k in enumerate(pick):
        ok &= bool(np.array_equal(result.trace(int(k)), ref_traces[q]))
    return int(len(pick)), ok
