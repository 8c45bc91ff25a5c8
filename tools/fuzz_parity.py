"""Time-bounded differential fuzzing of the batched CUDA path against the CPU oracle.

    python tools/fuzz_parity.py [seconds] [first_seed]

Every iteration draws a scoring scheme (BLOSUM62 / NUC / a random small-integer matrix full of
ties), a gap penalty (linear or affine, including extension 0 and open == extension), a mode
(global / semi-global / local), a length law chosen to sit on the kernels' seams (1-residue
sequences, multiples of the 16-row stripe +-1, ragged mixes of tiny and long pairs, n >> m and
m >> n), a pair count that is rarely a multiple of the 64-pair group, and a route (cost model /
packed / general).  Scores and first-optimal traces of the resident batch (traced, score only,
trace statistics) and of the pipelined one-shot entry point must equal the oracle's bit for bit.
Banded iterations draw random cropped bands.  One JSON line per iteration on stdout (a mismatch
adds an "error" key; the seed reproduces it); exit code 1 if any iteration failed.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import biotite_b200 as bt                      # noqa: E402
from biotite_b200.banded import crop_band      # noqa: E402
from oracle import oracle                      # noqa: E402

SEAMS = np.array([1, 2, 15, 16, 17, 31, 32, 33, 47, 48, 49, 63, 64, 65])


def draw_lengths(rng, n_pairs):
    law = rng.integers(0, 6)
    if law == 0:
        return rng.integers(1, 21, n_pairs), rng.integers(1, 21, n_pairs), "tiny"
    if law == 1:
        return rng.choice(SEAMS, n_pairs), rng.choice(SEAMS, n_pairs), "seams"
    if law == 2:
        return rng.integers(50, 300, n_pairs), rng.integers(50, 300, n_pairs), "medium"
    if law == 3:
        long1 = rng.random(n_pairs) < 0.1
        long2 = rng.random(n_pairs) < 0.1
        return (np.where(long1, rng.integers(600, 1200, n_pairs), rng.integers(1, 40, n_pairs)),
                np.where(long2, rng.integers(600, 1200, n_pairs), rng.integers(1, 40, n_pairs)), "ragged")
    if law == 4:
        return rng.integers(1, 12, n_pairs), rng.integers(200, 700, n_pairs), "n<<m"
    return rng.integers(200, 700, n_pairs), rng.integers(1, 12, n_pairs), "n>>m"


def draw_scoring(rng):
    kind = rng.integers(0, 3)
    if kind == 0:
        return bt.SubstitutionMatrix.std_protein_matrix().score_matrix(), 24, "BLOSUM62"
    if kind == 1:
        return bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix(), 15, "NUC"
    k = int(rng.integers(2, 9))
    m = rng.integers(-4, 5, (k, k)).astype(np.int32)
    if rng.random() < 0.5:
        m = ((m + m.T) // 2).astype(np.int32)
    return np.ascontiguousarray(m), k, "random%d" % k


def draw_gap(rng):
    if rng.random() < 0.35:
        return int(-rng.integers(1, 13))
    open_ = int(-rng.integers(1, 16))
    ext = int(-rng.integers(0, -open_ + 1))
    return (open_, ext)


def make_batch(rng, n_pairs, len1, len2, alphabet):
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, alphabet, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, alphabet, off2[-1]).astype(np.uint8)
    homologous = np.flatnonzero(rng.random(n_pairs) < 0.4)
    for k in homologous:
        n = int(min(len1[k], len2[k]))
        shift = int(rng.integers(0, len2[k] - n + 1))
        seg = codes1[off1[k]:off1[k] + n].copy()
        flip = rng.random(n) < 0.12
        seg[flip] = rng.integers(0, alphabet, int(flip.sum()))
        codes2[off2[k] + shift:off2[k] + shift + n] = seg
    return codes1, off1, codes2, off2


def count_gaps(trace, terminal_penalty):
    """(gap-opening, gap-extending) columns of a trace as the reference's guide-tree distance counts them."""
    if not terminal_penalty:
        full = np.nonzero((trace != -1).all(axis=1))[0]
        if full.size == 0:
            return 0, 0
        trace = trace[full[0]:full[-1] + 1]
    opens = exts = 0
    for col in (0, 1):
        gap = trace[:, col] == -1
        prev = np.concatenate(([False], gap[:-1]))
        opens += int(np.count_nonzero(gap & ~prev))
        exts += int(np.count_nonzero(gap & prev))
    return opens, exts


def compare(tag, res, ref_scores, ref_traces, info):
    if not np.array_equal(res.scores, ref_scores):
        bad = np.flatnonzero(res.scores != ref_scores)
        return "%s: %d scores differ, first pair %d (%d vs %d)" % (
            tag, len(bad), bad[0], res.scores[bad[0]], ref_scores[bad[0]])
    if ref_traces is not None:
        traces = res.traces()
        for k, ref in enumerate(ref_traces):
            if not np.array_equal(traces[k], ref):
                return "%s: trace of pair %d differs" % (tag, k)
    return None


def iteration(seed):
    rng = np.random.default_rng(seed)
    banded = rng.random() < 0.25
    route = ["", "packed", "general"][int(rng.choice(3, p=[0.3, 0.55, 0.15]))]
    if route:
        os.environ["B200ALIGN_ROUTE"] = route
    else:
        os.environ.pop("B200ALIGN_ROUTE", None)
    gap = draw_gap(rng)
    info = {"seed": seed, "route": route or "auto", "gap": gap}
    if banded:
        n_pairs = int(rng.integers(64, 400))
        matrix, alphabet, info["matrix"] = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix(), 4, "NUC"
        local = bool(rng.random() < 0.3)
        c1, c2, o1, o2, bands = [], [], [0], [0], []
        while len(bands) < n_pairs:
            n = int(rng.integers(1, 200))
            m = n + int(rng.integers(0, 120))
            a = rng.integers(0, 4, n).astype(np.uint8)
            b = rng.integers(0, 4, m).astype(np.uint8)
            if rng.random() < 0.6:
                shift = int(rng.integers(0, m - n + 1))
                b[shift:shift + n] = a
                flip = rng.random(m) < 0.1
                b[flip] = rng.integers(0, 4, int(flip.sum()))
            lo = int(rng.integers(-n - 5, m + 5))
            hi = lo + int(rng.integers(0, 90))
            try:
                band = crop_band(n, m, (lo, hi))
            except ValueError:
                continue
            c1.append(a); c2.append(b); o1.append(o1[-1] + n); o2.append(o2[-1] + m); bands.append(band)
        codes1, codes2 = np.concatenate(c1), np.concatenate(c2)
        off1, off2 = np.array(o1, dtype=np.int64), np.array(o2, dtype=np.int64)
        bands = np.array(bands, dtype=np.int64)
        info.update(kind="banded", pairs=n_pairs, local=local)
        ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, True, local, bands=bands)
        with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, True, local, bands) as batch:
            info["kernel"] = batch.kernel
            batch.run()
            err = compare("banded traced", batch.fetch(), ref_scores, ref_traces, info)
        if err is None:
            with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, True, local, bands, score_only=True) as batch:
                batch.run()
                err = compare("banded score only", batch.fetch(), ref_scores, None, info)
        return info, err

    n_pairs = int(rng.choice([64, 65, 127, 128, 129, int(rng.integers(64, 3000)), int(rng.integers(3000, 12000))]))
    len1, len2, law = draw_lengths(rng, n_pairs)
    if law in ("ragged", "n<<m", "n>>m", "medium") and n_pairs > 3000:
        n_pairs = 3000
        len1, len2 = len1[:n_pairs], len2[:n_pairs]
    matrix, alphabet, info["matrix"] = draw_scoring(rng)
    mode = int(rng.integers(0, 3))
    local, term = mode == 2, mode != 1
    info.update(kind="optimal", pairs=n_pairs, lengths=law, local=local, terminal_penalty=term)
    codes1, off1, codes2, off2 = make_batch(rng, n_pairs, len1, len2, alphabet)
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local) as batch:
        info["kernel"] = batch.kernel
        batch.run()
        err = compare("resident traced", batch.fetch(), ref_scores, ref_traces, info)
    if err is None:
        with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local, score_only=True) as batch:
            batch.run()
            err = compare("resident score only", batch.fetch(), ref_scores, None, info)
    if err is None:
        one = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local)
        err = compare("one-shot traced", one, ref_scores, ref_traces, info)
    if err is None:
        one = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local, score_only=True)
        err = compare("one-shot score only", one, ref_scores, None, info)
    if err is None and not local:
        # trace statistics: alignment length of the first-optimal trace, no step bytes
        with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local, stats=True) as batch:
            batch.run()
            res = batch.fetch()
        want = np.array([len(t) for t in ref_traces], dtype=np.int64)
        if not np.array_equal(res.scores, ref_scores) or not np.array_equal(res.trace_len, want):
            err = "trace statistics: scores or alignment lengths differ"
        else:
            for k in range(min(n_pairs, 2000)):
                if tuple(res.gap_counts[k]) != count_gaps(ref_traces[k], term):
                    err = "trace statistics: gap counts of pair %d differ" % k
                    break
    return info, err


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0 = time.time()
    done = 0
    kernels = {}
    failed = []
    while time.time() - t0 < budget:
        try:
            info, err = iteration(seed)
        except Exception as exc:   # a library error is a finding too
            info, err = {"seed": seed}, "%s: %s" % (type(exc).__name__, exc)
        kernels[info.get("kernel", "?")] = kernels.get(info.get("kernel", "?"), 0) + 1
        if err is not None:
            info["error"] = err
            failed.append(seed)
        print(json.dumps(info, default=str), flush=True)
        done += 1
        seed += 1
        if len(failed) >= 10:
            break
    print(json.dumps({"fuzz": "MISMATCH" if failed else "ok", "iterations": done, "failed_seeds": failed,
                      "seconds": round(time.time() - t0, 1), "kernels": kernels}), flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
