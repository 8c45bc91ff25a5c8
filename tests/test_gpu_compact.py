"""
Compact result form of the one-shot batch call (``bt_align_batch_compact``): int32 trace
lengths and 2-bit steps must expand to exactly the traces of the oracle, on the pipelined
packed route, the resident routes behind it, and with the tag kernel switched off.
"""

import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle

pytestmark = pytest.mark.gpu

BLOSUM = bt.SubstitutionMatrix.std_protein_matrix
NUC = bt.SubstitutionMatrix.std_nucleotide_matrix


def _batch(rng, n_pairs, lo, hi, alphabet=20, homolog_every=3):
    len1 = rng.integers(lo, hi, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    codes1 = rng.integers(0, alphabet, off1[-1]).astype(np.uint8)
    seqs2 = []
    for k in range(n_pairs):
        if homolog_every and k % homolog_every == 0:
            s = codes1[off1[k]:off1[k + 1]].copy()
            mut = rng.random(len(s)) < 0.15
            s[mut] = rng.integers(0, alphabet, int(mut.sum()))
            keep = rng.random(len(s)) >= 0.03
            s = s[keep] if keep.any() else s[:1]
        else:
            s = rng.integers(0, alphabet, rng.integers(lo, hi)).astype(np.uint8)
        seqs2.append(s)
    off2 = np.concatenate(([0], np.cumsum([len(s) for s in seqs2]))).astype(np.int64)
    return codes1, off1, np.concatenate(seqs2).astype(np.uint8), off2


def _check(res, codes1, off1, codes2, off2, matrix, gap, term, local, ids):
    for k in ids:
        traces, score = oracle.align_optimal(codes1[off1[k]:off1[k + 1]], codes2[off2[k]:off2[k + 1]],
                                             matrix, gap, term, local, False, 1)
        assert score == res.scores[k], f"score of pair {k}"
        assert np.array_equal(traces[0], res.trace(k)), f"trace of pair {k}"


@pytest.mark.parametrize("tag", ["1", "0"])
@pytest.mark.parametrize("gap,local,term", [((-10, -1), False, True), ((-10, -1), False, False),
                                            ((-10, -1), True, True), (-7, False, True), ((-3, 0), False, True)])
def test_compact_matches_oracle_pipelined(gap, local, term, tag, monkeypatch):
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    monkeypatch.setenv("B200ALIGN_TAG", tag)
    monkeypatch.setenv("B200ALIGN_WIN_GROUPS", "3")   # several windows even for a small batch
    rng = np.random.default_rng(7)
    n = 700
    codes1, off1, codes2, off2 = _batch(rng, n, 1, 150)
    matrix = BLOSUM().score_matrix()
    out = bt.compact_out(off1, off2)
    res = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local, out=out, compact=True)
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(n):
        assert np.array_equal(traces[k], ref_traces[k]), f"pair {k}"
    assert res.trace_len.dtype == np.int32
    # only the used words were promised: the end of the used range is reported
    assert 0 < res.ops2_words <= len(res.ops2)
    # the classic form of the same call agrees
    classic = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(classic.scores, res.scores)
    assert np.array_equal(classic.trace_len, res.trace_len)
    assert np.array_equal(classic.first, res.first)


def test_compact_resident_routes(monkeypatch):
    """Small batches (general kernel) and banded batches return the compact form too."""
    rng = np.random.default_rng(11)
    matrix = BLOSUM().score_matrix()
    codes1, off1, codes2, off2 = _batch(rng, 40, 1, 90)
    res = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, compact=True)
    _check(res, codes1, off1, codes2, off2, matrix, (-10, -1), True, False, range(40))
    # banded, through the general / banded kernels
    nuc = NUC().score_matrix()
    codes1, off1, codes2, off2 = _batch(rng, 70, 60, 120, alphabet=4, homolog_every=1)
    len1, len2 = np.diff(off1), np.diff(off2)
    swap = len2 < len1
    assert not swap.all()
    keep = np.nonzero(~swap)[0]
    # keep only pairs with len1 <= len2 (the wrapper's swap is not what is tested here)
    c1 = np.concatenate([codes1[off1[k]:off1[k + 1]] for k in keep])
    c2 = np.concatenate([codes2[off2[k]:off2[k + 1]] for k in keep])
    o1 = np.concatenate(([0], np.cumsum(len1[keep]))).astype(np.int64)
    o2 = np.concatenate(([0], np.cumsum(len2[keep]))).astype(np.int64)
    bands = np.array([bt.banded.crop_band(int(a), int(b), (-12, 14)) for a, b in zip(len1[keep], len2[keep])],
                     dtype=np.int64)
    res = bt.align_batch_arrays(c1, o1, c2, o2, nuc, (-10, -1), True, False, bands=bands, compact=True)
    for k in range(len(keep)):
        traces, score = oracle.align_banded(c1[o1[k]:o1[k + 1]], c2[o2[k]:o2[k + 1]], nuc, (-10, -1),
                                            int(bands[k, 0]), int(bands[k, 1]), False, False, 1)
        assert score == res.scores[k]
        assert np.array_equal(traces[0], res.trace(k))


def test_compact_c2_shape_rescoring():
    """C2-shaped batch at a size the oracle cannot check exhaustively: every expanded trace
    consumes both sequences in order and re-scores to the reported score."""
    rng = np.random.default_rng(3)
    n_pairs = 30000
    codes1, off1, codes2, off2 = _batch(rng, n_pairs, 200, 400, homolog_every=2)
    matrix = BLOSUM().score_matrix()
    out = bt.compact_out(off1, off2)
    res = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, out=out, compact=True)
    flat, rows_off = res.traces_flat()
    for col, off in ((0, off1), (1, off2)):
        present = flat[:, col] != -1
        counts = np.add.reduceat(present.astype(np.int64), rows_off[:-1])
        assert np.array_equal(counts, np.diff(off))
    pair_of_row = np.repeat(np.arange(n_pairs), np.diff(rows_off))
    a, b = flat[:, 0], flat[:, 1]
    both = (a != -1) & (b != -1)
    sub = np.zeros(len(flat), dtype=np.int64)
    sub[both] = matrix[codes1[off1[pair_of_row[both]] + a[both]], codes2[off2[pair_of_row[both]] + b[both]]]
    total = np.add.reduceat(sub, rows_off[:-1])
    for col in (a, b):
        gap = col == -1
        prev = np.concatenate(([False], gap[:-1]))
        prev[rows_off[:-1]] = False
        total += np.add.reduceat(np.where(gap, np.where(prev, -1, -10), 0), rows_off[:-1])
    assert np.array_equal(total, res.scores.astype(np.int64))
    # bytes that crossed the link: far below the byte form's capacity
    assert res.ops2_words * 4 < 0.3 * (off1[-1] + off2[-1])
    _check(res, codes1, off1, codes2, off2, matrix, (-10, -1), True, False, rng.choice(n_pairs, 48, replace=False))


def test_invalid_code_in_packed_batch_is_recoverable(monkeypatch):
    """A code outside the matrix in a pipelined packed batch raises ValueError (the reference
    panics, scoring.rs:63) without poisoning the context: the next call succeeds."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(5)
    codes1, off1, codes2, off2 = _batch(rng, 256, 20, 120)
    matrix = BLOSUM().score_matrix()
    bad = codes2.copy()
    bad[off2[100] + 3] = 200
    for compact in (False, True):
        with pytest.raises(ValueError, match="outside the substitution matrix"):
            bt.align_batch_arrays(codes1, off1, bad, off2, matrix, (-10, -1), True, False, compact=compact)
        bad1 = codes1.copy()
        bad1[off1[17]] = 24
        with pytest.raises(ValueError, match="outside the substitution matrix"):
            bt.align_batch_arrays(bad1, off1, codes2, off2, matrix, (-10, -1), True, True, compact=compact)
        res = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, compact=compact)
        _check(res, codes1, off1, codes2, off2, matrix, (-10, -1), True, False, range(0, 256, 37))


def test_out_arrays_are_validated():
    rng = np.random.default_rng(9)
    codes1, off1, codes2, off2 = _batch(rng, 64, 10, 40)
    matrix = BLOSUM().score_matrix()
    n = 64
    good = (np.empty(n, np.int32), np.empty(n, np.int64), np.empty((n, 2), np.int32),
            np.empty(int(off1[-1] + off2[-1]), np.uint8))
    bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), out=good)
    for k, wrong in enumerate([np.empty(n, np.int64), np.empty(n - 1, np.int64), np.empty((2, n), np.int32).T,
                               np.empty(3, np.uint8)]):
        out = list(good)
        out[k] = wrong
        with pytest.raises(ValueError, match="out:"):
            bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), out=tuple(out))


def test_multi_device_call_matches_single_device():
    """bt_align_batch_multi over every device of the box (and over the same device listed twice:
    two shards, two host threads) against the one-device result and the oracle."""
    from biotite_b200 import _cabi
    lib = _cabi.load()
    rng = np.random.default_rng(21)
    n = 1500
    codes1, off1, codes2, off2 = _batch(rng, n, 30, 200)
    matrix = BLOSUM().score_matrix()
    single = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, compact=True)
    ref = single.traces()
    device_lists = [list(range(lib.bt_device_count())), [0, 0], [0, 0, 0]]
    for devices in device_lists:
        out = bt.compact_out(off1, off2, n_devices=max(2, len(devices)))
        res = bt.align_batch_multi(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, devices=devices, out=out)
        assert np.array_equal(res.scores, single.scores)
        assert np.array_equal(res.trace_len, single.trace_len)
        got = res.traces()
        for k in range(n):
            assert np.array_equal(got[k], ref[k]), (devices, k)
        # shards are contiguous, cover the batch and hold about the same number of DP cells
        shards = res.shards
        assert shards[0] == 0 and shards[-1] == n and np.all(np.diff(shards) >= 0)
        cells = np.diff(off1) * np.diff(off2)
        per = np.add.reduceat(cells, shards[:-1][np.diff(shards) > 0])
        assert per.max() - per.min() <= 2 * cells.max()
    _check(single, codes1, off1, codes2, off2, matrix, (-10, -1), True, False, range(0, n, 97))
    # score only, semi-global, small batch (general route in every shard)
    res = bt.align_batch_multi(codes1[:off1[40]], off1[:41], codes2[:off2[40]], off2[:41], matrix, (-10, -1), False, False,
                               score_only=True, devices=[0, 0])
    ref_scores, _ = oracle.batch(codes1[:off1[40]], off1[:41], codes2[:off2[40]], off2[:41], matrix, (-10, -1), False, False)
    assert np.array_equal(res.scores, ref_scores)
    with pytest.raises(ValueError, match="does not exist"):
        bt.align_batch_multi(codes1, off1, codes2, off2, matrix, (-10, -1), devices=[99])
