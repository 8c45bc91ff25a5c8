"""The batch planner on the device (csrc/device_plan.cuh) against the host planner and a numpy
restatement, and the generated batches (pairs enumerated on the GPU) against index arrays and the
oracle.  Reference loops they replace: one align_optimal call per pair of a database scan
(pairwise.py:117-244) and the guide-tree loop of align_multiple (multiple.pyx:322-333)."""
import ctypes
import os

import numpy as np
import pytest

import biotite_b200 as bt
from biotite_b200 import _cabi
from biotite_b200.batch import DeviceBatch, PackedSequences
from oracle import oracle

pytestmark = pytest.mark.gpu


def _device_plan(len1, len2, score_only=1, affine=True):
    lib = _cabi.load()
    n = len(len1)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    groups = (n + 63) // 64
    order = np.full(groups * 64, -7, dtype=np.int32)
    nm = np.zeros(groups, dtype=np.uint32)
    dir_off = np.zeros(groups, dtype=np.int64)
    _cabi.check(lib.bt_debug_device_plan(_cabi.ptr(off1), _cabi.ptr(off2), n, int(score_only), int(affine),
                                         _cabi.ptr(order), _cabi.ptr(nm), _cabi.ptr(dir_off)))
    return order.reshape(groups, 64), nm, dir_off


def _restatement(len1, len2):
    n = len(len1)
    ids = np.lexsort((len2, len1))[::-1].astype(np.int32)
    slots = np.full((n + 63) // 64 * 64, -1, dtype=np.int32)
    slots[:n] = ids
    slots = slots.reshape(-1, 64)
    nmax = np.array([len1[s[s >= 0]].max() for s in slots])
    mmax = np.array([len2[s[s >= 0]].max() for s in slots])
    cost = -(-nmax // 16) * mmax
    perm = np.argsort(-cost, kind="stable")
    return slots[perm], nmax[perm], mmax[perm]


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 85_249, 300_000])
def test_device_plan_equals_the_restatement(n):
    rng = np.random.default_rng(n)
    len1 = np.clip(np.rint(rng.normal(300, 100, n)), 1, 600).astype(np.int64)
    len2 = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 1, 5000).astype(np.int64)
    for score_only, affine in ((1, True), (0, True), (0, False)):
        order, nm, dir_off = _device_plan(len1, len2, score_only, affine)
        expected, nmax, mmax = _restatement(len1, len2)
        assert np.array_equal(order, expected)
        assert np.array_equal(nm >> 16, nmax) and np.array_equal(nm & 0xffff, mmax)
        words = np.zeros(len(nm), dtype=np.int64) if score_only else -(-nmax // 16) * mmax * 32 * (4 if affine else 2)
        assert np.array_equal(dir_off, np.concatenate(([0], np.cumsum(words)[:-1])))


def test_device_plan_equals_the_host_planner():
    """One window of the host planner's batch-wide order (bt_debug_plan, ragged batch) is the device plan."""
    lib = _cabi.load()
    rng = np.random.default_rng(5)
    n = 64 * 700 + 13
    len1 = np.clip(np.rint(rng.normal(300, 100, n)), 50, 600).astype(np.int64)
    len2 = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    params, _ = _cabi.make_params(np.zeros((24, 24), dtype=np.int32), (-10, -1), True, False, False, np.uint8)
    cap = (n + 63) // 64 * 64
    host = np.full(cap, -7, dtype=np.int32)
    n_groups, n_windows = ctypes.c_int64(0), ctypes.c_int64(0)
    _cabi.check(lib.bt_debug_plan(ctypes.byref(params), _cabi.ptr(off1), _cabi.ptr(off2), n, 1, 4, _cabi.ptr(host), cap,
                                  ctypes.byref(n_groups), None, None, 0, ctypes.byref(n_windows), None, None))
    order, _, _ = _device_plan(len1, len2)
    assert n_windows.value == 1
    assert np.array_equal(order.reshape(-1), host)


def _pool(rng, lengths):
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    return PackedSequences(rng.integers(0, 20, int(off[-1])).astype(np.uint8), off, bt.ProteinSequence.alphabet)


@pytest.mark.parametrize("local,terminal", [(True, True), (False, True), (False, False)])
def test_generated_cross_batch_matches_the_oracle(local, terminal):
    rng = np.random.default_rng(2)
    q = _pool(rng, rng.integers(1, 400, 37))
    d = _pool(rng, np.concatenate((rng.integers(1, 900, 600), [0, 1, 3100, 8001])))
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    with DeviceBatch.generated(q.codes, q.offsets, d.codes, d.offsets, matrix, (-10, -1), terminal, local) as batch:
        batch.run()
        assert batch.kernel == "interseq_s16x2"
        assert batch.cells == int(q.lengths.sum()) * int(d.lengths.sum())
        scores = batch.fetch().scores.reshape(len(q), len(d))
    idx1 = np.repeat(np.arange(len(q)), len(d))
    idx2 = np.tile(np.arange(len(d)), len(q))
    c1 = np.concatenate([q.codes[q.offsets[i]:q.offsets[i + 1]] for i in idx1])
    c2 = np.concatenate([d.codes[d.offsets[j]:d.offsets[j + 1]] for j in idx2])
    o1 = np.concatenate(([0], np.cumsum(q.lengths[idx1]))).astype(np.int64)
    o2 = np.concatenate(([0], np.cumsum(d.lengths[idx2]))).astype(np.int64)
    ref, _ = oracle.batch(c1, o1, c2, o2, matrix, (-10, -1), terminal, local, score_only=True)
    assert np.array_equal(scores.reshape(-1), ref)


def test_generated_triangle_matches_index_arrays_and_is_symmetric():
    rng = np.random.default_rng(3)
    # a few sequences above the 16-bit bound of the shorter sequence (2976 for BLOSUM62): both in one pair
    pool = _pool(rng, np.concatenate((np.clip(np.rint(rng.lognormal(np.log(270), 0.7, 300)), 30, 2500).astype(np.int64),
                                      [3000, 3100, 1])))
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    timings = {}
    got = bt.all_vs_all_scores(pool, matrix, (-10, -1), terminal_penalty=False, timings=timings)
    lengths = pool.lengths
    assert timings["cells"] == (int(lengths.sum()) ** 2 + int((lengths ** 2).sum())) // 2
    assert np.array_equal(got, got.T)
    ref = bt.all_vs_all_scores(pool, matrix, (-10, -1), terminal_penalty=False, chunk_limit=0)   # index-array path
    assert np.array_equal(got, ref)
    # a sample against the oracle, including the pair of the two longest sequences
    n = len(pool)
    pick = [(n - 3, n - 2), (0, n - 1), (n - 1, n - 1)] + [tuple(rng.integers(0, n, 2)) for _ in range(40)]
    for i, j in pick:
        a, b = pool.codes[pool.offsets[i]:pool.offsets[i + 1]], pool.codes[pool.offsets[j]:pool.offsets[j + 1]]
        _, score = oracle.align_optimal(a, b, matrix.score_matrix(), (-10, -1), False, False, True, 1)
        assert got[i, j] == score


def test_search_scores_uses_the_generated_batches():
    rng = np.random.default_rng(4)
    q = _pool(rng, rng.integers(50, 400, 20))
    d = _pool(rng, rng.integers(30, 700, 5000))
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    timings = {}
    got = bt.search_scores(q, d, matrix, (-10, -1), chunk_size=1700, timings=timings)
    assert timings["kernel"] == "interseq_s16x2"
    pick = rng.integers(0, len(d), 64)
    for qi in (0, 7, 19):
        a = q.codes[q.offsets[qi]:q.offsets[qi + 1]]
        for j in pick:
            b = d.codes[d.offsets[j]:d.offsets[j + 1]]
            _, score = oracle.align_optimal(a, b, matrix.score_matrix(), (-10, -1), True, True, True, 1)
            assert got[qi, j] == score


def test_streamed_windows_planned_on_the_device_equal_the_host_planned_ones():
    """The one-shot path plans every window on the device; B200ALIGN_PLAN=host keeps the host-thread planner."""
    rng = np.random.default_rng(6)
    n = 3000
    len1, len2 = rng.integers(20, 500, n), rng.integers(20, 500, n)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, 20, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, 20, off2[-1]).astype(np.uint8)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    os.environ["B200ALIGN_ROUTE"] = "packed"
    try:
        dev = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, compact=True)
        os.environ["B200ALIGN_PLAN"] = "host"
        host = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, compact=True)
    finally:
        os.environ.pop("B200ALIGN_PLAN", None)
        os.environ.pop("B200ALIGN_ROUTE", None)
    assert np.array_equal(dev.scores, host.scores) and np.array_equal(dev.trace_len, host.trace_len)
    a, _ = dev.traces_flat()
    b, _ = host.traces_flat()
    assert np.array_equal(a, b)
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False)
    assert np.array_equal(dev.scores, ref_scores)
    assert all(np.array_equal(dev.trace(k), ref_traces[k]) for k in range(0, n, 37))


@pytest.mark.parametrize("k,chunk", [(1, 1000), (10, 333), (64, 5000), (700, 256)])
def test_top_k_selection_on_the_device(k, chunk):
    """search_scores(top_k=...) equals a stable descending sort of the full score matrix (many ties:
    short sequences over four letters), including chunk seams, rows shorter than k and the pairs
    that run on the 32-bit kernels."""
    rng = np.random.default_rng(k)
    q = _pool(rng, rng.integers(5, 60, 33))
    lengths = np.concatenate((rng.integers(1, 80, 600), [0, 9000]))
    d = PackedSequences(rng.integers(0, 4, int(lengths.sum())).astype(np.uint8),
                        np.concatenate(([0], np.cumsum(lengths))).astype(np.int64), bt.ProteinSequence.alphabet)
    q = PackedSequences(q.codes % 4, q.offsets, q.alphabet)
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    full = bt.search_scores(q, d, matrix, (-10, -1), chunk_size=chunk)
    scores, index = bt.search_scores(q, d, matrix, (-10, -1), chunk_size=chunk, top_k=k)
    kk = min(k, len(d))
    order = np.argsort(-full.astype(np.int64), axis=1, kind="stable")[:, :kk]
    assert scores.shape == (len(q), k) and index.shape == (len(q), k)
    assert np.array_equal(index[:, :kk], order)
    assert np.array_equal(scores[:, :kk], np.take_along_axis(full, order, axis=1))
    assert np.all(index[:, kk:] == -1)


def test_generated_batches_with_empty_and_tiny_pools():
    """No pairs, one pair, a pool of empty sequences: the device planner and the 32-bit side batch."""
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    rng = np.random.default_rng(8)
    empty = PackedSequences(np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int64), bt.ProteinSequence.alphabet)
    one = _pool(rng, [37])
    zeros = PackedSequences(np.zeros(0, dtype=np.uint8), np.zeros(4, dtype=np.int64), bt.ProteinSequence.alphabet)
    few = _pool(rng, [12, 300, 1])
    assert bt.search_scores(one, empty, matrix, (-10, -1)).shape == (1, 0)
    assert bt.search_scores(empty, few, matrix, (-10, -1)).shape == (0, 3)
    s, idx = bt.search_scores(one, empty, matrix, (-10, -1), top_k=3)
    assert s.shape == (1, 3) and np.all(idx == -1)
    got = bt.search_scores(few, zeros, matrix, (-10, -1), local=False)          # global against empty sequences
    ref = np.array([[-10 - (n - 1)] * 3 for n in (12, 300, 1)])
    assert np.array_equal(got, ref)
    sq = bt.all_vs_all_scores(one, matrix, (-10, -1))
    a = one.codes
    _, score = oracle.align_optimal(a, a, matrix.score_matrix(), (-10, -1), True, False, True, 1)
    assert sq.shape == (1, 1) and sq[0, 0] == score
