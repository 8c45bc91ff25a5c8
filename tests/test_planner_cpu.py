"""Invariants of the host-side planner of the packed route (csrc/interseq_kernel.cuh
interseq_plan_host, through the bt_debug_plan test hook; no GPU involved): every pair is placed
exactly once, groups are homogeneous, windows follow the wave size / direction-table budget,
uniform batches keep windows of consecutive pairs, ragged batches are ordered by cost."""
import ctypes

import numpy as np
import pytest

from biotite_b200 import _cabi

WAVE = 148 * 11         # groups per wave of the affine kernels without a device (B200 default): the stripe
WAVE_TAG = 148 * 11     # profiles of the 20 standard letters fit 11 warps per SM (multi-warp blocks)


def _plan(len1, len2, score_only=0, waves=4, affine=True):
    lib = _cabi.load()
    n = len(len1)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    params, _ = _cabi.make_params(np.zeros((24, 24), dtype=np.int32), (-10, -1) if affine else -10, True, False,
                                  False, np.uint8)
    cap = (n + 63) // 64 * 64
    order = np.full(cap, -7, dtype=np.int32)
    win_groups = np.zeros(4096, dtype=np.int64)
    win_bytes = np.zeros(4096, dtype=np.int64)
    n_groups, n_windows, cells = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
    _cabi.check(lib.bt_debug_plan(ctypes.byref(params), _cabi.ptr(off1), _cabi.ptr(off2), n, score_only, waves,
                                  _cabi.ptr(order), cap, ctypes.byref(n_groups), _cabi.ptr(win_groups),
                                  _cabi.ptr(win_bytes), 4096, ctypes.byref(n_windows), ctypes.byref(cells), None))
    g, w = n_groups.value, n_windows.value
    return order[:g * 64].reshape(g, 64), win_groups[:w], win_bytes[:w], cells.value


def _check_permutation(order, n):
    ids = order[order >= 0]
    assert len(ids) == n and np.array_equal(np.sort(ids), np.arange(n))
    assert np.all(order[order < 0] == -1)


def _group_cost(order, len1, len2):
    cost = np.zeros(len(order), dtype=np.int64)
    for g, slots in enumerate(order):
        ids = slots[slots >= 0]
        cost[g] = -(-len1[ids].max() // 16) * len2[ids].max()
    return cost


def test_uniform_batch_keeps_consecutive_windows():
    rng = np.random.default_rng(0)
    n = 400_000
    len1 = np.clip(np.rint(rng.normal(300, 30, n)), 200, 400).astype(np.int64)
    len2 = np.clip(np.rint(rng.normal(300, 30, n)), 200, 400).astype(np.int64)
    order, win_groups, win_bytes, cells = _plan(len1, len2)
    _check_permutation(order, n)
    assert cells == int((len1 * len2).sum())
    assert win_groups.sum() == len(order) and np.all(win_groups <= 4 * WAVE_TAG)
    assert np.all(win_bytes > 0)
    start = 0
    lo = 0
    for count in win_groups:                    # each window holds a run of consecutive pairs
        ids = order[start:start + count]
        ids = ids[ids >= 0]
        assert ids.min() == lo and ids.max() == lo + len(ids) - 1
        lo += len(ids)
        start += count
    # groups are homogeneous: sorted by (len1, len2) inside a window
    spread = np.array([len1[s[s >= 0]].max() - len1[s[s >= 0]].min() for s in order])
    assert np.median(spread) <= 1


def test_ragged_batch_is_ordered_by_cost():
    rng = np.random.default_rng(1)
    n = 300_000
    len1 = np.clip(np.rint(rng.normal(300, 100, n)), 50, 600).astype(np.int64)
    len2 = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
    order, win_groups, win_bytes, cells = _plan(len1, len2, score_only=1, waves=1)
    _check_permutation(order, n)
    assert np.all(win_bytes == 0)               # score only: no direction tables
    cost = _group_cost(order, len1, len2)
    full = cost[:len(cost) - 1]
    assert np.all(np.diff(full) <= 0)           # descending cost over the whole batch
    assert win_groups.sum() == len(order)
    # 4688 groups need three one-wave windows; they are balanced rather than 2 full + remainder
    assert len(win_groups) == -(-len(order) // WAVE) == 3 and win_groups.max() <= WAVE
    assert win_groups.max() - win_groups.min() <= 1


def test_direction_budget_cuts_windows_of_long_pairs():
    n = 64 * 40
    len1 = np.full(n, 7000, dtype=np.int64)
    len2 = np.full(n, 7000, dtype=np.int64)
    order, win_groups, win_bytes, _ = _plan(len1, len2)
    _check_permutation(order, n)
    per_group = -(-7000 // 16) * 7000 * 32 * 4 * 4
    assert np.all(win_bytes == win_groups * per_group)
    assert np.all(win_bytes <= 40 << 30) and len(win_groups) > 1
    assert np.all((win_groups[:-1] + 1) * per_group > 40 << 30)     # every window but the last is full


@pytest.mark.parametrize("n", [64, 65, 127, 1000, 85_249])
def test_odd_pair_counts_and_linear_mode(n):
    rng = np.random.default_rng(n)
    len1 = rng.integers(1, 500, n).astype(np.int64)
    len2 = rng.integers(1, 500, n).astype(np.int64)
    for affine in (True, False):
        order, win_groups, win_bytes, cells = _plan(len1, len2, waves=1, affine=affine)
        _check_permutation(order, n)
        assert len(order) == (n + 63) // 64 and win_groups.sum() == len(order)
        assert cells == int((len1 * len2).sum())
        stripes = np.array([-(-len1[s[s >= 0]].max() // 16) for s in order])
        mmax = np.array([len2[s[s >= 0]].max() for s in order])
        start = 0
        for count, b in zip(win_groups, win_bytes):   # direction words: uint4 (affine) / uint2 (linear) per lane
            sel = slice(start, start + count)
            assert b == int((stripes[sel] * mmax[sel]).sum()) * 32 * (4 if affine else 2) * 4
            start += count


def test_length_range_is_checked():
    with pytest.raises(OverflowError):
        _plan(np.array([9000] * 64), np.array([100] * 64))


@pytest.mark.parametrize("n", [1000, 300_000, 1_200_037])
def test_ragged_order_equals_a_numpy_restatement(n):
    """The batch-wide order is exactly: pairs sorted by (len1, len2) ascending (stable), reversed,
    cut into groups of 64 slots, groups stably sorted by descending cost (threaded counting passes
    over 64-bit records in the library; one thread below 262 144 pairs)."""
    rng = np.random.default_rng(n)
    len1 = np.clip(np.rint(rng.normal(300, 100, n)), 50, 600).astype(np.int64)
    len2 = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n)), 30, 5000).astype(np.int64)
    order, win_groups, _, cells = _plan(len1, len2, score_only=1, waves=1)
    ids = np.lexsort((len2, len1))[::-1].astype(np.int32)
    slots = np.full((n + 63) // 64 * 64, -1, dtype=np.int32)
    slots[:n] = ids
    slots = slots.reshape(-1, 64)
    cost = _group_cost(slots, len1, len2)
    expected = slots[np.argsort(-cost, kind="stable")]
    assert np.array_equal(order, expected)
    assert cells == int((len1 * len2).sum()) and win_groups.sum() == len(order)
