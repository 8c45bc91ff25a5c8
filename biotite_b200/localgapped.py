"""
``align_local_gapped`` - drop-in for ``biotite.sequence.align.align_local_gapped``
(reference wrapper: ``src/biotite/sequence/align/localgapped.py:57-311``): local gapped
alignment extending from a seed with the X-drop heuristic.

As in the reference, the Python layer splits the problem at the seed into an upstream region
(both sequences reversed) and a downstream region, hands them to the native ``align_region``
- here ``bt_xdrop_regions`` of ``libb200align.so`` (``csrc/xdrop_kernel.cuh``), which extends
all regions of a call concurrently on the GPU - and shifts / combines the traces.
:func:`align_local_gapped_batch` extends many seeds in one call (the gapped stage of a
seed-and-extend search, ``doc/tutorial/sequence/align_heuristic.rst:196-243``).
"""

from __future__ import annotations

import ctypes
import itertools

import numpy as np

from . import _cabi
from .alignment import Alignment
from .pairwise import prepare_scoring

__all__ = ["align_local_gapped", "align_local_gapped_batch"]


def _check_arguments(gap_penalty, max_number, max_table_size, threshold):
    if isinstance(gap_penalty, int):
        if gap_penalty >= 0:
            raise ValueError("Gap penalty must be negative")
    elif isinstance(gap_penalty, tuple):
        if gap_penalty[0] >= 0 or gap_penalty[1] >= 0:
            raise ValueError("Gap penalty must be negative")
    else:
        raise TypeError("Gap penalty must be either integer or tuple")
    if max_number < 1:
        raise ValueError("Maximum number of returned alignments must be at least 1")
    if max_table_size is None:
        max_table_size = np.iinfo(np.int64).max
    elif max_table_size <= 0:
        raise ValueError("Maximum table size must be a positve value")
    if threshold < 0:
        raise ValueError("The threshold value must be a non-negative integer")
    return int(max_table_size)


def _directions(direction):
    if direction == "both":
        return True, True
    if direction == "upstream":
        return True, False
    if direction == "downstream":
        return False, True
    raise ValueError(f"Direction '{direction}' is invalid")


def _native_align_regions(regions1, regions2, scoring, gap_penalty, threshold, max_number,
                          score_only, max_table_size, dtype):
    """
    ``align_region`` (localgapped.rs:57-70) for a list of regions at once.  Returns
    ``(list of trace lists, scores)``; raises ``MemoryError`` like the reference if one region
    exceeds `max_table_size`.
    """
    lib = _cabi.load()
    n = len(regions1)
    off1 = np.zeros(n + 1, dtype=np.int64)
    off2 = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(r) for r in regions1], out=off1[1:])
    np.cumsum([len(r) for r in regions2], out=off2[1:])
    codes1 = np.ascontiguousarray(np.concatenate(regions1) if n else np.zeros(0), dtype)
    codes2 = np.ascontiguousarray(np.concatenate(regions2) if n else np.zeros(0), dtype)
    params, matrix = _cabi.make_params(scoring, gap_penalty, True, False, False, dtype)
    scores = np.zeros(n, dtype=np.int32)
    status = np.zeros(n, dtype=np.int32)
    handles = (ctypes.c_void_p * max(n, 1))()
    _cabi.check(lib.bt_xdrop_regions(
        ctypes.byref(params), _cabi.ptr(codes1), _cabi.ptr(off1), _cabi.ptr(codes2), _cabi.ptr(off2),
        n, _cabi.ptr(matrix), int(threshold), int(max_table_size), int(bool(score_only)),
        int(max_number), _cabi.ptr(scores), _cabi.ptr(status), handles))
    all_traces = []
    try:
        for r in range(n):
            traces = []
            if not score_only and handles[r]:
                for k in range(lib.bt_tracelist_count(handles[r])):
                    length = lib.bt_tracelist_length(handles[r], k)
                    trace = np.empty((length, 2), dtype=np.int64)
                    _cabi.check(lib.bt_tracelist_copy(handles[r], k, _cabi.ptr(trace)))
                    traces.append(trace)
            all_traces.append(traces)
    finally:
        for r in range(n):
            if handles[r]:
                lib.bt_tracelist_free(handles[r])
    if status.any():
        raise MemoryError("Maximum table size exceeded")
    return all_traces, scores


def _seed_score(scoring, code1, code2, seed):
    if isinstance(scoring, tuple):
        return scoring[0] if code1[seed[0]] == code2[seed[1]] else scoring[1]
    return int(scoring[code1[seed[0]], code2[seed[1]]])


def _check_seed(seed, len1, len2):
    s1, s2 = int(seed[0]), int(seed[1])
    if s1 < 0 or s2 < 0:
        raise IndexError("Seed must contain positive indices")
    if s1 >= len1 or s2 >= len2:
        raise IndexError(
            f"Seed {(s1, s2)} is out of bounds for the sequences of length {len1} and {len2}")
    return s1, s2


def _combine(seed, upstream, downstream, up_traces, down_traces, max_number):
    """Shift the region traces to sequence positions and join them at the seed
    (localgapped.py:243-309)."""
    s1, s2 = seed
    ups = []
    for trace in up_traces:
        trace = trace[::-1].copy()
        mask = trace != -1
        trace[mask] *= -1
        trace[mask[:, 0], 0] += s1 - 1
        trace[mask[:, 1], 1] += s2 - 1
        ups.append(trace)
    downs = []
    for trace in down_traces:
        trace = trace.copy()
        trace[trace[:, 0] != -1, 0] += s1 + 1
        trace[trace[:, 1] != -1, 1] += s2 + 1
        downs.append(trace)
    seed_row = np.array([[s1, s2]], dtype=np.int64)
    if upstream and downstream:
        return [np.concatenate([u, seed_row, d])
                for u, d in itertools.islice(itertools.product(ups, downs), max_number)]
    if upstream:
        return [np.concatenate([u, seed_row]) for u in ups]
    if downstream:
        return [np.concatenate([seed_row, d]) for d in downs]
    return [seed_row]


def align_local_gapped(seq1, seq2, matrix, seed, threshold, gap_penalty=-10, max_number=1,
                       direction="both", score_only=False, max_table_size=None):
    """
    Local gapped alignment extending from `seed` into one or both directions until the score
    falls more than `threshold` below the best score found (X-drop).  Same parameters, return
    values and exceptions as the reference's ``align_local_gapped``.

    Returns
    -------
    alignments : list of Alignment, or int if `score_only`
    """
    scoring = prepare_scoring(matrix, seq1, seq2)
    max_table_size = _check_arguments(gap_penalty, max_number, max_table_size, threshold)
    s1, s2 = _check_seed(seed, len(seq1), len(seq2))
    upstream, downstream = _directions(direction)
    if s1 == 0 or s2 == 0:
        upstream = False
    dtype = np.promote_types(seq1.code.dtype, seq2.code.dtype)
    code1 = np.ascontiguousarray(seq1.code, dtype)
    code2 = np.ascontiguousarray(seq2.code, dtype)
    regions1, regions2 = [], []
    if upstream:
        regions1.append(code1[s1 - 1::-1])
        regions2.append(code2[s2 - 1::-1])
    if downstream:
        regions1.append(code1[s1 + 1:])
        regions2.append(code2[s2 + 1:])
    traces, scores = _native_align_regions(regions1, regions2, scoring, gap_penalty, threshold,
                                           max_number, score_only, max_table_size, dtype)
    total = int(scores.sum()) + _seed_score(scoring, code1, code2, (s1, s2))
    if score_only:
        return total
    up = traces[0] if upstream else []
    down = traces[-1] if downstream else []
    combined = _combine((s1, s2), upstream, downstream, up, down, max_number)
    return [Alignment([seq1, seq2], trace, total) for trace in combined]


def align_local_gapped_batch(seqs1, seqs2, matrix, seeds, threshold, gap_penalty=-10,
                             direction="both", score_only=False, max_table_size=None):
    """
    ``align_local_gapped(seqs1[k], seqs2[k], matrix, seeds[k], threshold, gap_penalty,
    max_number=1, direction, score_only, max_table_size)`` for every ``k`` with one native
    call: all upstream and downstream regions are extended concurrently on the GPU.

    Returns a list with one :class:`Alignment` (the first co-optimal one) per seed, or an
    int32 array of scores if `score_only`.
    """
    n = len(seeds)
    if len(seqs1) != n or len(seqs2) != n:
        raise ValueError("seqs1, seqs2 and seeds must have the same length")
    max_table_size = _check_arguments(gap_penalty, 1, max_table_size, threshold)
    want_up, want_down = _directions(direction)
    if n == 0:
        return np.zeros(0, dtype=np.int32) if score_only else []
    scoring = prepare_scoring(matrix, seqs1[0], seqs2[0])
    dtype = np.result_type(*[s.code.dtype for s in seqs1], *[s.code.dtype for s in seqs2])
    regions1, regions2, plan = [], [], []
    for k in range(n):
        if k:
            prepare_scoring(matrix, seqs1[k], seqs2[k])
        s1, s2 = _check_seed(seeds[k], len(seqs1[k]), len(seqs2[k]))
        code1 = np.ascontiguousarray(seqs1[k].code, dtype)
        code2 = np.ascontiguousarray(seqs2[k].code, dtype)
        upstream = want_up and s1 != 0 and s2 != 0
        first = len(regions1)
        if upstream:
            regions1.append(code1[s1 - 1::-1])
            regions2.append(code2[s2 - 1::-1])
        if want_down:
            regions1.append(code1[s1 + 1:])
            regions2.append(code2[s2 + 1:])
        plan.append((s1, s2, upstream, first, _seed_score(scoring, code1, code2, (s1, s2))))
    traces, scores = _native_align_regions(regions1, regions2, scoring, gap_penalty, threshold, 1,
                                           score_only, max_table_size, dtype)
    totals = np.empty(n, dtype=np.int32)
    out = []
    for k, (s1, s2, upstream, first, seed_score) in enumerate(plan):
        count = int(upstream) + int(want_down)
        totals[k] = int(scores[first:first + count].sum()) + seed_score
        if score_only:
            continue
        up = traces[first] if upstream else []
        down = traces[first + count - 1] if want_down else []
        trace = _combine((s1, s2), upstream, want_down, up, down, 1)[0]
        out.append(Alignment([seqs1[k], seqs2[k]], trace, int(totals[k])))
    return totals if score_only else out
