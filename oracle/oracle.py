"""
ctypes binding of the CPU oracle (``oracle/align_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The
product package ``biotite_b200`` never imports this module.

The two array-level functions mirror the reference's native entry points
(``src/rust/sequence/align/pairwise.rs:57-70`` and ``banded.rs:63-77``): same
argument order and meaning, same return value ``(list of (L,2) int64 traces,
score)``.  :func:`banded_wrapper` additionally restates the Python-side
preparation of ``src/biotite/sequence/align/banded.py:215-261`` (swap so that
the first sequence is the shorter one, band negation, matrix transposition,
band cropping, trace flip) on plain arrays, so the product's own wrapper can be
checked against an independent restatement.

Parity status: pinned.  ``tests/test_oracle_golden.py`` checks this oracle
against the reference's 540 golden alignments, its 7 known-answer cases and its
docstring outputs (fixtures in ``tests/golden/``).
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle.so"

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> Path:
    """Compile ``liboracle.so`` with gcc if missing or stale."""
    src = _HERE / "align_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "liboracle.so"], check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        common = [_i64p, ctypes.c_int64, _i64p, ctypes.c_int64, _i32p, ctypes.c_int64,
                  ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int,
                  ctypes.c_int32, ctypes.c_int32]
        lib.oracle_align_optimal.restype = ctypes.c_void_p
        lib.oracle_align_optimal.argtypes = common + [
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64]
        lib.oracle_align_banded.restype = ctypes.c_void_p
        lib.oracle_align_banded.argtypes = common + [
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64]
        lib.oracle_align_region.restype = ctypes.c_void_p
        lib.oracle_align_region.argtypes = common + [
            ctypes.c_int32, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
        lib.oracle_result_score.restype = ctypes.c_int32
        lib.oracle_result_score.argtypes = [ctypes.c_void_p]
        lib.oracle_result_count.restype = ctypes.c_int64
        lib.oracle_result_count.argtypes = [ctypes.c_void_p]
        lib.oracle_result_length.restype = ctypes.c_int64
        lib.oracle_result_length.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        lib.oracle_result_copy.restype = None
        lib.oracle_result_copy.argtypes = [ctypes.c_void_p, ctypes.c_int64, _i64p]
        lib.oracle_result_free.restype = None
        lib.oracle_result_free.argtypes = [ctypes.c_void_p]
        lib.oracle_batch.restype = ctypes.c_int
        lib.oracle_batch.argtypes = [
            _i64p, _i64p, _i64p, _i64p, ctypes.c_int64, _i32p, ctypes.c_int64,
            ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int, ctypes.c_int32,
            ctypes.c_int32, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64p, _i32p, _i64p,
            _i64p, _i64p, ctypes.c_int]
        _lib = lib
    return _lib


def _ptr(arr, typ):
    return arr.ctypes.data_as(typ) if arr is not None else None


def _scoring_args(scoring):
    if isinstance(scoring, tuple):
        return None, 0, 0, int(scoring[0]), int(scoring[1])
    mat = np.ascontiguousarray(scoring, dtype=np.int32)
    return mat, mat.shape[0], mat.shape[1], 0, 0


def _gap_args(gap_penalty):
    if isinstance(gap_penalty, tuple):
        return 1, int(gap_penalty[0]), int(gap_penalty[1])
    return 0, int(gap_penalty), int(gap_penalty)


def _collect(lib, handle):
    try:
        score = int(lib.oracle_result_score(handle))
        traces = []
        for k in range(lib.oracle_result_count(handle)):
            length = lib.oracle_result_length(handle, k)
            out = np.empty((length, 2), dtype=np.int64)
            if length:
                lib.oracle_result_copy(handle, k, _ptr(out, _i64p))
            traces.append(out)
        return traces, score
    finally:
        lib.oracle_result_free(handle)


def align_optimal(code1, code2, scoring, gap_penalty, terminal_penalty, local,
                  score_only, max_number):
    """Same contract as the reference's native ``align_optimal``."""
    lib = _load()
    c1 = np.ascontiguousarray(code1, dtype=np.int64)
    c2 = np.ascontiguousarray(code2, dtype=np.int64)
    mat, n_rows, n_cols, match, mismatch = _scoring_args(scoring)
    affine, gap_open, gap_ext = _gap_args(gap_penalty)
    handle = lib.oracle_align_optimal(
        _ptr(c1, _i64p), len(c1), _ptr(c2, _i64p), len(c2), _ptr(mat, _i32p), n_rows,
        n_cols, match, mismatch, affine, gap_open, gap_ext, int(terminal_penalty),
        int(local), int(score_only), int(max_number))
    return _collect(lib, handle)


def align_banded(code1, code2, scoring, gap_penalty, lower_diag, upper_diag, local,
                 score_only, max_number):
    """Same contract as the reference's native ``align_banded`` (band already cropped)."""
    lib = _load()
    c1 = np.ascontiguousarray(code1, dtype=np.int64)
    c2 = np.ascontiguousarray(code2, dtype=np.int64)
    mat, n_rows, n_cols, match, mismatch = _scoring_args(scoring)
    affine, gap_open, gap_ext = _gap_args(gap_penalty)
    handle = lib.oracle_align_banded(
        _ptr(c1, _i64p), len(c1), _ptr(c2, _i64p), len(c2), _ptr(mat, _i32p), n_rows,
        n_cols, match, mismatch, affine, gap_open, gap_ext, int(lower_diag),
        int(upper_diag), int(local), int(score_only), int(max_number))
    return _collect(lib, handle)


def align_region(code1, code2, scoring, gap_penalty, threshold, max_number, score_only,
                 max_table_size):
    """Same contract as the reference's native ``align_region`` (localgapped.rs:57-70)."""
    lib = _load()
    c1 = np.ascontiguousarray(code1, dtype=np.int64)
    c2 = np.ascontiguousarray(code2, dtype=np.int64)
    mat, n_rows, n_cols, match, mismatch = _scoring_args(scoring)
    affine, gap_open, gap_ext = _gap_args(gap_penalty)
    status = ctypes.c_int(0)
    handle = lib.oracle_align_region(
        _ptr(c1, _i64p), len(c1), _ptr(c2, _i64p), len(c2), _ptr(mat, _i32p), n_rows,
        n_cols, match, mismatch, affine, gap_open, gap_ext, int(threshold), int(score_only),
        int(max_number), int(min(max_table_size, 2**62)), ctypes.byref(status))
    traces, score = _collect(lib, handle)
    if status.value:
        raise MemoryError("Maximum table size exceeded")
    return traces, score


def align_local_gapped(code1, code2, scoring, seed, threshold, gap_penalty=-10, max_number=1,
                       direction="both", score_only=False, max_table_size=None,
                       region=None):
    """
    The reference's ``align_local_gapped`` wrapper (localgapped.py:186-311) over `region`
    (default: this module's ``align_region``): split at the seed, reverse the upstream part,
    shift and combine the traces.  Returns ``(traces, score)``.
    """
    import itertools
    region = region or align_region
    if max_table_size is None:
        max_table_size = np.iinfo(np.int64).max
    s1, s2 = int(seed[0]), int(seed[1])
    upstream = direction in ("both", "upstream") and s1 != 0 and s2 != 0
    downstream = direction in ("both", "downstream")
    code1, code2 = np.asarray(code1), np.asarray(code2)
    total, up_traces, down_traces = 0, [], []
    if upstream:
        traces, score = region(code1[s1 - 1::-1], code2[s2 - 1::-1], scoring, gap_penalty,
                               threshold, max_number, score_only, max_table_size)
        total += score
        for trace in traces:
            trace = trace[::-1].copy()
            mask = trace != -1
            trace[mask] *= -1
            trace[mask[:, 0], 0] += s1 - 1
            trace[mask[:, 1], 1] += s2 - 1
            up_traces.append(trace)
    if downstream:
        traces, score = region(code1[s1 + 1:], code2[s2 + 1:], scoring, gap_penalty, threshold,
                               max_number, score_only, max_table_size)
        total += score
        for trace in traces:
            trace = trace.copy()
            trace[trace[:, 0] != -1, 0] += s1 + 1
            trace[trace[:, 1] != -1, 1] += s2 + 1
            down_traces.append(trace)
    if isinstance(scoring, tuple):
        total += scoring[0] if code1[s1] == code2[s2] else scoring[1]
    else:
        total += int(scoring[code1[s1], code2[s2]])
    if score_only:
        return [], total
    seed_row = np.array([[s1, s2]], dtype=np.int64)
    if upstream and downstream:
        combos = itertools.islice(itertools.product(up_traces, down_traces), max_number)
        traces = [np.concatenate([u, seed_row, d]) for u, d in combos]
    elif upstream:
        traces = [np.concatenate([u, seed_row]) for u in up_traces]
    elif downstream:
        traces = [np.concatenate([seed_row, d]) for d in down_traces]
    else:
        traces = [seed_row]
    return traces, total


def prepare_band(len1, len2, band):
    """
    ``banded.py:215-236`` on plain numbers: returns ``(swapped, lower, upper)``
    in the (possibly swapped) frame; raises ``ValueError`` like the reference.
    """
    swapped = len2 < len1
    if swapped:
        len1, len2 = len2, len1
        band = (-band[0], -band[1])
    lower, upper = min(band), max(band)
    if len1 + upper <= 0 or lower >= len2:
        raise ValueError(
            "Alignment band is out of range, the band allows no overlap "
            "between both sequences")
    lower = max(lower, -len1 + 1)
    upper = min(upper, len2 - 1)
    if upper - lower + 1 < 1:
        raise ValueError("The width of the band is 0")
    return swapped, lower, upper


def banded_wrapper(code1, code2, scoring, band, gap_penalty, local, score_only,
                   max_number):
    """
    ``align_banded`` as seen from Python in the reference: takes the sequences
    in caller order, returns traces in caller order.
    """
    swapped, lower, upper = prepare_band(len(code1), len(code2), band)
    if swapped:
        code1, code2 = code2, code1
        if not isinstance(scoring, tuple):
            scoring = np.ascontiguousarray(np.asarray(scoring).T)
    traces, score = align_banded(code1, code2, scoring, gap_penalty, lower, upper,
                                 local, score_only, max_number)
    if swapped:
        traces = [np.ascontiguousarray(np.flip(t, axis=1)) for t in traces]
    return traces, score


def batch(codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True,
          local=False, score_only=False, bands=None, n_threads=None):
    """
    ``max_number=1`` alignment of many pairs on `n_threads` host threads.

    Returns ``(scores int32[n], traces)`` where `traces` is a list of ``(L,2)``
    int64 arrays (``None`` when `score_only`).  `bands` (``(n,2)`` int64,
    already cropped, first sequence the shorter one) selects ``align_banded``.
    """
    lib = _load()
    codes1 = np.ascontiguousarray(codes1, dtype=np.int64)
    codes2 = np.ascontiguousarray(codes2, dtype=np.int64)
    off1 = np.ascontiguousarray(off1, dtype=np.int64)
    off2 = np.ascontiguousarray(off2, dtype=np.int64)
    n_pairs = len(off1) - 1
    mat, n_rows, n_cols, match, mismatch = _scoring_args(scoring)
    affine, gap_open, gap_ext = _gap_args(gap_penalty)
    if bands is not None:
        bands = np.ascontiguousarray(bands, dtype=np.int64)
    scores = np.empty(n_pairs, dtype=np.int32)
    if score_only:
        t_out = t_off = t_len = None
    else:
        cap = np.diff(off1) + np.diff(off2)
        t_off = np.concatenate(([0], np.cumsum(cap))).astype(np.int64)
        t_out = np.empty((int(t_off[-1]), 2), dtype=np.int64)
        t_len = np.zeros(n_pairs, dtype=np.int64)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    lib.oracle_batch(
        _ptr(codes1, _i64p), _ptr(off1, _i64p), _ptr(codes2, _i64p), _ptr(off2, _i64p),
        n_pairs, _ptr(mat, _i32p), n_rows, n_cols, match, mismatch, affine, gap_open,
        gap_ext, int(terminal_penalty), int(local), int(score_only), _ptr(bands, _i64p),
        _ptr(scores, _i32p), _ptr(t_out, _i64p), _ptr(t_off, _i64p), _ptr(t_len, _i64p),
        int(n_threads))
    if score_only:
        return scores, None
    traces = [t_out[t_off[k]:t_off[k] + t_len[k]] for k in range(n_pairs)]
    return scores, traces
