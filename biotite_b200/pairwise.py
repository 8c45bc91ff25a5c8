"""
``align_optimal`` - drop-in for ``biotite.sequence.align.align_optimal``
(reference wrapper: ``src/biotite/sequence/align/pairwise.py:117-291``).

The Python layer does what the reference's does - argument checks, promotion
of both code arrays to one unsigned dtype, resolution of the scoring scheme -
and then calls ``bt_align_pair`` of ``libb200align.so`` where the reference
calls its native ``align_optimal`` (``src/rust/sequence/align/pairwise.rs:57-70``).
The DP fill and the traceback run on the GPU; co-optimal enumeration for
``max_number > 1`` walks the direction table the GPU produced.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from .alignment import Alignment
from .matrix import SubstitutionMatrix

__all__ = ["align_optimal", "align_ungapped", "prepare_scoring"]


def align_ungapped(seq1, seq2, matrix, score_only=False):
    """
    Score two equally long sequences position by position; the alignment has the trivial trace
    ``[[0, 0], [1, 1], ...]`` (reference ``pairwise.py:36-92``: no table, one gather and a sum on
    the host there as well).
    """
    if len(seq1) != len(seq2):
        raise ValueError(f"Different sequence lengths ({len(seq1):d} and {len(seq2):d})")
    scoring = prepare_scoring(matrix, seq1, seq2)
    if isinstance(scoring, tuple):
        n_equal = int(np.count_nonzero(seq1.code == seq2.code))
        score = n_equal * scoring[0] + (len(seq1) - n_equal) * scoring[1]
    else:
        score = int(scoring[seq1.code, seq2.code].sum(dtype=np.int64))
    if score_only:
        return score
    positions = np.arange(len(seq1), dtype=np.int64)
    return Alignment([seq1, seq2], np.stack((positions, positions), axis=1), score)


def _check_gap_penalty(gap_penalty):
    # bool is an int in Python, as in the reference
    if isinstance(gap_penalty, int):
        if gap_penalty > 0:
            raise ValueError("Gap penalty must be negative")
    elif isinstance(gap_penalty, tuple):
        if gap_penalty[0] > 0 or gap_penalty[1] > 0:
            raise ValueError("Gap penalty must be negative")
    else:
        raise TypeError("Gap penalty must be either integer or tuple")


def prepare_scoring(scoring, seq1, seq2, check_matrix=True):
    """
    ``(match, mismatch)`` tuples become tuples of plain ints; a
    :class:`SubstitutionMatrix` is checked against the sequences' alphabets
    (prefix ``extends`` relation) and resolved to its int32 score array.
    """
    if isinstance(scoring, tuple):
        match_score, mismatch_score = scoring
        return (int(match_score), int(mismatch_score))
    if check_matrix:
        if seq1 is None or seq2 is None:
            raise TypeError("Sequences are required to check the matrix")
        if not scoring.get_alphabet1().extends(
            seq1.get_alphabet()
        ) or not scoring.get_alphabet2().extends(seq2.get_alphabet()):
            raise ValueError("The sequences' alphabets do not fit the matrix")
    return scoring.score_matrix()


def _native_align_pair(code1, code2, scoring, gap_penalty, terminal_penalty, local,
                       score_only, max_number, band=None):
    """
    Same contract as the reference's native functions: returns
    ``(list of (L,2) int64 traces, score)``; the list is empty if `score_only`.
    """
    lib = _cabi.load()
    P, matrix = _cabi.make_params(
        scoring, gap_penalty, terminal_penalty, local, band is not None, code1.dtype
    )
    lower, upper = (0, 0) if band is None else (int(band[0]), int(band[1]))
    score = ctypes.c_int32(0)
    handle = ctypes.c_void_p(None)
    status = lib.bt_align_pair(
        ctypes.byref(P), _cabi.ptr(code1), len(code1), _cabi.ptr(code2), len(code2),
        _cabi.ptr(matrix), lower, upper, int(bool(score_only)), int(max_number),
        ctypes.byref(score), ctypes.byref(handle),
    )
    _cabi.check(status)
    traces = []
    if not score_only:
        try:
            for k in range(lib.bt_tracelist_count(handle)):
                length = lib.bt_tracelist_length(handle, k)
                trace = np.empty((length, 2), dtype=np.int64)
                _cabi.check(lib.bt_tracelist_copy(handle, k, _cabi.ptr(trace)))
                traces.append(trace)
        finally:
            lib.bt_tracelist_free(handle)
    return traces, int(score.value)


def align_optimal(seq1, seq2, matrix, gap_penalty=-10, terminal_penalty=True,
                  local=False, max_number=1000, score_only=False):
    """
    Optimal global (Needleman-Wunsch), semi-global or local (Smith-Waterman)
    alignment of two sequences with a linear or affine (Gotoh) gap penalty.

    Parameters
    ----------
    seq1, seq2 : Sequence
        The sequences to be aligned.
    matrix : SubstitutionMatrix or tuple(int, int)
        The substitution matrix, or a ``(match, mismatch)`` score pair.
    gap_penalty : int or tuple(int, int), optional
        Linear penalty, or ``(opening, extension)``; values must not be positive.
    terminal_penalty : bool, optional
        If false, terminal gaps are free (semi-global).  Ignored if `local`.
    local : bool, optional
        Local instead of global alignment.
    max_number : int, optional
        Maximum number of co-optimal alignments returned.
    score_only : bool, optional
        Return only the optimal score.

    Returns
    -------
    alignments : list of Alignment, or int if `score_only`
    """
    _check_gap_penalty(gap_penalty)
    if max_number < 1:
        raise ValueError("Maximum number of returned alignments must be at least 1")

    common_dtype = np.promote_types(seq1.code.dtype, seq2.code.dtype)
    traces, score = _native_align_pair(
        np.ascontiguousarray(seq1.code, common_dtype),
        np.ascontiguousarray(seq2.code, common_dtype),
        prepare_scoring(matrix, seq1, seq2),
        gap_penalty, terminal_penalty, local, score_only, max_number,
    )
    if score_only:
        return score
    return [Alignment([seq1, seq2], trace, score) for trace in traces]
