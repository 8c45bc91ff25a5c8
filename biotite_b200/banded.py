"""
``align_banded`` - drop-in for ``biotite.sequence.align.align_banded``
(reference wrapper: ``src/biotite/sequence/align/banded.py:46-261``; native
entry ``src/rust/sequence/align/banded.rs:63-77``).
"""

from __future__ import annotations

import numpy as np

from .alignment import Alignment
from .matrix import SubstitutionMatrix
from .pairwise import _check_gap_penalty, _native_align_pair, prepare_scoring

__all__ = ["align_banded"]


def crop_band(len1, len2, band):
    """
    The reference's band preparation (``banded.py:225-236``) for sequences that
    are already ordered shorter-first: returns the cropped ``(lower, upper)``
    or raises ``ValueError``.
    """
    lower, upper = min(band), max(band)
    if len1 + upper <= 0 or lower >= len2:
        raise ValueError(
            "Alignment band is out of range, the band allows no overlap "
            "between both sequences"
        )
    lower = max(lower, -len1 + 1)
    upper = min(upper, len2 - 1)
    if upper - lower + 1 < 1:
        raise ValueError("The width of the band is 0")
    return lower, upper


def align_banded(seq1, seq2, matrix, band, gap_penalty=-10, local=False,
                 max_number=1000, score_only=False):
    """
    Semi-global or local alignment restricted to the diagonal band
    ``band[0] <= j - i <= band[1]`` of the DP table.

    Parameters
    ----------
    seq1, seq2 : Sequence
    matrix : SubstitutionMatrix or tuple(int, int)
    band : tuple(int, int)
        Lower and upper diagonal; a diagonal ``d`` relates positions ``i`` of
        `seq1` and ``j = i + d`` of `seq2`.
    gap_penalty : int or tuple(int, int), optional
    local : bool, optional
    max_number : int, optional
    score_only : bool, optional

    Returns
    -------
    alignments : list of Alignment, or int if `score_only`
    """
    _check_gap_penalty(gap_penalty)
    if max_number < 1:
        raise ValueError("Maximum number of returned alignments must be at least 1")

    # The shorter sequence runs along the rows of the band table
    swapped = len(seq2) < len(seq1)
    if swapped:
        seq1, seq2 = seq2, seq1
        band = (-band[0], -band[1])
        if isinstance(matrix, SubstitutionMatrix):
            matrix = matrix.transpose()
    lower, upper = crop_band(len(seq1), len(seq2), band)

    common_dtype = np.promote_types(seq1.code.dtype, seq2.code.dtype)
    traces, score = _native_align_pair(
        np.ascontiguousarray(seq1.code, common_dtype),
        np.ascontiguousarray(seq2.code, common_dtype),
        prepare_scoring(matrix, seq1, seq2),
        gap_penalty, True, local, score_only, max_number, band=(lower, upper),
    )
    if score_only:
        return score
    if swapped:
        return [Alignment([seq2, seq1], np.flip(t, axis=1), score) for t in traces]
    return [Alignment([seq1, seq2], t, score) for t in traces]
