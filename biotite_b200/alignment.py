"""
:class:`Alignment` and the helper functions that consume alignment traces.

Counterpart of the reference's ``src/biotite/sequence/align/alignment.py``
(``Alignment`` ``:36-318``, ``get_codes`` ``:321-369``, ``get_symbols``
``:372-416``, ``get_sequence_identity`` ``:419-482``, ``score`` ``:563-644``,
``find_terminal_gaps`` ``:647-705``, ``remove_terminal_gaps`` ``:708-762``,
``remove_gaps`` ``:765-784``).  A *trace* is an ``(L, n_sequences)`` integer
array of sequence positions with ``-1`` standing for a gap.  ``score()`` is an
independent re-scorer and is what pins the DP score in the tests.
"""

from __future__ import annotations

import numbers
import textwrap
from collections.abc import Sequence as SequenceABC

import numpy as np

__all__ = [
    "Alignment",
    "get_codes",
    "get_symbols",
    "get_sequence_identity",
    "get_pairwise_sequence_identity",
    "score",
    "find_terminal_gaps",
    "remove_terminal_gaps",
    "remove_gaps",
]


class Alignment:
    """
    A set of sequences plus the trace that aligns them.

    Parameters
    ----------
    sequences : list of Sequence
    trace : ndarray, shape=(L, n), integer
        Position of each sequence in every alignment column, ``-1`` = gap.
    score : int, optional
    """

    def __init__(self, sequences, trace, score=None):
        self.sequences = list(sequences)
        self.trace = trace
        self.score = score

    @staticmethod
    def from_strings(sequence_strings, sequence_factory, gap_character="-"):
        sequences = [
            sequence_factory(s.replace(gap_character, "")) for s in sequence_strings
        ]
        return Alignment(
            sequences, Alignment.trace_from_strings(sequence_strings, gap_character)
        )

    @staticmethod
    def trace_from_strings(sequence_strings, gap_character="-"):
        if len(sequence_strings) < 2:
            raise ValueError("An alignment must contain at least two sequences")
        length = len(sequence_strings[0])
        if any(len(s) != length for s in sequence_strings):
            raise IndexError("All sequence strings must have the same length")
        trace = np.full((length, len(sequence_strings)), -1, dtype=np.int64)
        for k, s in enumerate(sequence_strings):
            present = np.frombuffer(s.encode("ascii"), dtype=np.uint8) != ord(
                gap_character
            )
            trace[present, k] = np.arange(np.count_nonzero(present))
        return trace

    def _gapped_str(self, seq_index):
        seq = self.sequences[seq_index]
        column = np.asarray(self.trace)[:, seq_index]
        symbols = seq.symbols
        return "".join("-" if p == -1 else str(symbols[p]) for p in column.tolist())

    def get_gapped_sequences(self):
        return [self._gapped_str(i) for i in range(len(self.sequences))]

    def __str__(self):
        if all(_is_single_letter(seq.get_alphabet()) for seq in self.sequences):
            wrapper = textwrap.TextWrapper(break_on_hyphens=False)
            wrapped = [wrapper.wrap(s) for s in self.get_gapped_sequences()]
            blocks = []
            for row in range(len(wrapped[0])):
                blocks.append("\n".join(lines[row] for lines in wrapped))
            return "\n\n".join(blocks)
        return super().__str__()

    def __repr__(self):
        return (
            f"Alignment([{', '.join(repr(s) for s in self.sequences)}], "
            f"np.{np.array_repr(np.asarray(self.trace))}, score={self.score})"
        )

    def __getitem__(self, index):
        if isinstance(index, tuple):
            if len(index) != 2:
                raise IndexError("Only 2D indices are allowed")
            col_index, seq_index = index
            if isinstance(col_index, numbers.Integral) or isinstance(
                seq_index, numbers.Integral
            ):
                raise IndexError(
                    "Integers are invalid indices for alignments, "
                    "a single sequence or alignment column cannot be selected"
                )
            return Alignment(
                _select_sequences(self.sequences, seq_index),
                self.trace[index],
                self.score,
            )
        return Alignment(self.sequences, self.trace[index], self.score)

    def __iter__(self):
        raise NotImplementedError("'Alignment' object is not iterable")

    def __len__(self):
        return len(self.trace)

    def __eq__(self, other):
        return (
            isinstance(other, Alignment)
            and self.sequences == other.sequences
            and np.array_equal(self.trace, other.trace)
            and self.score == other.score
        )

    __hash__ = None


def _select_sequences(sequences, index):
    if isinstance(index, np.ndarray) and index.dtype == bool:
        return [s for s, keep in zip(sequences, index) if keep]
    if isinstance(index, (list, tuple, np.ndarray)):
        return [sequences[i] for i in index]
    if isinstance(index, slice):
        return sequences[index]
    raise IndexError(f"Invalid alignment index type '{type(index).__name__}'")


def _is_single_letter(alphabet):
    if alphabet.is_letter_alphabet():
        return True
    return all(len(str(sym)) == 1 for sym in alphabet)


def get_codes(alignment):
    """
    Symbol codes per alignment column: ``(n_sequences, L)`` int64, ``-1`` = gap.
    """
    trace = np.asarray(alignment.trace)
    codes = np.full((trace.shape[1], trace.shape[0]), -1, dtype=np.int64)
    for k, seq in enumerate(alignment.sequences):
        column = trace[:, k]
        present = column != -1
        codes[k, present] = seq.code[column[present]]
    return codes


def get_symbols(alignment):
    """Like :func:`get_codes`, but symbols; gaps are ``None``."""
    codes = get_codes(alignment)
    out = []
    for k, seq in enumerate(alignment.sequences):
        alph = seq.get_alphabet()
        out.append([None if c == -1 else alph.decode(int(c)) for c in codes[k]])
    return out


def get_sequence_identity(alignment, mode="not_terminal"):
    """
    Fraction of identical columns; `mode` is ``"all"``, ``"not_terminal"`` or
    ``"shortest"`` (denominator as in the reference).
    """
    codes = get_codes(alignment)
    if mode == "all":
        length = len(alignment)
    elif mode == "not_terminal":
        start, stop = find_terminal_gaps(alignment)
        if stop <= start:
            raise ValueError("Cannot calculate non-terminal identity, "
                             "at least two sequences have no overlap")
        codes = codes[:, start:stop]
        length = stop - start
    elif mode == "shortest":
        length = min(len(seq) for seq in alignment.sequences)
    else:
        raise ValueError(f"'{mode}' is an invalid calculation mode")
    matches = np.count_nonzero(
        np.all(codes == codes[0], axis=0) & (codes[0] != -1)
    )
    return matches / length


def get_pairwise_sequence_identity(alignment, mode="not_terminal"):
    """
    :func:`get_sequence_identity` of every pair of rows: an ``(n, n)`` float array
    (reference ``alignment.py:485-560``; the denominator of a pair counts only the two rows).
    """
    codes = get_codes(alignment)
    n = len(codes)
    present = codes != -1
    # columns where both rows carry the same symbol, summed per pair
    matches = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        matches[i] = np.count_nonzero((codes == codes[i]) & present & present[i], axis=1)
    if mode == "all":
        return matches / len(alignment)
    if mode == "shortest":
        lengths = np.array([len(seq) for seq in alignment.sequences])
        return matches / np.minimum.outer(lengths, lengths)
    if mode != "not_terminal":
        raise ValueError(f"'{mode}' is an invalid calculation mode")
    # first and one-past-last column each row occupies; a pair overlaps on [max first, min stop)
    any_present = present.any(axis=1)
    first = np.where(any_present, present.argmax(axis=1), codes.shape[1])
    stop = np.where(any_present, codes.shape[1] - present[:, ::-1].argmax(axis=1), 0)
    length = np.minimum.outer(stop, stop) - np.maximum.outer(first, first)
    if np.any(length <= 0):
        raise ValueError("Cannot calculate non-terminal identity, as the two sequences have no overlap")
    return matches / length


def score(alignment, matrix, gap_penalty=-10, terminal_penalty=True):
    """
    Re-score an alignment from its trace: sum of substitution scores over all
    sequence pairs per column, plus gap penalties per sequence (first gap
    column of a run costs the opening penalty, following ones the extension
    penalty).  With ``terminal_penalty=False`` terminal gap columns are free.
    """
    codes = get_codes(alignment)
    table = matrix.score_matrix()
    total = 0
    n_seq = codes.shape[0]
    for a in range(n_seq):
        for b in range(a + 1, n_seq):
            both = (codes[a] != -1) & (codes[b] != -1)
            total += int(table[codes[a, both], codes[b, both]].sum())

    if isinstance(gap_penalty, SequenceABC):
        gap_open, gap_ext = int(gap_penalty[0]), int(gap_penalty[1])
    elif isinstance(gap_penalty, numbers.Integral):
        gap_open = gap_ext = int(gap_penalty)
    else:
        raise TypeError("Gap penalty must be either integer or tuple")

    if terminal_penalty:
        start, stop = 0, codes.shape[1]
    else:
        start, stop = find_terminal_gaps(alignment)
    for row in codes:
        is_gap = row[start:stop] == -1
        if not is_gap.any():
            continue
        prev = np.concatenate(([False], is_gap[:-1]))
        n_open = int(np.count_nonzero(is_gap & ~prev))
        n_ext = int(np.count_nonzero(is_gap & prev))
        total += n_open * gap_open + n_ext * gap_ext
    return total


def find_terminal_gaps(alignment):
    """
    Column slice ``(start, stop)`` that excludes terminal gaps: columns before
    every sequence has started or after any sequence has ended.
    """
    trace = np.asarray(alignment.trace)
    firsts, lasts = [], []
    for k in range(trace.shape[1]):
        present = np.nonzero(trace[:, k] != -1)[0]
        firsts.append(present[0])
        lasts.append(present[-1])
    return int(np.max(firsts)), int(np.min(lasts)) + 1


def remove_terminal_gaps(alignment):
    start, stop = find_terminal_gaps(alignment)
    if stop < start:
        raise ValueError(
            "Cannot remove terminal gaps, since at least two sequences have "
            "no overlap and the resulting alignment would be empty"
        )
    return alignment[start:stop]


def remove_gaps(alignment):
    keep = (np.asarray(alignment.trace) != -1).all(axis=1)
    return alignment[keep]
