"""
:class:`SubstitutionMatrix` - the scoring table the alignment kernels index by
symbol *code*.

Counterpart of the reference's ``src/biotite/sequence/align/matrix.py``
(constructor ``:156-197``, ``score_matrix`` ``:234-244``, ``transpose``
``:246-259``, ``as_positional`` ``:312-432``, NCBI parser ``:470-502``,
standard matrices ``:539-569``, dict fill ``:678-684``).  The device side
receives ``score_matrix()`` - a C-contiguous read-only int32 array of shape
``(len(alphabet1), len(alphabet2))`` - exactly as the reference hands it to its
native layer.
"""

from __future__ import annotations

import functools
import json
from pathlib import Path

import numpy as np

from .alphabet import Alphabet
from .sequence import NucleotideSequence, PositionalSequence, ProteinSequence

__all__ = ["SubstitutionMatrix"]

_DB_PATH = Path(__file__).resolve().parent / "substitution_matrices.json"


@functools.lru_cache(maxsize=1)
def _database():
    with open(_DB_PATH) as f:
        return json.load(f)


class SubstitutionMatrix:
    """
    Immutable mapping ``(symbol1, symbol2) -> int score`` over two alphabets.

    Parameters
    ----------
    alphabet1, alphabet2 : Alphabet
        Alphabets of the first and second sequence in an alignment.
    score_matrix : ndarray (integer, shape ``(len(alphabet1), len(alphabet2))``), dict or str
        The scores: a code-indexed array, a ``{(sym1, sym2): score}`` dict that
        must cover every symbol pair, or the name of a matrix of the built-in
        collection (``"BLOSUM62"``, ``"NUC"``, ``"PAM250"``, ...).
    """

    def __init__(self, alphabet1, alphabet2, score_matrix):
        self._alph1 = alphabet1
        self._alph2 = alphabet2
        if isinstance(score_matrix, str):
            score_matrix = SubstitutionMatrix.dict_from_db(score_matrix)
        if isinstance(score_matrix, dict):
            matrix = np.zeros((len(alphabet1), len(alphabet2)), dtype=np.int32)
            for i, sym1 in enumerate(alphabet1.get_symbols()):
                for j, sym2 in enumerate(alphabet2.get_symbols()):
                    # A missing pair raises KeyError, as in the reference
                    matrix[i, j] = int(score_matrix[sym1, sym2])
            self._matrix = matrix
        elif isinstance(score_matrix, np.ndarray):
            shape = (len(alphabet1), len(alphabet2))
            if score_matrix.shape != shape:
                raise ValueError(
                    f"Matrix has shape {score_matrix.shape}, but {shape} is required"
                )
            if not np.issubdtype(score_matrix.dtype, np.integer):
                raise TypeError("Score matrix must be an integer ndarray")
            self._matrix = np.ascontiguousarray(score_matrix, dtype=np.int32).copy()
            info = np.iinfo(np.int32)
            if np.any(self._matrix == info.max) or np.any(self._matrix == info.min):
                raise ValueError(
                    "Score values are too large. "
                    "Maybe it was converted from a float matrix containing inf values?"
                )
        else:
            raise TypeError(
                "Matrix must be either a dictionary, an 2-D ndarray or a string"
            )
        self._matrix.setflags(write=False)

    # -- accessors --------------------------------------------------------------
    @property
    def shape(self):
        return (len(self._alph1), len(self._alph2))

    def get_alphabet1(self):
        return self._alph1

    def get_alphabet2(self):
        return self._alph2

    def score_matrix(self) -> np.ndarray:
        return self._matrix

    def get_score_by_code(self, code1, code2):
        return self._matrix[code1, code2]

    def get_score(self, symbol1, symbol2):
        return self._matrix[self._alph1.encode(symbol1), self._alph2.encode(symbol2)]

    def transpose(self):
        return SubstitutionMatrix(
            self._alph2, self._alph1, np.ascontiguousarray(self._matrix.T)
        )

    def is_symmetric(self) -> bool:
        return self._alph1 == self._alph2 and np.array_equal(
            self._matrix, self._matrix.T
        )

    def as_positional(self, sequence1, sequence2):
        """
        Position-specific equivalent: returns ``(pos_matrix, pos_seq1, pos_seq2)``
        with ``pos_matrix[i, j] == self[sequence1[i], sequence2[j]]``.
        """
        pos_seq1 = PositionalSequence(sequence1)
        pos_seq2 = PositionalSequence(sequence2)
        pos_scores = self._matrix[
            tuple(np.meshgrid(sequence1.code, sequence2.code, indexing="ij"))
        ]
        pos_matrix = SubstitutionMatrix(
            pos_seq1.get_alphabet(), pos_seq2.get_alphabet(), pos_scores
        )
        return pos_matrix, pos_seq1, pos_seq2

    # -- comparison / printing ----------------------------------------------------
    def __eq__(self, other):
        return (
            isinstance(other, SubstitutionMatrix)
            and self._alph1 == other._alph1
            and self._alph2 == other._alph2
            and np.array_equal(self._matrix, other._matrix)
        )

    def __ne__(self, other):
        return not self == other

    __hash__ = None

    def __repr__(self):
        return (
            f"SubstitutionMatrix({self._alph1!r}, {self._alph2!r}, "
            f"np.{np.array_repr(self._matrix)})"
        )

    def __str__(self):
        lines = [" " + "".join(f" {str(s):>3}" for s in self._alph2)]
        for i, sym in enumerate(self._alph1):
            lines.append(
                f"{str(sym):>1}"
                + "".join(f" {int(v):>3d}" for v in self._matrix[i])
            )
        return "\n".join(lines)

    # -- built-in collection ------------------------------------------------------
    @staticmethod
    def dict_from_str(string):
        """
        Parse a matrix in NCBI text format into ``{(sym1, sym2): score}``.

        As in the reference (``matrix.py:491-501``) the entry for
        ``(row_symbol, column_symbol)`` is taken from the *transposed* table;
        all shipped matrices are symmetric, so this is unobservable for them.
        """
        lines = [ln.strip() for ln in string.split("\n")]
        lines = [ln for ln in lines if ln and ln[0] != "#"]
        symbols2 = lines[0].split()
        symbols1 = [ln.split()[0] for ln in lines[1:]]
        table = np.array([ln.split()[1:] for ln in lines[1:]]).astype(int).T
        return {
            (s1, s2): int(table[i, j])
            for i, s1 in enumerate(symbols1)
            for j, s2 in enumerate(symbols2)
        }

    @staticmethod
    def dict_from_db(matrix_name):
        try:
            entry = _database()[matrix_name]
        except KeyError:
            raise FileNotFoundError(
                f"Matrix '{matrix_name}' is not in the built-in collection"
            ) from None
        table = np.array(entry["scores"], dtype=int).T
        return {
            (s1, s2): int(table[i, j])
            for i, s1 in enumerate(entry["rows"])
            for j, s2 in enumerate(entry["cols"])
        }

    @staticmethod
    def list_db():
        return list(_database().keys())

    @staticmethod
    @functools.lru_cache(maxsize=None)
    def std_protein_matrix():
        """BLOSUM62 over the 24-symbol protein alphabet."""
        return SubstitutionMatrix(
            ProteinSequence.alphabet, ProteinSequence.alphabet, "BLOSUM62"
        )

    @staticmethod
    @functools.lru_cache(maxsize=None)
    def std_nucleotide_matrix():
        """NUC (match 5, mismatch -4) over the 15-symbol ambiguous DNA alphabet."""
        return SubstitutionMatrix(
            NucleotideSequence.alphabet_amb, NucleotideSequence.alphabet_amb, "NUC"
        )
