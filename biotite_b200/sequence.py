"""
Sequence containers whose ``.code`` arrays feed the alignment kernels.

Drop-in counterparts of the reference's ``Sequence`` base class
(``src/biotite/sequence/sequence.py:30-384``) and of the concrete types in
``src/biotite/sequence/seqtypes.py`` (``GeneralSequence`` ``:36-104``,
``NucleotideSequence`` ``:107-397``, ``ProteinSequence`` ``:399-643``,
``PositionalSequence`` ``:646-732``, ``PurePositionalSequence`` ``:735-758``),
reduced to what the pairwise alignment path touches: symbol <-> code
conversion, the code dtype rule, slicing, equality and copying.
"""

from __future__ import annotations

import copy as _copy

import numpy as np

from .alphabet import Alphabet, AlphabetError, LetterAlphabet, code_dtype

__all__ = [
    "Sequence",
    "GeneralSequence",
    "NucleotideSequence",
    "ProteinSequence",
    "PositionalSequence",
    "PurePositionalSequence",
]


class Sequence:
    """
    Base class: an alphabet plus an unsigned integer code array.

    Subclasses provide :meth:`get_alphabet`.  Assigning ``seq.code`` casts to
    the alphabet's dtype *without validation*, as the reference does
    (``sequence.py:198-203``); the tests build random sequences this way.
    """

    def __init__(self, sequence=()):
        self.symbols = sequence

    @staticmethod
    def dtype(alphabet_size: int):
        return code_dtype(alphabet_size)

    def get_alphabet(self) -> Alphabet:
        raise NotImplementedError

    @property
    def alphabet(self) -> Alphabet:
        return self.get_alphabet()

    @property
    def code(self) -> np.ndarray:
        return self._seq_code

    @code.setter
    def code(self, value):
        dtype = code_dtype(len(self.get_alphabet()))
        self._seq_code = np.asarray(value).astype(dtype, copy=False)

    @property
    def symbols(self):
        return self.get_alphabet().decode_multiple(self.code)

    @symbols.setter
    def symbols(self, value):
        alph = self.get_alphabet()
        dtype = code_dtype(len(alph))
        if len(value) == 0:
            self._seq_code = np.zeros(0, dtype=dtype)
        else:
            self._seq_code = np.asarray(alph.encode_multiple(value)).astype(dtype)

    def __copy_create__(self):
        return type(self)()

    def copy(self, new_seq_code=None):
        clone = self.__copy_create__()
        for key, val in self.__dict__.items():
            if key != "_seq_code" and key not in clone.__dict__:
                clone.__dict__[key] = val
        clone.code = self.code.copy() if new_seq_code is None else new_seq_code
        return clone

    def reverse(self, copy=True):
        rev = np.flip(self.code, axis=0)
        if copy:
            rev = np.copy(rev)
        return self.copy(rev)

    def is_valid(self) -> bool:
        return bool((self.code < len(self.get_alphabet())).all())

    def get_symbol_frequency(self):
        counts = np.bincount(self.code, minlength=len(self.get_alphabet()))
        return {sym: int(counts[i]) for i, sym in enumerate(self.get_alphabet())}

    def __getitem__(self, index):
        alph = self.get_alphabet()
        sub = self.code[index]
        if isinstance(sub, np.ndarray):
            return self.copy(sub.copy())
        return alph.decode(int(sub))

    def __setitem__(self, index, item):
        alph = self.get_alphabet()
        if isinstance(index, (int, np.integer)):
            self._seq_code[index] = alph.encode(item)
        else:
            if isinstance(item, Sequence):
                self._seq_code[index] = item.code
            elif isinstance(item, np.ndarray):
                self._seq_code[index] = item          # an array is taken as codes (sequence.py:312-313)
            else:
                self._seq_code[index] = alph.encode_multiple(item)

    def __len__(self):
        return len(self._seq_code)

    def __iter__(self):
        alph = self.get_alphabet()
        for c in self.code:
            yield alph.decode(int(c))

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return False
        if self.get_alphabet() != other.get_alphabet():
            return False
        return np.array_equal(self.code, other.code)

    __hash__ = None

    def __str__(self):
        alph = self.get_alphabet()
        if isinstance(alph, LetterAlphabet):
            return alph.decode_multiple(self.code, as_bytes=True).tobytes().decode("ascii")
        return ", ".join(str(s) for s in self.symbols)

    def __add__(self, other):
        if self.get_alphabet().extends(other.get_alphabet()):
            return self.copy(np.concatenate((self.code, other.code)))
        if other.get_alphabet().extends(self.get_alphabet()):
            return other.copy(np.concatenate((self.code, other.code)))
        raise ValueError("The sequences alphabets are not compatible")


class GeneralSequence(Sequence):
    """A sequence over an arbitrary user supplied :class:`Alphabet`."""

    def __init__(self, alphabet, sequence=()):
        self._alphabet = alphabet
        super().__init__(sequence)

    def __copy_create__(self):
        return GeneralSequence(self._alphabet)

    def get_alphabet(self):
        return self._alphabet

    def __repr__(self):
        return f"GeneralSequence({self._alphabet!r}, {list(self.symbols)!r})"


class NucleotideSequence(Sequence):
    """
    DNA sequence.  Uses the 4-letter alphabet ``ACGT`` when possible and falls
    back to the 15-letter IUPAC alphabet (which *extends* the former, so the
    15x15 NUC matrix scores both) unless `ambiguous` forces the choice.
    """

    alphabet_unamb = LetterAlphabet("ACGT")
    alphabet_amb = LetterAlphabet("ACGTRYWSMKHBVDN")

    _complement = dict(zip("ACGTMRWSYKVHDBN", "TGCAKYWSRMBDHVN"))

    def __init__(self, sequence=(), ambiguous=None):
        if isinstance(sequence, str):
            sequence = sequence.upper()
        else:
            sequence = [s.upper() for s in sequence]
        if ambiguous is None:
            try:
                self._alphabet = NucleotideSequence.alphabet_unamb
                code = self._alphabet.encode_multiple(sequence)
            except AlphabetError:
                self._alphabet = NucleotideSequence.alphabet_amb
                code = self._alphabet.encode_multiple(sequence)
        else:
            self._alphabet = (
                NucleotideSequence.alphabet_amb
                if ambiguous
                else NucleotideSequence.alphabet_unamb
            )
            code = self._alphabet.encode_multiple(sequence)
        self.code = code

    def __copy_create__(self):
        return NucleotideSequence(
            ambiguous=self._alphabet is NucleotideSequence.alphabet_amb
            or self._alphabet == NucleotideSequence.alphabet_amb
        )

    def get_alphabet(self):
        return self._alphabet

    def complement(self):
        table = np.array(
            [
                NucleotideSequence.alphabet_amb.encode(self._complement[s])
                for s in NucleotideSequence.alphabet_amb
            ],
            dtype=np.uint8,
        )
        return self.copy(table[self.code])

    @staticmethod
    def unambiguous_alphabet():
        return NucleotideSequence.alphabet_unamb

    @staticmethod
    def ambiguous_alphabet():
        return NucleotideSequence.alphabet_amb

    def __repr__(self):
        amb = self._alphabet == NucleotideSequence.alphabet_amb
        return f'NucleotideSequence("{self}", ambiguous={amb})'


class ProteinSequence(Sequence):
    """
    Protein sequence over the 24-symbol alphabet ``ACDEFGHIKLMNPQRSTVWYBZX*``.
    Three-letter residue names are accepted and converted.
    """

    alphabet = LetterAlphabet("ACDEFGHIKLMNPQRSTVWYBZX*")

    _dict_1to3 = {
        "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY",
        "H": "HIS", "I": "ILE", "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN",
        "P": "PRO", "Q": "GLN", "R": "ARG", "S": "SER", "T": "THR", "V": "VAL",
        "W": "TRP", "Y": "TYR", "B": "ASX", "Z": "GLX", "X": "UNK", "*": "TER",
    }  # fmt: skip
    _dict_3to1 = {v: k for k, v in _dict_1to3.items()}
    _dict_3to1["SEC"] = "C"
    _dict_3to1["MSE"] = "M"

    def __init__(self, sequence=()):
        if isinstance(sequence, str):
            sequence = sequence.upper()
        else:
            table = ProteinSequence._dict_3to1
            sequence = [
                table[s.upper()] if len(s) == 3 else s.upper() for s in sequence
            ]
        super().__init__(sequence)

    def get_alphabet(self):
        return ProteinSequence.alphabet

    def remove_stops(self):
        stop = ProteinSequence.alphabet.encode("*")
        return self.copy(self.code[self.code != stop])

    @staticmethod
    def convert_letter_3to1(symbol):
        return ProteinSequence._dict_3to1[symbol.upper()]

    @staticmethod
    def convert_letter_1to3(symbol):
        return ProteinSequence._dict_1to3[symbol.upper()]

    def __repr__(self):
        return f'ProteinSequence("{self}")'


class PositionalSequence(Sequence):
    """
    A sequence whose code *is* the position: symbol ``k`` of the private
    alphabet stands for "position k of `original_sequence`".  Used together
    with position-specific substitution matrices
    (:meth:`SubstitutionMatrix.as_positional`).
    """

    class Symbol:
        __slots__ = ("original_alphabet", "original_code", "position", "symbol")

        def __init__(self, original_alphabet, original_code, position):
            self.original_alphabet = original_alphabet
            self.original_code = original_code
            self.position = position
            self.symbol = original_alphabet.decode(original_code)

        def _key(self):
            return (id(self.original_alphabet), self.original_code, self.position)

        def __eq__(self, other):
            return (
                isinstance(other, PositionalSequence.Symbol)
                and self.original_alphabet == other.original_alphabet
                and self.original_code == other.original_code
                and self.position == other.position
            )

        def __hash__(self):
            return hash((self.original_code, self.position))

        def __str__(self):
            return str(self.symbol)

        def __repr__(self):
            return f"Symbol({self.symbol!r}@{self.position})"

    def __init__(self, original_sequence):
        self._orig_alphabet = original_sequence.get_alphabet()
        self._alphabet = Alphabet(
            [
                PositionalSequence.Symbol(self._orig_alphabet, int(c), pos)
                for pos, c in enumerate(original_sequence.code.tolist())
            ]
        )
        self.code = np.arange(len(original_sequence))

    def __copy_create__(self):
        return _copy.copy(self)

    def reconstruct(self):
        seq = GeneralSequence(self._orig_alphabet)
        seq.code = np.array([s.original_code for s in self._alphabet])
        return seq

    def get_alphabet(self):
        return self._alphabet

    def __str__(self):
        return "".join(str(s) for s in self.symbols)

    def __repr__(self):
        return f"PositionalSequence({self.reconstruct()!r})"


class PurePositionalSequence(Sequence):
    """Placeholder sequence ``0, 1, ..., length-1`` over the alphabet of positions."""

    def __init__(self, length):
        self._alphabet = Alphabet(range(length))
        self.code = np.arange(length)

    def __copy_create__(self):
        return _copy.copy(self)

    def get_alphabet(self):
        return self._alphabet

    def __repr__(self):
        return f"PurePositionalSequence({len(self)})"
