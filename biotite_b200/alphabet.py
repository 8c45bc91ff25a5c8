"""
Symbol alphabets: the mapping between symbols and the integer *codes* the
alignment kernels consume.

Mirrors the behaviour of the reference's ``biotite.sequence.alphabet``
(``src/biotite/sequence/alphabet.py:28-157`` for :class:`Alphabet`,
``:297-523`` for :class:`LetterAlphabet`): the prefix ``extends`` relation,
``encode``/``decode`` and their ``*_multiple`` array forms, and
:class:`AlphabetError` for unknown symbols.  Only what the alignment hot path
and its tests need is provided.
"""

from __future__ import annotations

import numpy as np

__all__ = ["Alphabet", "LetterAlphabet", "AlphabetError", "code_dtype"]


class AlphabetError(Exception):
    """Raised when a symbol or a code is not part of an alphabet."""


def code_dtype(alphabet_size: int):
    """
    Smallest unsigned dtype able to hold every code of an alphabet of the
    given size (reference: ``sequence.py:359-384``).
    """
    for dtype in (np.uint8, np.uint16, np.uint32):
        if alphabet_size <= np.iinfo(dtype).max + 1:
            return dtype
    return np.uint64


class Alphabet:
    """
    An ordered collection of hashable symbols; the position of a symbol is its
    code.

    Parameters
    ----------
    symbols : iterable
        The symbols in code order.
    """

    def __init__(self, symbols):
        self._symbols = tuple(symbols)
        if not self._symbols:
            raise ValueError("Symbol list is empty")
        self._codes = {sym: i for i, sym in enumerate(self._symbols)}

    def get_symbols(self):
        return self._symbols

    def extends(self, alphabet) -> bool:
        """
        ``True`` if `alphabet` is a prefix of this alphabet, i.e. every code of
        `alphabet` means the same symbol here.
        """
        if alphabet is self:
            return True
        n = len(alphabet)
        if n > len(self):
            return False
        return tuple(alphabet.get_symbols()) == tuple(self.get_symbols()[:n])

    def encode(self, symbol) -> int:
        try:
            return self._codes[symbol]
        except KeyError:
            raise AlphabetError(f"Symbol {symbol!r} is not in the alphabet") from None

    def decode(self, code):
        if code < 0 or code >= len(self._symbols):
            raise AlphabetError(f"'{code}' is not a valid code")
        return self._symbols[code]

    def encode_multiple(self, symbols, dtype=np.int64):
        return np.array([self.encode(s) for s in symbols], dtype=dtype)

    def decode_multiple(self, code):
        return [self.decode(int(c)) for c in code]

    def is_letter_alphabet(self) -> bool:
        return all(isinstance(s, (str, bytes)) and len(s) == 1 for s in self._symbols)

    def __len__(self):
        return len(self._symbols)

    def __iter__(self):
        return iter(self.get_symbols())

    def __contains__(self, symbol):
        return symbol in self._codes

    def __hash__(self):
        return hash(self._symbols)

    def __eq__(self, other):
        if other is self:
            return True
        if not isinstance(other, Alphabet):
            return False
        return tuple(self.get_symbols()) == tuple(other.get_symbols())

    def __str__(self):
        return str(list(self.get_symbols()))

    def __repr__(self):
        return f"Alphabet({list(self.get_symbols())!r})"


class LetterAlphabet(Alphabet):
    """
    An alphabet of single printable ASCII characters.  Encoding and decoding of
    whole strings go through a 256-entry byte table, so they are vectorised.
    """

    _INVALID = 255

    def __init__(self, symbols):
        if isinstance(symbols, bytes):
            symbols = symbols.decode("ascii")
        letters = []
        for sym in symbols:
            if isinstance(sym, bytes):
                sym = sym.decode("ascii")
            if not isinstance(sym, str) or len(sym) != 1 or not (32 <= ord(sym) < 127):
                raise ValueError(f"Symbol {sym!r} is not a single printable letter")
            letters.append(sym)
        if len(letters) > self._INVALID:
            raise ValueError("Too many symbols for a letter alphabet")
        super().__init__(letters)
        self._ascii = np.frombuffer("".join(letters).encode("ascii"), dtype=np.uint8)
        self._table = np.full(256, self._INVALID, dtype=np.uint8)
        self._table[self._ascii] = np.arange(len(letters), dtype=np.uint8)

    def encode(self, symbol) -> int:
        if isinstance(symbol, bytes):
            symbol = symbol.decode("ascii")
        return super().encode(symbol)

    def encode_multiple(self, symbols, dtype=None):
        if isinstance(symbols, str):
            raw = np.frombuffer(symbols.encode("ascii", errors="replace"), dtype=np.uint8)
        elif isinstance(symbols, bytes):
            raw = np.frombuffer(symbols, dtype=np.uint8)
        elif isinstance(symbols, np.ndarray) and symbols.dtype == np.uint8:
            raw = symbols
        else:
            joined = "".join(
                s.decode("ascii") if isinstance(s, bytes) else str(s) for s in symbols
            )
            raw = np.frombuffer(joined.encode("ascii", errors="replace"), dtype=np.uint8)
        code = self._table[raw]
        bad = np.nonzero(code == self._INVALID)[0]
        if bad.size:
            raise AlphabetError(
                f"Symbol {chr(int(raw[bad[0]]))!r} is not in the alphabet"
            )
        if dtype is not None:
            code = code.astype(dtype, copy=False)
        return code

    def decode_multiple(self, code, as_bytes=False):
        code = np.asarray(code)
        if code.size and (code.min() < 0 or code.max() >= len(self)):
            raise AlphabetError("Sequence code contains invalid values")
        raw = self._ascii[code.astype(np.int64, copy=False)]
        if as_bytes:
            return raw.view("S1")
        return np.array(list(raw.tobytes().decode("ascii")), dtype="U1")

    def is_letter_alphabet(self) -> bool:
        return True

    def __repr__(self):
        return f"LetterAlphabet({''.join(self.get_symbols())!r})"
