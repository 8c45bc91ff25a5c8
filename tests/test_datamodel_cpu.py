"""Alphabet / Sequence behaviour the alignment path relies on, restated from the reference's tests
(tests/sequence/test_alphabet.py, test_sequence.py, test_seqtypes.py) for the drop-in classes."""
import numpy as np
import pytest

import biotite_b200 as seq

CASES = {"A": [0], "D": [3], "ABC": [0, 1, 2], "ABAFF": [0, 1, 0, 5, 5]}


@pytest.mark.parametrize("cls", [seq.Alphabet, seq.LetterAlphabet])
@pytest.mark.parametrize("symbols", list(CASES))
def test_encoding_and_decoding(cls, symbols):
    """test_alphabet.py:23-65."""
    alph, code = cls("ABCDEF"), CASES[symbols]
    if len(symbols) == 1:
        assert alph.encode(symbols) == code[0] and alph.decode(np.uint8(code[0])) == symbols
    else:
        assert list(alph.encode_multiple(symbols)) == code
        assert list(alph.decode_multiple(np.array(code, dtype=np.uint8))) == list(symbols)
    assert len(alph) == 6 and "D" in alph and "G" not in alph


@pytest.mark.parametrize("cls", [seq.Alphabet, seq.LetterAlphabet])
def test_alphabet_errors(cls):
    """test_alphabet.py:68-92: symbols and codes outside the alphabet raise AlphabetError."""
    alph = cls("ABCDEF")
    for call, arg in ((alph.encode, "G"), (alph.encode, 42), (alph.decode, 6), (alph.decode, -1),
                      (alph.encode_multiple, "G"), (alph.encode_multiple, [42]),
                      (alph.decode_multiple, np.array([6])), (alph.decode_multiple, np.array([-1]))):
        with pytest.raises(seq.AlphabetError):
            call(arg)


@pytest.mark.parametrize("symbols", ["ABC", b"ABC", ["A", "B", "C"], np.array(["A", "B", "C"]),
                                     np.array([b"A", b"B", b"C"])])
def test_letter_alphabet_input_types(symbols):
    """test_alphabet.py:95-123."""
    alph = seq.LetterAlphabet("ABCDEF")
    assert list(alph.decode_multiple(alph.encode_multiple(symbols))) == ["A", "B", "C"]


def test_alphabet_extension_is_a_prefix_relation():
    """test_sequence.py:74-84; what `prepare_scoring` checks a matrix against (pairwise.py:283-290)."""
    a1, a2, a3, a4 = seq.Alphabet("abc"), seq.Alphabet("abc"), seq.Alphabet("acb"), seq.Alphabet("abcde")
    assert a1.extends(a1) and a2.extends(a1) and not a3.extends(a1) and a4.extends(a1) and not a1.extends(a4)
    # a 4-letter DNA sequence fits the 15 x 15 NUC matrix because the ambiguous alphabet starts with ACGT
    assert seq.NucleotideSequence.alphabet_amb.extends(seq.NucleotideSequence.alphabet_unamb)


def test_sequence_round_trip_access_and_validity():
    """test_sequence.py:10-31."""
    text = "AATGCGTTA"
    dna = seq.NucleotideSequence(text)
    assert str(dna) == text and dna[2] == text[2] and "".join(dna) == text and str(dna[3:-2]) == "GCGT"
    raw = seq.NucleotideSequence()
    raw.code = np.array([0, 1, 0, 3, 3])
    assert raw.is_valid()
    raw.code = np.array([0, 1, 4, 3, 3])          # no validation on assignment (sequence.py:198-203)
    assert not raw.is_valid()
    with pytest.raises(seq.AlphabetError):
        seq.NucleotideSequence("AATGCGTUTA")


def test_sequence_manipulation_and_concatenation():
    """test_sequence.py:34-59."""
    dna = seq.NucleotideSequence("ACGTA")
    copy = dna.copy(); copy[2] = "C"
    assert str(copy) == "ACCTA" and str(dna) == "ACGTA"
    copy = dna.copy(); copy[0:2] = copy[3:5]
    assert str(copy) == "TAGTA"
    copy = dna.copy(); copy[np.array([True, False, False, False, True])] = "T"
    assert str(copy) == "TCGTT"
    copy = dna.copy(); copy[1:4] = np.array([0, 1, 2])
    assert str(copy) == "AACGA"
    for a, b in (("AAGTTA", "CGA"), ("AAGTTA", "NNN"), ("NNN", "AAGTTA")):
        assert str(seq.NucleotideSequence(a) + seq.NucleotideSequence(b)) == a + b
    assert seq.NucleotideSequence("ACGCGAGAAAGCGGG").get_symbol_frequency() == {"A": 5, "C": 3, "G": 7, "T": 0}


def test_sequence_types():
    """test_seqtypes.py:10-18, :74-83, :96-104 and the code dtype rule the kernels rely on."""
    assert str(seq.NucleotideSequence("acgt")) == "ACGT"
    assert seq.NucleotideSequence("ACGT").get_alphabet() == seq.NucleotideSequence.alphabet_unamb
    assert seq.NucleotideSequence("ACGTN").get_alphabet() == seq.NucleotideSequence.alphabet_amb
    assert seq.NucleotideSequence("ACGT", ambiguous=True).get_alphabet() == seq.NucleotideSequence.alphabet_amb
    assert str(seq.ProteinSequence(["Ala", "Trp", "Cys"])) == "AWC" and str(seq.ProteinSequence("acw")) == "ACW"
    assert seq.ProteinSequence("ACW").code.dtype == np.uint8
    positional = seq.PositionalSequence(seq.ProteinSequence("ACW"))
    assert str(positional) == "ACW" and list(positional.code) == [0, 1, 2]
