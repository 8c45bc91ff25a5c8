"""The built-in substitution matrices against the reference's data files, and the reference's own
matrix tests (tests/sequence/align/test_matrix.py) restated for the drop-in class."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

import biotite_b200 as bt
from tests.golden.make_matrix_checksums import canonical

CHECKSUMS = json.loads((Path(__file__).parent / "golden" / "matrix_checksums.json").read_text())


def test_collection_is_the_reference_collection():
    assert sorted(bt.SubstitutionMatrix.list_db()) == sorted(CHECKSUMS)


@pytest.mark.parametrize("name", sorted(CHECKSUMS))
def test_matrix_entries_match_the_reference_file(name):
    """Every (symbol, symbol) -> score entry, as SHA-256 of the sorted entry list (fixture written from
    the reference's .mat files by tests/golden/make_matrix_checksums.py)."""
    entries = bt.SubstitutionMatrix.dict_from_db(name)
    assert hashlib.sha256(canonical(entries).encode()).hexdigest() == CHECKSUMS[name]


@pytest.mark.parametrize("name", [n for n in sorted(CHECKSUMS) if n not in ("NUC", "GONNET", "3Di", "PB", "CLESUM")])
def test_protein_matrices_load_in_alphabet_order(name):
    """test_matrix.py:12-27 (every matrix but the five that do not cover the protein alphabet loads over
    it); rows and columns follow the alphabet's order, not the file's (matrix.py:678-684)."""
    alph = bt.ProteinSequence.alphabet
    matrix = bt.SubstitutionMatrix(alph, alph, name)
    entries = bt.SubstitutionMatrix.dict_from_db(name)
    table = matrix.score_matrix()
    assert table.shape == (len(alph), len(alph)) and table.dtype == np.int32
    for i, a in enumerate(alph.get_symbols()):
        for j, b in enumerate(alph.get_symbols()):
            assert table[i, j] == entries[(a, b)]
            assert matrix.get_score(a, b) == entries[(a, b)]
    assert not table.flags.writeable


def test_default_matrices_and_known_entries():
    """test_matrix.py:52-57 and the docstring values of matrix.py."""
    blosum, nuc = bt.SubstitutionMatrix.std_protein_matrix(), bt.SubstitutionMatrix.std_nucleotide_matrix()
    assert blosum.get_score("W", "W") == 11 and blosum.get_score("A", "R") == -1 and blosum.get_score("L", "I") == 2
    assert nuc.get_score("A", "A") == 5 and nuc.get_score("A", "C") == -4
    assert blosum.shape == (24, 24) and nuc.shape == (15, 15)
    assert blosum.is_symmetric() and nuc.is_symmetric() and blosum.transpose() == blosum


def test_matrix_str():
    """test_matrix.py:60-75: the printed table of a small custom matrix."""
    alph1, alph2 = bt.Alphabet("abc"), bt.Alphabet("def")
    matrix = bt.SubstitutionMatrix(alph1, alph2, np.arange(9).reshape(3, 3))
    assert str(matrix) == "\n".join(["    d   e   f", "a   0   1   2", "b   3   4   5", "c   6   7   8"])


@pytest.mark.parametrize("seed", range(5))
def test_as_positional_scores_like_the_original(seed):
    """test_matrix.py:78-109: the positional matrix gives every position pair the score of its symbols."""
    rng = np.random.default_rng(seed)
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    s1, s2 = bt.ProteinSequence(), bt.ProteinSequence()
    s1.code = rng.integers(0, 24, int(rng.integers(1, 40)))
    s2.code = rng.integers(0, 24, int(rng.integers(1, 40)))
    pos, p1, p2 = matrix.as_positional(s1, s2)
    assert pos.shape == (len(s1), len(s2))
    assert np.array_equal(pos.score_matrix(), matrix.score_matrix()[np.ix_(s1.code, s2.code)])
    assert np.array_equal(p1.code, np.arange(len(s1))) and np.array_equal(p2.code, np.arange(len(s2)))
    # the reference's own assertions: same printed sequences, same score symbol by symbol
    assert str(p1) == str(s1) and str(p2) == str(s2)
    for i in range(len(s1)):
        for j in range(len(s2)):
            assert pos.get_score(p1[i], p2[j]) == matrix.get_score(s1[i], s2[j])


def test_constructor_errors():
    alph = bt.ProteinSequence.alphabet
    with pytest.raises(FileNotFoundError):
        bt.SubstitutionMatrix.dict_from_db("NO_SUCH_MATRIX")
    with pytest.raises((ValueError, KeyError)):
        bt.SubstitutionMatrix(alph, alph, {("A", "A"): 1})            # does not cover the alphabet
    with pytest.raises(ValueError):
        bt.SubstitutionMatrix(alph, alph, np.zeros((3, 3), dtype=np.int32))
