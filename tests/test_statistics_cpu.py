"""EValueEstimator: the argument checks that run before any alignment (reference
tests/sequence/align/test_statistics.py:180-195, statistics.py:162-195) and the E-value formula
(statistics.py:219-295).  The sampling itself is a GPU batch (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

import biotite_b200 as bt


def test_invalid_scoring_scheme():
    alph = bt.ProteinSequence.alphabet
    matrix = bt.SubstitutionMatrix(alph, alph, np.ones((len(alph), len(alph)), dtype=int))
    with pytest.raises(ValueError):          # positive expected score of two random symbols
        bt.EValueEstimator.from_samples(alph, matrix, -10, np.ones(len(alph)))


def test_argument_checks():
    alph = bt.ProteinSequence.alphabet
    blosum = bt.SubstitutionMatrix.std_protein_matrix()
    with pytest.raises(IndexError):          # one frequency per symbol
        bt.EValueEstimator.from_samples(alph, blosum, -10, np.ones(len(alph) - 1))
    with pytest.raises(ValueError):
        bt.EValueEstimator.from_samples(alph, blosum, -10, -np.ones(len(alph)))
    skewed = blosum.score_matrix().copy()
    skewed[0, 1] += 1
    with pytest.raises(ValueError):          # symmetric matrices only
        bt.EValueEstimator.from_samples(alph, bt.SubstitutionMatrix(alph, alph, skewed), -10, np.ones(len(alph)))
    with pytest.raises(ValueError):          # the matrix must cover the alphabet
        bt.EValueEstimator.from_samples(alph, bt.SubstitutionMatrix.std_nucleotide_matrix(), -10, np.ones(len(alph)))


def test_log_evalue_formula():
    est = bt.EValueEstimator(0.267, 0.041)          # BLOSUM62, (-11, -1) in Altschul & Gish
    assert est.lam == 0.267 and est.k == 0.041
    expected = np.log10(0.041 * 300 * 3_000_000) - 0.267 * 50 / np.log(10)
    assert est.log_evalue(50, 300, 3_000_000) == pytest.approx(expected)
    many = est.log_evalue(np.array([30, 40, 50]), 300, 3_000_000)
    assert many.shape == (3,) and many[2] == pytest.approx(expected) and np.all(np.diff(many) < 0)
    # ten points of score at lambda 0.267 are 10 * 0.267 / ln 10 decades
    assert many[0] - many[1] == pytest.approx(10 * 0.267 / np.log(10))
