"""Host-side alignment container and its helpers (reference ``alignment.py``): traces from
gapped strings, code / symbol tables, identities, the independent re-scorer, terminal gaps."""
import numpy as np
import pytest

import biotite_b200 as bt


def _alignment(rows):
    return bt.Alignment.from_strings(rows, bt.ProteinSequence)


def _random_alignment(rng, n_seq, n_col, gap_rate=0.25):
    """Rows over a 5-letter subset (many chance matches); every column keeps one symbol."""
    letters = np.array(list("ACDEG"))
    table = letters[rng.integers(0, 5, (n_seq, n_col))]
    gaps = rng.random((n_seq, n_col)) < gap_rate
    gaps[rng.integers(0, n_seq, n_col), np.arange(n_col)] = False
    for r in range(n_seq):                       # a symbol at both ends of some rows, none in others
        if not (~gaps[r]).any():
            gaps[r, rng.integers(0, n_col)] = False
    table[gaps] = "-"
    return _alignment(["".join(row) for row in table])


def test_trace_round_trip_and_tables():
    ali = _alignment(["AC-DE", "A-GDE", "--GD-"])
    assert ali.trace.tolist() == [[0, 0, -1], [1, -1, -1], [-1, 1, 0], [2, 2, 1], [3, 3, -1]]
    assert [str(s) for s in ali.sequences] == ["ACDE", "AGDE", "GD"]
    assert ali.get_gapped_sequences() == ["AC-DE", "A-GDE", "--GD-"]
    codes = bt.get_codes(ali)
    assert codes.shape == (3, 5) and codes[2].tolist()[:2] == [-1, -1]
    assert bt.get_symbols(ali)[1] == ["A", None, "G", "D", "E"]
    sub = ali[1:4, [0, 2]]
    assert sub.get_gapped_sequences() == ["C-D", "-GD"]
    assert bt.Alignment.trace_from_strings(["AC-DE", "A-GDE", "--GD-"]).tolist() == ali.trace.tolist()


def test_identity_modes():
    ali = _alignment(["--ACDE", "GAACE-"])
    # matching columns: A, C (E is shifted); 6 columns, 3 inside both sequences, shortest has 4 symbols
    assert bt.get_sequence_identity(ali, "all") == pytest.approx(2 / 6)
    assert bt.get_sequence_identity(ali, "not_terminal") == pytest.approx(2 / 3)
    assert bt.get_sequence_identity(ali, "shortest") == pytest.approx(2 / 4)
    with pytest.raises(ValueError):
        bt.get_sequence_identity(ali, "nonsense")
    with pytest.raises(ValueError):
        bt.get_sequence_identity(_alignment(["AC--", "--DE"]), "not_terminal")


@pytest.mark.parametrize("mode", ["all", "not_terminal", "shortest"])
def test_pairwise_identity_is_the_identity_of_every_row_pair(mode):
    rng = np.random.default_rng(7)
    for n_seq, n_col in ((2, 12), (5, 40), (9, 25)):
        ali = _random_alignment(rng, n_seq, n_col, gap_rate=0.15)
        got = bt.get_pairwise_sequence_identity(ali, mode)
        assert got.shape == (n_seq, n_seq) and np.allclose(got, got.T)
        for i in range(n_seq):
            for j in range(n_seq):
                pair = ali[:, [i, j]]
                if mode == "all":     # the reference divides by the columns of the whole alignment
                    expected = bt.get_sequence_identity(pair, "all") * len(pair) / len(ali)
                else:
                    expected = bt.get_sequence_identity(pair, mode)
                assert got[i, j] == pytest.approx(expected)
    with pytest.raises(ValueError):
        bt.get_pairwise_sequence_identity(_alignment(["AC--", "--DE"]), "not_terminal")
    with pytest.raises(ValueError):
        bt.get_pairwise_sequence_identity(ali, "nonsense")


def test_terminal_gaps():
    ali = _alignment(["--ACDE-", "GAAC-EW", "-AACDE-"])
    assert bt.find_terminal_gaps(ali) == (2, 6)
    assert bt.remove_terminal_gaps(ali).get_gapped_sequences() == ["ACDE", "AC-E", "ACDE"]
    # the docstring example of the reference (alignment.py:723-745)
    doc = bt.Alignment.from_strings(["AAAAACTGATTC---", "--AAACTG-TTCA--", "-----CTGATTCAAA"], bt.NucleotideSequence)
    assert bt.find_terminal_gaps(doc) == (5, 12)
    assert bt.remove_terminal_gaps(doc).get_gapped_sequences() == ["CTGATTC", "CTG-TTC", "CTGATTC"]
    # touching ends give an empty alignment, a hole between the sequences is an error (as in the reference)
    assert len(bt.remove_terminal_gaps(_alignment(["AC--", "--DE"]))) == 0
    with pytest.raises(ValueError):
        bt.remove_terminal_gaps(_alignment(["AC---", "---DE"]))


def test_rescorer_counts_opens_extensions_and_terminal_gaps():
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    ali = _alignment(["ACD--EGA", "-CDAAEG-"])
    m = matrix.score_matrix()
    alph = bt.ProteinSequence.alphabet
    pairs = sum(int(m[alph.encode(a), alph.encode(b)]) for a, b in ("CC", "DD", "EE", "GG"))
    # gaps: one leading and one trailing column in the second row, one run of two in the first
    assert bt.score(ali, matrix, (-10, -1), terminal_penalty=True) == pairs + 3 * -10 + -1
    assert bt.score(ali, matrix, (-10, -1), terminal_penalty=False) == pairs + -10 + -1
    assert bt.score(ali, matrix, -7, terminal_penalty=True) == pairs + 4 * -7
    with pytest.raises(TypeError):
        bt.score(ali, matrix, 1.5)


def test_rescorer_agrees_with_the_oracle_on_random_pairs():
    from oracle import oracle
    rng = np.random.default_rng(3)
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    for gap, terminal in (((-10, -1), True), ((-10, -1), False), (-6, True), (-6, False)):
        for _ in range(10):
            a, b = bt.ProteinSequence(), bt.ProteinSequence()
            a.code = rng.integers(0, 20, rng.integers(5, 60)).astype(np.uint8)
            b.code = rng.integers(0, 20, rng.integers(5, 60)).astype(np.uint8)
            traces, best = oracle.align_optimal(a.code, b.code, matrix.score_matrix(), gap, terminal, False, False, 1)
            ali = bt.Alignment([a, b], traces[0], best)
            assert bt.score(ali, matrix, gap, terminal) == best


def test_align_ungapped():
    """The known answer of the reference's test (tests/sequence/align/test_pairwise.py:14-23) and the error paths."""
    seq1, seq2 = bt.NucleotideSequence("ACCTGA"), bt.NucleotideSequence("ACTGGT")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    ali = bt.align_ungapped(seq1, seq2, matrix)
    assert ali.score == 3 and str(ali) == "ACCTGA\nACTGGT"
    assert ali.trace.tolist() == [[k, k] for k in range(6)]
    assert bt.align_ungapped(seq1, seq2, matrix, score_only=True) == 3
    assert bt.align_ungapped(seq1, seq2, (2, -1), score_only=True) == 3 * 2 + 3 * -1
    with pytest.raises(ValueError):
        bt.align_ungapped(seq1, bt.NucleotideSequence("ACT"), matrix)
    with pytest.raises(ValueError):
        bt.align_ungapped(bt.ProteinSequence("ACDEFG"), seq2, matrix)


def test_reference_known_answers():
    """tests/sequence/align/test_alignment.py:11-57: strings round trip, symbol tables, identity values."""
    rows = ["A-CCTGA----", "----T-ATGCT"]
    assert str(bt.Alignment.from_strings(rows, bt.NucleotideSequence)).split("\n") == rows
    rows = ["HAKLPRDD--WKL--", "HA--PRDDADWKLHH", "HA----DDADWKLHH"]
    symbols = bt.get_symbols(_alignment(rows))
    assert ["".join(s if s is not None else "-" for s in row) for row in symbols] == rows
    ali = _alignment(["--HAKLPRDD--WL--", "FRHA--QRTDADWLHH"])
    for mode, value in zip(("all", "not_terminal", "shortest"), (6 / 16, 6 / 12, 6 / 10)):
        assert bt.get_sequence_identity(ali, mode=mode) == value
        assert bt.get_pairwise_sequence_identity(ali, mode=mode)[0, 1] == value
    assert np.all(np.diag(bt.get_pairwise_sequence_identity(ali, "shortest")) == 1)
