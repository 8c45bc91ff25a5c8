"""
Pins the CPU oracle against everything the reference's own tests hold for the
align_optimal / align_banded path:

* the 360 + 180 legacy golden alignments (reference test
  ``tests/sequence/align/test_align.py:234-254``),
* score self-consistency with the independent re-scorer (``test_align.py:71-102``),
* the 7 known-answer cases with full co-optimal sets
  (``tests/sequence/align/test_pairwise.py:27-72``),
* docstring outputs (``pairwise.py:201-212``, ``alignment.py:81-96``,
  ``banded.py:181-200``) and the README alignment (``README.rst:100-115``).
"""

import numpy as np
import pytest

import biotite_b200 as bt
from tests.util import (
    cas9_sequences,
    gap_param,
    golden,
    oracle_align_banded,
    oracle_align_optimal,
)


def _cases(method):
    return [c for c in golden()["cases"] if c["method"] == method]


@pytest.mark.parametrize("case", _cases("align_optimal"), ids=lambda c: c["file"])
def test_golden_align_optimal(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    p = case["params"]
    gap = gap_param(p["gap_penalty"])
    ali = oracle_align_optimal(
        seqs[i], seqs[j], matrix, gap_penalty=gap, terminal_penalty=p["terminal_penalty"],
        local=p["local"], max_number=1)[0]
    assert ali.get_gapped_sequences() == case["gapped"]
    # the FASTA files carry no score: pin it through the independent re-scorer
    assert ali.score == bt.score(ali, matrix, gap, p["terminal_penalty"])
    assert ali.score == oracle_align_optimal(
        seqs[i], seqs[j], matrix, gap_penalty=gap, terminal_penalty=p["terminal_penalty"],
        local=p["local"], score_only=True)


@pytest.mark.parametrize("case", _cases("align_banded"), ids=lambda c: c["file"])
def test_golden_align_banded(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    p = case["params"]
    gap = gap_param(p["gap_penalty"])
    ali = oracle_align_banded(
        seqs[i], seqs[j], matrix, tuple(p["band"]), gap_penalty=gap, local=p["local"],
        max_number=1)[0]
    assert ali.get_gapped_sequences() == case["gapped"]
    assert ali.score == bt.score(ali, matrix, gap, terminal_penalty=False)
    assert ali.score == oracle_align_banded(
        seqs[i], seqs[j], matrix, tuple(p["band"]), gap_penalty=gap, local=p["local"],
        score_only=True)


# (local, terminal_penalty, gap_penalty, seq1, seq2, expected co-optimal set)
KNOWN_ANSWERS = [
    (False, True, -7, "TATGGGTATCC", "TATGTATAA",
     {"TATGGGTATCC\nTATG--TATAA", "TATGGGTATCC\nTAT-G-TATAA", "TATGGGTATCC\nTAT--GTATAA"}, 13),
    (True, True, -6, "TATGGGTATCC", "TATGTATAA",
     {"TATGGGTAT\nTATG--TAT", "TATGGGTAT\nTAT-G-TAT", "TATGGGTAT\nTAT--GTAT"}, 23),
    (False, True, (-7, -1), "TACTATGGGTATCC", "TCATATGTATAA",
     {"TACTATGGGTATCC\nTCATATG--TATAA", "TACTATGGGTATCC\nTCATAT--GTATAA"}, 16),
    (True, True, (-7, -1), "TACTATGGGTATCC", "TCATATGTATAA",
     {"TATGGGTAT\nTATG--TAT", "TATGGGTAT\nTAT--GTAT"}, 27),
    (False, True, (-7, -1), "T", "TTT", {"T--\nTTT", "--T\nTTT"}, -3),
    (False, True, -7, "TAAAGCGAAAT", "TGCGT", {"TAAAGCGAAAT\nT---GCG---T"}, -17),
    (False, False, -7, "TAAAGCGAAAT", "TGCGT", {"TAAAGCGAAAT\n---TGCGT---"}, 7),
]


@pytest.mark.parametrize("local,term,gap,s1,s2,expect,score", KNOWN_ANSWERS)
def test_known_answers(local, term, gap, s1, s2, expect, score):
    seq1, seq2 = bt.NucleotideSequence(s1), bt.NucleotideSequence(s2)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    alis = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=gap,
                                terminal_penalty=term, local=local)
    assert {str(a) for a in alis} == expect
    assert len(alis) == len(expect)
    for a in alis:
        assert a.score == score == bt.score(a, matrix, gap, term)


def test_docstring_align_optimal_order():
    seq1 = bt.NucleotideSequence("ATACGCTTGCT")
    seq2 = bt.NucleotideSequence("AGGCGCAGCT")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    alis = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=-6)
    assert [str(a) for a in alis] == [
        "ATACGCTTGCT\nAGGCGCA-GCT",
        "ATACGCTTGCT\nAGGCGC-AGCT",
    ]
    assert alis[0].score == 17


def test_docstring_alignment_trace():
    seq1 = bt.NucleotideSequence("CGTCAT")
    seq2 = bt.NucleotideSequence("TCATGC")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    ali = oracle_align_optimal(seq1, seq2, matrix)[0]
    assert str(ali) == "CGTCAT--\n--TCATGC"
    assert ali.trace.tolist() == [
        [0, -1], [1, -1], [2, 0], [3, 1], [4, 2], [5, 3], [-1, 4], [-1, 5]]
    assert ali.score == -20
    assert bt.get_codes(ali).tolist() == [
        [1, 2, 3, 1, 0, 3, -1, -1], [-1, -1, 3, 1, 0, 3, 2, 1]]


def test_docstring_align_banded():
    # banded.py:181-200 (diagonal -6 found by the k-mer match, buffer 5)
    seq1 = bt.NucleotideSequence("GCGCGCTATATTATGCGCGC")
    seq2 = bt.NucleotideSequence("TATAAT")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    ali = oracle_align_banded(seq1, seq2, matrix, band=(-6 - 5, -6 + 5),
                              gap_penalty=(-6, -1))[0]
    assert str(ali) == "TATATTAT\nTATA--AT"


def test_docstring_statistics_local_affine():
    # statistics.py:65-76
    query = bt.NucleotideSequence("CGACGGCGTCTACGAGTCAACATCATTC")
    hit = bt.NucleotideSequence("GCTTTATTACGGGTTTACGAGTTCAACATCACGAAAACAA")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    ali = oracle_align_optimal(query, hit, matrix, (-12, -2), local=True)[0]
    assert str(ali) == "ACGGCGTCTACGAGT-CAACATCA\nACGG-GTTTACGAGTTCAACATCA"
    assert ali.score == 77


def test_readme_avidin_streptavidin():
    g = golden()["readme_avidin_streptavidin"]
    top, bottom = g["gapped"]
    seq1 = bt.ProteinSequence(top.replace("-", ""))
    seq2 = bt.ProteinSequence(bottom.replace("-", ""))
    assert (len(seq1), len(seq2)) == (152, 159)
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    alis = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=(-10, -1),
                                terminal_penalty=False)
    assert len(alis) == 6
    assert alis[0].get_gapped_sequences() == [top, bottom]
    assert alis[0].score == 133
    # with max_number=1 the all-first-priority path comes out, which is the LAST
    # of the enumeration above
    first = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=(-10, -1),
                                 terminal_penalty=False, max_number=1)[0]
    assert np.array_equal(first.trace, alis[-1].trace)


def test_empty_local_trace():
    seq1 = bt.NucleotideSequence("AAAA")
    seq2 = bt.NucleotideSequence("CCCC")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    for gap in (-5, (-5, -1)):
        alis = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=gap, local=True,
                                    max_number=3)
        assert len(alis) == 3
        assert all(a.trace.shape == (0, 2) and a.score == 0 for a in alis)


def test_reference_affine_is_three_state():
    # a gap in one sequence cannot be followed directly by a
    # gap in the other, so 'A' vs 'C' with cheap gaps scores the mismatch.
    seq1, seq2 = bt.NucleotideSequence("A"), bt.NucleotideSequence("C")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    assert oracle_align_optimal(seq1, seq2, matrix, gap_penalty=(-1, -1), score_only=True) == -4
    assert oracle_align_optimal(seq1, seq2, matrix, gap_penalty=-1, score_only=True) == -2
