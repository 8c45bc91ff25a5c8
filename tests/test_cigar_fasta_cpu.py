"""
Host-side CIGAR / gapped-FASTA functions: the known answers of the reference's docstrings
(src/biotite/sequence/align/cigar.py:100-150, :272-314) and the two round-trip tests of the
reference's tests/sequence/align/test_cigar.py, with alignments computed by the CPU oracle.
"""
import re

import numpy as np
import pytest

import biotite_b200 as bt
from tests.util import oracle_align_optimal

REF = "TATAAAAGGTTTCCGACCGTAGGTAGCTGA"
SEG = "CCCCGGTTTGACCGTATGTAG"


def test_docstring_known_answers():
    ref, seg = bt.NucleotideSequence(REF), bt.NucleotideSequence(SEG)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    assert str(bt.read_alignment_from_cigar("9M2D12M", 3, ref, seg)) == \
        "AAAAGGTTTCCGACCGTAGGTAG\nCCCCGGTTT--GACCGTATGTAG"
    assert str(bt.read_alignment_from_cigar("4X5=2D7=1X4=", 3, ref, seg)) == \
        "AAAAGGTTTCCGACCGTAGGTAG\nCCCCGGTTT--GACCGTATGTAG"
    assert str(bt.read_alignment_from_cigar("3D9M2D12M4D", 0, ref, seg)) == \
        "TATAAAAGGTTTCCGACCGTAGGTAGCTGA\n---CCCCGGTTT--GACCGTATGTAG----"
    assert str(bt.read_alignment_from_cigar("4S5M2D12M", 7, ref, seg)) == \
        "GGTTTCCGACCGTAGGTAG\nGGTTT--GACCGTATGTAG"
    hard = bt.NucleotideSequence("GGTTTGACCGTATGTAG")
    assert str(bt.read_alignment_from_cigar("4H5M2D12M", 7, ref, hard)) == \
        "GGTTTCCGACCGTAGGTAG\nGGTTT--GACCGTATGTAG"
    tuples = [(bt.CigarOp.MATCH, 9), (bt.CigarOp.DELETION, 2), (bt.CigarOp.MATCH, 12)]
    assert str(bt.read_alignment_from_cigar(tuples, 3, ref, seg)) == \
        "AAAAGGTTTCCGACCGTAGGTAG\nCCCCGGTTT--GACCGTATGTAG"

    semi = oracle_align_optimal(ref, seg, matrix, terminal_penalty=False, max_number=1)[0]
    assert str(semi) == "TATAAAAGGTTTCCGACCGTAGGTAGCTGA\n---CCCCGGTTT--GACCGTATGTAG----"
    assert bt.write_alignment_to_cigar(semi) == "9M2D12M"
    assert bt.write_alignment_to_cigar(semi, introns=[(12, 14)]) == "9M2N12M"
    assert bt.write_alignment_to_cigar(semi, distinguish_matches=True) == "4X5=2D7=1X4="
    assert bt.write_alignment_to_cigar(semi, include_terminal_gaps=True) == "3D9M2D12M4D"
    local = oracle_align_optimal(ref, seg, matrix, local=True, max_number=1)[0]
    assert str(local) == "GGTTTCCGACCGTAGGTAG\nGGTTT--GACCGTATGTAG"
    assert bt.write_alignment_to_cigar(local, hard_clip=False) == "4S5M2D12M"
    assert bt.write_alignment_to_cigar(local, hard_clip=True) == "4H5M2D12M"
    ops = bt.write_alignment_to_cigar(semi, as_string=False)
    assert [(bt.CigarOp(op).name, n) for op, n in ops] == [("MATCH", 9), ("DELETION", 2), ("MATCH", 12)]


def test_errors():
    ref, seg = bt.NucleotideSequence(REF), bt.NucleotideSequence(SEG)
    ali = bt.read_alignment_from_cigar("9M2D12M", 3, ref, seg)
    with pytest.raises(ValueError):
        bt.write_alignment_to_cigar(ali, introns=[(5, 5)])
    with pytest.raises(ValueError):
        bt.write_alignment_to_cigar(ali, introns=[(-1, 5)])
    with pytest.raises(ValueError):
        bt.write_alignment_to_cigar(ali, introns=[(3, 6)])   # not inside a reference gap
    with pytest.raises(ValueError):
        bt.read_alignment_from_cigar(np.zeros((3,), dtype=int), 0, ref, seg)
    with pytest.raises(ValueError):
        bt.read_alignment_from_cigar("3P", 0, ref, seg)
    assert bt.CigarOp.from_cigar_symbol("=") == bt.CigarOp.EQUAL
    assert bt.CigarOp.INTRON.to_cigar_symbol() == "N"


def _generate_cigar(seed):
    rng = np.random.default_rng(seed)
    cigar = ""
    for _ in range(rng.choice(range(1, 10))):
        cigar += f"{rng.integers(1, 100)}M"
        cigar += f"{rng.integers(1, 100)}{'ID'[rng.integers(0, 2)]}"
    cigar += f"{rng.integers(1, 100)}M"
    if rng.choice([True, False]):
        cigar = f"{rng.integers(1, 100)}S" + cigar
    return cigar


@pytest.mark.parametrize("cigar", [_generate_cigar(i) for i in range(20)])
def test_cigar_conversion(cigar):
    """CIGAR -> Alignment -> CIGAR (reference test_cigar.py:78-99)."""
    ref = bt.NucleotideSequence("A" * 4000)
    seg = bt.NucleotideSequence("A" * 4000)
    alignment = bt.read_alignment_from_cigar(cigar, 0, ref, seg)
    test = re.sub(r"\d+S$", "", bt.write_alignment_to_cigar(alignment))
    assert test == cigar


def _mutate(original, rng, count=30):
    code = original.code.copy()
    idx = rng.choice(len(code), rng.integers(count), replace=False)
    code[idx] = rng.integers(len(original.alphabet), size=len(idx))
    idx = rng.choice(len(code), rng.integers(count), replace=False)
    code = np.insert(code, idx, rng.integers(len(original.alphabet), size=len(idx)))
    code = np.delete(code, rng.choice(len(code), rng.integers(count), replace=False))
    mutant = bt.NucleotideSequence(ambiguous=False)
    mutant.code = code.astype(np.uint8)
    return mutant


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("distinguish_matches", [False, True])
@pytest.mark.parametrize("include_terminal_gaps", [False, True])
def test_alignment_conversion(seed, local, distinguish_matches, include_terminal_gaps):
    """Alignment -> CIGAR -> Alignment (reference test_cigar.py:102-153)."""
    rng = np.random.default_rng(seed)
    ref = bt.NucleotideSequence(ambiguous=False)
    ref.code = rng.integers(0, len(ref.alphabet), 400, dtype=np.uint8)
    seg = _mutate(ref[rng.integers(0, 80):rng.integers(320, 400)], rng)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    if local:
        ali = oracle_align_optimal(ref, seg, matrix, local=True, max_number=1)[0]
    else:
        ali = oracle_align_optimal(ref, seg, matrix, terminal_penalty=False, max_number=1)[0]
        if not include_terminal_gaps:
            ali = bt.remove_terminal_gaps(ali)
    ali.score = None
    cigar = bt.write_alignment_to_cigar(ali, distinguish_matches=distinguish_matches,
                                        include_terminal_gaps=include_terminal_gaps)
    test = bt.read_alignment_from_cigar(cigar, ali.trace[0, 0], ref, seg)
    assert test == ali


def test_fasta_round_trip():
    ref, seg = bt.NucleotideSequence(REF * 4), bt.NucleotideSequence(SEG * 4)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    ali = oracle_align_optimal(ref, seg, matrix, terminal_penalty=False, max_number=1)[0]
    text = bt.alignment_to_fasta(ali, ["reference", "segment"])
    lines = text.splitlines()
    assert lines[0] == ">reference" and max(len(x) for x in lines) <= 80
    back = bt.alignment_from_fasta(text, lambda s: bt.NucleotideSequence(s))
    assert back.get_gapped_sequences() == ali.get_gapped_sequences()
    with pytest.raises(ValueError):
        bt.alignment_to_fasta(ali, ["only one"])
