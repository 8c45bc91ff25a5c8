"""
Pins the oracle's X-drop restatement (oracle/align_oracle.c run_region, following
src/rust/sequence/align/localgapped.rs) on the reference's own vectors: the 90
align_local_gapped legacy alignments (tests/sequence/data/legacy_consistency) and the
cyclotide comparison with align_optimal of tests/sequence/align/test_localgapped.py:10-50.
"""
import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle
from tests.util import cas9_sequences, gap_param, golden, oracle_align_optimal


def _cases():
    return [c for c in golden()["cases"] if c["method"] == "align_local_gapped"]


def test_golden_count():
    assert len(_cases()) == 90


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["file"])
def test_oracle_reproduces_legacy_alignment(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    prm = case["params"]
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    traces, score = oracle.align_local_gapped(seqs[i].code, seqs[j].code, matrix, tuple(prm["seed"]),
                                              prm["threshold"], gap_param(prm["gap_penalty"]), 1)
    ali = bt.Alignment([seqs[i], seqs[j]], traces[0], score)
    assert ali.get_gapped_sequences() == case["gapped"]
    # scores are pinned through the independent re-scorer (reference test_align.py:89-97)
    assert bt.score(ali, bt.SubstitutionMatrix.std_protein_matrix(), gap_param(prm["gap_penalty"])) == score


@pytest.mark.parametrize("gap_penalty", [-10, (-10, -1)])
@pytest.mark.parametrize("seed", [(0, 0), (11, 11), (20, 19), (30, 29)])
@pytest.mark.parametrize("threshold", [20, 100, 500])
@pytest.mark.parametrize("direction", ["both", "upstream", "downstream"])
def test_oracle_cyclotides_against_align_optimal(gap_penalty, seed, threshold, direction):
    seq1 = bt.ProteinSequence("gvpcaescvwipctvtallgcsckdkvcyld")
    seq2 = bt.ProteinSequence("gipcgescvfipcissvvgcsckskvcyld")
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    refs = oracle_align_optimal(seq1, seq2, matrix, gap_penalty=gap_penalty, local=True)
    for ali in refs:
        k = np.where(ali.trace[:, 0] == seed[0])[0][0]
        if direction == "upstream":
            ali.trace = ali.trace[:k + 1]
        elif direction == "downstream":
            ali.trace = ali.trace[k:]
        ali.score = bt.score(ali, matrix, gap_penalty)
    traces, score = oracle.align_local_gapped(seq1.code, seq2.code, matrix.score_matrix(), seed, threshold,
                                              gap_penalty, 1000, direction)
    assert score == refs[0].score
    assert len(traces) == len(refs)
    for trace in traces:
        assert any(np.array_equal(trace, ref.trace) for ref in refs)
    _, only = oracle.align_local_gapped(seq1.code, seq2.code, matrix.score_matrix(), seed, threshold,
                                        gap_penalty, 1000, direction, score_only=True)
    assert only == score


def test_oracle_max_table_size():
    rng = np.random.default_rng(0)
    code = rng.integers(0, 4, 3000)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()
    with pytest.raises(MemoryError):
        oracle.align_local_gapped(code, code, matrix, (1500, 1500), 100, -10, 1, "both", False, 100_000)
    traces, _ = oracle.align_local_gapped(code, code, matrix, (1500, 1500), 100, -10, 1, "both", False, 10**9)
    assert len(traces[0]) == len(code)
