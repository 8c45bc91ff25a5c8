"""
align_local_gapped on the GPU (csrc/xdrop_kernel.cuh through bt_xdrop_regions) against the
reference's golden alignments, its test cases (tests/sequence/align/test_localgapped.py) and
the CPU oracle on randomised inputs.  Bit-exact scores and traces, same co-optimal order.
"""
import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle
from tests.util import cas9_sequences, gap_param, golden, oracle_align_optimal

pytestmark = pytest.mark.gpu


def _cases():
    return [c for c in golden()["cases"] if c["method"] == "align_local_gapped"]


@pytest.mark.parametrize("case", _cases()[::3], ids=lambda c: c["file"])
def test_golden_single_calls(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    prm = case["params"]
    ali = bt.align_local_gapped(seqs[i], seqs[j], bt.SubstitutionMatrix.std_protein_matrix(), tuple(prm["seed"]),
                                prm["threshold"], gap_param(prm["gap_penalty"]), max_number=1)[0]
    assert ali.get_gapped_sequences() == case["gapped"]


def test_golden_all_cases_batched():
    """All 90 legacy alignments in two native calls (affine, linear)."""
    seqs = cas9_sequences()
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    for affine in (True, False):
        cases = [c for c in _cases() if isinstance(c["params"]["gap_penalty"], list) == affine]
        gap = gap_param(cases[0]["params"]["gap_penalty"])
        assert all(gap_param(c["params"]["gap_penalty"]) == gap for c in cases)
        alis = bt.align_local_gapped_batch([seqs[c["seq_indices"][0]] for c in cases],
                                           [seqs[c["seq_indices"][1]] for c in cases], matrix,
                                           [tuple(c["params"]["seed"]) for c in cases],
                                           cases[0]["params"]["threshold"], gap)
        for c, ali in zip(cases, alis):
            assert ali.get_gapped_sequences() == c["gapped"], c["file"]
            assert bt.score(ali, matrix, gap) == ali.score
        scores = bt.align_local_gapped_batch([seqs[c["seq_indices"][0]] for c in cases],
                                             [seqs[c["seq_indices"][1]] for c in cases], matrix,
                                             [tuple(c["params"]["seed"]) for c in cases],
                                             cases[0]["params"]["threshold"], gap, score_only=True)
        assert [int(x) for x in scores] == [a.score for a in alis]


@pytest.mark.parametrize("gap_penalty", [-10, (-10, -1)])
@pytest.mark.parametrize("seed", [(0, 0), (11, 11), (20, 19), (30, 29)])
@pytest.mark.parametrize("threshold", [20, 100, 500])
@pytest.mark.parametrize("direction", ["both", "upstream", "downstream"])
@pytest.mark.parametrize("score_only", [False, True])
def test_simple_alignment(gap_penalty, seed, threshold, direction, score_only):
    """Reference test_localgapped.py:10-50 (align_optimal computed by the oracle)."""
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
    result = bt.align_local_gapped(seq1, seq2, matrix, seed, threshold, gap_penalty, 1000, direction, score_only)
    if score_only:
        assert result == refs[0].score
    else:
        assert len(result) == len(refs)
        for ali in result:
            assert ali in refs


@pytest.mark.parametrize("gap_penalty", [-10, (-10, -1)])
@pytest.mark.parametrize("direction", ["both", "upstream", "downstream"])
@pytest.mark.parametrize("score_only", [False, True])
@pytest.mark.parametrize("should_raise", [False, True])
def test_max_table_size(gap_penalty, direction, score_only, should_raise):
    """Reference test_localgapped.py:53-111."""
    rng = np.random.default_rng(0)
    seq1 = bt.NucleotideSequence()
    seq1.code = rng.integers(len(seq1.alphabet), size=10000).astype(np.uint8)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    seed = (len(seq1) // 2, len(seq1) // 2)
    if should_raise:
        with pytest.raises(MemoryError):
            bt.align_local_gapped(seq1, seq1, matrix, seed, 100, gap_penalty, 1, direction, score_only, 1_000_000)
    else:
        result = bt.align_local_gapped(seq1, seq1, matrix, seed, 100, gap_penalty, 1, direction, score_only,
                                       1_000_000_000)
        if not score_only and direction == "both":
            assert len(result[0]) == len(seq1)


def _mutated_pair(rng, n, alphabet=20):
    a = rng.integers(0, alphabet, n).astype(np.uint8)
    b = a.copy()
    flip = rng.random(n) < 0.15
    b[flip] = rng.integers(0, alphabet, int(flip.sum()))
    keep = rng.random(n) > 0.03
    b = b[keep]
    ins = rng.random(len(b)) < 0.03
    b = np.insert(b, np.nonzero(ins)[0], rng.integers(0, alphabet, int(ins.sum()))).astype(np.uint8)
    return a, b


@pytest.mark.parametrize("gap_penalty", [-8, (-9, -2), (-3, -3), (-12, -1)])
@pytest.mark.parametrize("threshold", [0, 15, 60, 400])
def test_random_regions_match_oracle(gap_penalty, threshold):
    """align_region itself: scores, co-optimal traces and their order, tuple scoring included."""
    rng = np.random.default_rng(7)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    from biotite_b200.localgapped import _native_align_regions
    regions1, regions2 = [], []
    for k in range(24):
        a, b = _mutated_pair(rng, int(rng.integers(1, 400)))
        if k % 5 == 0:
            b = rng.integers(0, 20, int(rng.integers(0, 60))).astype(np.uint8)   # unrelated / empty
        regions1.append(a)
        regions2.append(b)
    for scoring in (matrix, (3, -2)):
        traces, scores = _native_align_regions(regions1, regions2, scoring, gap_penalty, threshold, 20, False,
                                               2**62, np.uint8)
        only = _native_align_regions(regions1, regions2, scoring, gap_penalty, threshold, 20, True, 2**62, np.uint8)[1]
        for k in range(len(regions1)):
            ref_traces, ref_score = oracle.align_region(regions1[k], regions2[k], scoring, gap_penalty, threshold,
                                                        20, False, 2**62)
            assert scores[k] == ref_score and only[k] == ref_score, k
            assert len(traces[k]) == len(ref_traces), k
            for t, r in zip(traces[k], ref_traces):
                assert np.array_equal(t, r), k


def test_wide_codes_and_long_regions():
    rng = np.random.default_rng(3)
    a, b = _mutated_pair(rng, 3000, alphabet=300)
    a, b = a.astype(np.uint16), b.astype(np.uint16)
    from biotite_b200.localgapped import _native_align_regions
    traces, scores = _native_align_regions([a], [b], (5, -4), (-10, -1), 80, 1, False, 2**62, np.uint16)
    ref_traces, ref_score = oracle.align_region(a, b, (5, -4), (-10, -1), 80, 1, False, 2**62)
    assert scores[0] == ref_score and np.array_equal(traces[0][0], ref_traces[0])


def test_argument_errors():
    seq = bt.ProteinSequence("ACDEFGHIKLMNPQRSTVWY")
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    with pytest.raises(ValueError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), 10, gap_penalty=0)
    with pytest.raises(TypeError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), 10, gap_penalty=-1.5)
    with pytest.raises(ValueError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), 10, max_number=0)
    with pytest.raises(ValueError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), -1)
    with pytest.raises(ValueError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), 10, direction="sideways")
    with pytest.raises(ValueError):
        bt.align_local_gapped(seq, seq, matrix, (3, 3), 10, max_table_size=0)
    with pytest.raises(IndexError):
        bt.align_local_gapped(seq, seq, matrix, (-1, 3), 10)
    with pytest.raises(IndexError):
        bt.align_local_gapped(seq, seq, matrix, (3, 20), 10)
    only_seed = bt.align_local_gapped(seq, seq, matrix, (0, 5), 10, direction="upstream")
    assert len(only_seed) == 1 and only_seed[0].trace.tolist() == [[0, 5]]
