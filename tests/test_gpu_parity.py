"""
GPU parity tests proper: the CUDA path, called through the C ABI, against the
CPU oracle and against the committed golden vectors.  Bit-exact scores,
identical traces (integer work: no tolerance).
"""

import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle
from tests.test_oracle_golden import KNOWN_ANSWERS
from tests.util import (
    cas9_sequences,
    gap_param,
    golden,
    oracle_align_banded,
    oracle_align_optimal,
)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["packed", "general"])
def route(request, monkeypatch):
    """Run a batch test once per kernel family (B200ALIGN_ROUTE overrides the cost model)."""
    monkeypatch.setenv("B200ALIGN_ROUTE", request.param)
    return request.param

NUC = bt.SubstitutionMatrix.std_nucleotide_matrix
BLOSUM = bt.SubstitutionMatrix.std_protein_matrix


def _same(alis, refs):
    assert len(alis) == len(refs)
    for a, r in zip(alis, refs):
        assert a.score == r.score
        assert np.array_equal(a.trace, r.trace)


# ---- reference golden vectors (every 3rd pair to bound the run time; all param sets) ----

def _golden_subset(method):
    cases = [c for c in golden()["cases"] if c["method"] == method]
    pairs = sorted({tuple(c["seq_indices"]) for c in cases})[::3]
    return [c for c in cases if tuple(c["seq_indices"]) in pairs]


@pytest.mark.parametrize("case", _golden_subset("align_optimal"), ids=lambda c: c["file"])
def test_golden_align_optimal(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    p = case["params"]
    gap = gap_param(p["gap_penalty"])
    ali = bt.align_optimal(seqs[i], seqs[j], BLOSUM(), gap_penalty=gap,
                           terminal_penalty=p["terminal_penalty"], local=p["local"],
                           max_number=1)[0]
    assert ali.get_gapped_sequences() == case["gapped"]
    assert ali.score == bt.score(ali, BLOSUM(), gap, p["terminal_penalty"])
    assert ali.score == bt.align_optimal(
        seqs[i], seqs[j], BLOSUM(), gap_penalty=gap, terminal_penalty=p["terminal_penalty"],
        local=p["local"], score_only=True)


@pytest.mark.parametrize("case", _golden_subset("align_banded"), ids=lambda c: c["file"])
def test_golden_align_banded(case):
    seqs = cas9_sequences()
    i, j = case["seq_indices"]
    p = case["params"]
    gap = gap_param(p["gap_penalty"])
    ali = bt.align_banded(seqs[i], seqs[j], BLOSUM(), tuple(p["band"]), gap_penalty=gap,
                          local=p["local"], max_number=1)[0]
    assert ali.get_gapped_sequences() == case["gapped"]
    assert ali.score == bt.score(ali, BLOSUM(), gap, terminal_penalty=False)


def test_golden_batch_all_cases():
    """All 540 golden alignments through the batched entry points."""
    seqs = cas9_sequences()
    cases = [c for c in golden()["cases"] if c["method"] in ("align_optimal", "align_banded")]
    assert len(cases) == 540
    groups = {}
    for c in cases:
        p = c["params"]
        key = (c["method"], str(p["gap_penalty"]), p["local"], p.get("terminal_penalty", True))
        groups.setdefault(key, []).append(c)
    for (method, _, local, term), group in groups.items():
        gap = gap_param(group[0]["params"]["gap_penalty"])
        s1 = [seqs[c["seq_indices"][0]] for c in group]
        s2 = [seqs[c["seq_indices"][1]] for c in group]
        if method == "align_optimal":
            res = bt.align_optimal_batch(s1, s2, BLOSUM(), gap, term, local)
        else:
            bands = np.array([c["params"]["band"] for c in group])
            res = bt.align_banded_batch(s1, s2, BLOSUM(), bands, gap, local)
        for k, c in enumerate(group):
            assert res.alignment(k).get_gapped_sequences() == c["gapped"], c["file"]


# ---- known answers: full co-optimal sets, ordered enumeration ------------------------------

@pytest.mark.parametrize("local,term,gap,s1,s2,expect,score", KNOWN_ANSWERS)
def test_known_answers(local, term, gap, s1, s2, expect, score):
    seq1, seq2 = bt.NucleotideSequence(s1), bt.NucleotideSequence(s2)
    alis = bt.align_optimal(seq1, seq2, NUC(), gap_penalty=gap, terminal_penalty=term, local=local)
    refs = oracle_align_optimal(seq1, seq2, NUC(), gap_penalty=gap, terminal_penalty=term, local=local)
    assert {str(a) for a in alis} == expect
    _same(alis, refs)
    assert all(a.score == score for a in alis)


def test_readme_example_enumeration_order():
    g = golden()["readme_avidin_streptavidin"]
    top, bottom = g["gapped"]
    seq1 = bt.ProteinSequence(top.replace("-", ""))
    seq2 = bt.ProteinSequence(bottom.replace("-", ""))
    alis = bt.align_optimal(seq1, seq2, BLOSUM(), gap_penalty=(-10, -1), terminal_penalty=False)
    assert len(alis) == 6 and alis[0].score == 133
    assert alis[0].get_gapped_sequences() == [top, bottom]


# ---- randomised property tests against the oracle ----------------------------------------------

def _random_seq(rng, n, alphabet_size=4, cls=None):
    if cls is None:
        seq = bt.NucleotideSequence()
    else:
        seq = cls()
    seq.code = rng.integers(0, alphabet_size, n)
    return seq


@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("term", [False, True])
@pytest.mark.parametrize("gap", [-5, -8, (-8, -8), (-7, -1), (-1, -1), (0, 0), (-3, 0)])
@pytest.mark.parametrize("seed", range(3))
def test_random_nucleotide_pairs(local, term, gap, seed):
    rng = np.random.default_rng(seed)
    seq1 = _random_seq(rng, rng.integers(1, 60))
    seq2 = _random_seq(rng, rng.integers(1, 60))
    for scoring in (NUC(), (5, -5)):
        alis = bt.align_optimal(seq1, seq2, scoring, gap_penalty=gap, terminal_penalty=term,
                                local=local, max_number=50)
        refs = oracle_align_optimal(seq1, seq2, scoring, gap_penalty=gap, terminal_penalty=term,
                                    local=local, max_number=50)
        _same(alis, refs)
        assert bt.align_optimal(seq1, seq2, scoring, gap_penalty=gap, terminal_penalty=term,
                                local=local, score_only=True) == refs[0].score


@pytest.mark.parametrize("n,m", [(0, 0), (0, 5), (5, 0), (1, 1), (1, 7), (7, 1)])
@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("gap", [-6, (-6, -2)])
def test_edge_shapes(n, m, local, gap):
    rng = np.random.default_rng(n * 10 + m)
    seq1, seq2 = _random_seq(rng, n), _random_seq(rng, m)
    for term in (True, False):
        alis = bt.align_optimal(seq1, seq2, NUC(), gap_penalty=gap, terminal_penalty=term,
                                local=local, max_number=5)
        refs = oracle_align_optimal(seq1, seq2, NUC(), gap_penalty=gap, terminal_penalty=term,
                                    local=local, max_number=5)
        _same(alis, refs)


def test_all_ties_and_all_negative():
    a = bt.NucleotideSequence("AAAAAAAA")
    c = bt.NucleotideSequence("CCCCCC")
    for seq1, seq2 in ((a, a), (a, c)):
        for local in (False, True):
            for gap in (-4, (-4, -4), (-5, -1)):
                alis = bt.align_optimal(seq1, seq2, NUC(), gap_penalty=gap, local=local, max_number=30)
                refs = oracle_align_optimal(seq1, seq2, NUC(), gap_penalty=gap, local=local, max_number=30)
                _same(alis, refs)


def test_wide_codes_and_positional_matrix():
    rng = np.random.default_rng(5)
    p1 = _random_seq(rng, 70, 20, bt.ProteinSequence)
    p2 = _random_seq(rng, 90, 20, bt.ProteinSequence)
    pos_matrix, q1, q2 = BLOSUM().as_positional(p1, p2)
    plain = bt.align_optimal(p1, p2, BLOSUM(), gap_penalty=(-10, -1), max_number=1)[0]
    # PositionalSequence codes are uint8 here; force a wide dtype through a big alphabet
    alph = bt.Alphabet(range(70000))
    w1 = bt.GeneralSequence(alph); w1.code = q1.code
    w2 = bt.GeneralSequence(alph); w2.code = q2.code
    big = np.zeros((70, 90), dtype=np.int32) + pos_matrix.score_matrix()
    from biotite_b200.pairwise import _native_align_pair
    traces, score = _native_align_pair(
        np.ascontiguousarray(w1.code, np.uint32), np.ascontiguousarray(w2.code, np.uint32),
        big, (-10, -1), True, False, False, 1)
    assert score == plain.score and np.array_equal(traces[0], plain.trace)
    pos = bt.align_optimal(q1, q2, pos_matrix, gap_penalty=(-10, -1), max_number=1)[0]
    assert pos.score == plain.score and np.array_equal(pos.trace, plain.trace)


def test_scaled_scores_need_int32():
    # test_statistics.py:118-177 scales scores x1000 -> 16-bit lanes would overflow
    rng = np.random.default_rng(11)
    s1 = _random_seq(rng, 300, 20, bt.ProteinSequence)
    s2 = _random_seq(rng, 300, 20, bt.ProteinSequence)
    alph = bt.ProteinSequence.alphabet
    scaled = bt.SubstitutionMatrix(alph, alph, BLOSUM().score_matrix() * 1000)
    for local in (False, True):
        got = bt.align_optimal(s1, s2, scaled, gap_penalty=(-10000, -1000), local=local, max_number=1)
        ref = oracle_align_optimal(s1, s2, scaled, gap_penalty=(-10000, -1000), local=local, max_number=1)
        _same(got, ref)
        res = bt.align_optimal_batch([s1, s2], [s2, s1], scaled, (-10000, -1000), True, local)
        assert res.scores[0] == ref[0].score and np.array_equal(res.trace(0), ref[0].trace)


def test_invalid_code_is_rejected():
    seq1 = bt.NucleotideSequence("ACGT")
    seq2 = bt.NucleotideSequence("ACGT")
    seq2.code = np.array([0, 1, 2, 200], dtype=np.uint8)
    with pytest.raises(ValueError, match="outside the substitution matrix"):
        bt.align_optimal(seq1, seq2, NUC())


# ---- banded -----------------------------------------------------------------------------------

@pytest.mark.parametrize("gap", [-10, (-10, -1)])
@pytest.mark.parametrize("local", [False, True])
@pytest.mark.parametrize("band_width", [2, 5, 20, 100])
def test_banded_cyclotide_like(gap, local, band_width):
    # same shape as tests/sequence/align/test_banded.py:11-43 (31 vs 30 residues)
    s1 = bt.ProteinSequence("gvpcgescvfipcisaaigcsckskvcyrnasdf".upper()[:31])
    s2 = bt.ProteinSequence("gipcgescvwipcitsaigcsckskvcyrnsd".upper()[:30])
    for a, b in ((s1, s2), (s2, s1)):
        got = bt.align_banded(a, b, BLOSUM(), (-band_width, band_width), gap_penalty=gap,
                              local=local, max_number=20)
        ref = oracle_align_banded(a, b, BLOSUM(), (-band_width, band_width), gap_penalty=gap,
                                  local=local, max_number=20)
        _same(got, ref)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("gap", [-7, (-7, -2)])
@pytest.mark.parametrize("local", [False, True])
def test_banded_random(seed, gap, local):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(5, 80)), int(rng.integers(5, 80))
    seq1, seq2 = _random_seq(rng, n), _random_seq(rng, m)
    lo = int(rng.integers(-n, m))
    band = (lo, lo + int(rng.integers(0, 30)))
    try:
        ref = oracle_align_banded(seq1, seq2, NUC(), band, gap_penalty=gap, local=local, max_number=10)
    except ValueError:
        with pytest.raises(ValueError):
            bt.align_banded(seq1, seq2, NUC(), band, gap_penalty=gap, local=local)
        return
    got = bt.align_banded(seq1, seq2, NUC(), band, gap_penalty=gap, local=local, max_number=10)
    _same(got, ref)
    assert bt.align_banded(seq1, seq2, NUC(), band, gap_penalty=gap, local=local,
                           score_only=True) == ref[0].score


def test_banded_long_excerpt_maps_back():
    # tests/sequence/align/test_banded.py:46-78 (shortened target)
    rng = np.random.default_rng(0)
    target = _random_seq(rng, 20000)
    start = 11111
    query = target[start:start + 500]
    diag = start
    alis = bt.align_banded(query, target, NUC(), (diag - 100 + 30, diag + 100 + 30))
    assert len(alis) == 1
    trace = alis[0].trace
    trace = trace[(trace != -1).all(axis=1)]
    assert np.array_equal(trace[:, 1] - trace[:, 0], np.full(len(trace), start))
    assert len(trace) == 500


# ---- batched entry points ---------------------------------------------------------------------------

def _random_batch(rng, n_pairs, lo, hi, alphabet=20):
    len1 = rng.integers(lo, hi, n_pairs)
    len2 = rng.integers(lo, hi, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, alphabet, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, alphabet, off2[-1]).astype(np.uint8)
    return codes1, off1, codes2, off2


@pytest.mark.parametrize("gap", [(-10, -1), -10, (-4, -4), (-11, 0)])
@pytest.mark.parametrize("local,term", [(False, True), (False, False), (True, True)])
def test_batch_matches_oracle(gap, local, term, route):
    rng = np.random.default_rng(42)
    codes1, off1, codes2, off2 = _random_batch(rng, 300, 1, 200)
    matrix = BLOSUM().score_matrix()
    # make a third of the pairs homologous so that traces have real structure
    for k in range(0, 300, 3):
        n = min(off1[k + 1] - off1[k], off2[k + 1] - off2[k])
        codes2[off2[k]:off2[k] + n] = codes1[off1[k]:off1[k] + n]
        flip = rng.random(n) < 0.15
        codes2[off2[k]:off2[k] + n][flip] = rng.integers(0, 20, int(flip.sum()))
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local) as batch:
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(300):
        assert np.array_equal(traces[k], ref_traces[k]), k
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local, score_only=True) as batch:
        batch.run()
        assert np.array_equal(batch.fetch().scores, ref_scores)
    # the pipelined one-shot entry point must agree with the resident-batch object
    one = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(one.scores, ref_scores)
    assert np.array_equal(one.trace_len, res.trace_len) and np.array_equal(one.first, res.first)
    assert np.array_equal(one.ops, res.ops) or all(
        np.array_equal(a, b) for a, b in zip(one.traces(), traces))


@pytest.mark.parametrize("local,term", [(False, True), (False, False), (True, True)])
@pytest.mark.parametrize("gap", [(-10, -1), (-12, -2), (-6, -6), -10, -3])
def test_batch_protein_300aa_all_modes(local, term, gap, monkeypatch):
    """Config-2/3/5 shaped pairs (200-400 aa, half homologous) through the packed kernels,
    every score and trace against the oracle; score-only variants too."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(7)
    n_pairs = 2500
    codes1, off1, codes2, off2 = _random_batch(rng, n_pairs, 200, 400)
    for k in range(0, n_pairs, 2):
        n = min(off1[k + 1] - off1[k], off2[k + 1] - off2[k])
        shift = int(rng.integers(0, 30))
        n = n - shift
        codes2[off2[k] + shift:off2[k] + shift + n] = codes1[off1[k]:off1[k] + n]
        flip = rng.random(n) < 0.2
        codes2[off2[k] + shift:off2[k] + shift + n][flip] = rng.integers(0, 20, int(flip.sum()))
    matrix = BLOSUM().score_matrix()
    linear_semi = isinstance(gap, int) and not local and not term  # no packed kernel for this one
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local) as batch:
        assert batch.kernel == ("general_i32" if linear_semi else "interseq_s16x2")
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(n_pairs):
        assert np.array_equal(traces[k], ref_traces[k]), k
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local, score_only=True) as batch:
        batch.run()
        assert np.array_equal(batch.fetch().scores, ref_scores)
    one = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, gap, term, local, score_only=True)
    assert np.array_equal(one.scores, ref_scores)


def test_batch_c2_shape_properties():
    """BASELINE config 2 shape (global affine BLOSUM62, ~300 aa) at a size the
    oracle cannot check exhaustively: re-score every trace instead."""
    rng = np.random.default_rng(1)
    n_pairs = 20000
    codes1, off1, codes2, off2 = _random_batch(rng, n_pairs, 200, 400)
    matrix = BLOSUM().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False) as batch:
        batch.run()
        res = batch.fetch()
    flat, rows_off = res.traces_flat()
    # every trace consumes both sequences completely, in order
    for col, off in ((0, off1), (1, off2)):
        present = flat[:, col] != -1
        counts = np.add.reduceat(present.astype(np.int64), rows_off[:-1])
        assert np.array_equal(counts, np.diff(off))
    # rescoring: substitution scores + affine gap costs == reported score
    pair_of_row = np.repeat(np.arange(n_pairs), np.diff(rows_off))
    a, b = flat[:, 0], flat[:, 1]
    both = (a != -1) & (b != -1)
    sub = np.zeros(len(flat), dtype=np.int64)
    sub[both] = matrix[codes1[off1[pair_of_row[both]] + a[both]], codes2[off2[pair_of_row[both]] + b[both]]]
    total = np.add.reduceat(sub, rows_off[:-1])
    for col in (a, b):
        gap = col == -1
        prev = np.concatenate(([False], gap[:-1]))
        prev[rows_off[:-1]] = False
        total += np.add.reduceat(np.where(gap, np.where(prev, -1, -10), 0), rows_off[:-1])
    assert np.array_equal(total, res.scores.astype(np.int64))
    # spot-check against the oracle
    idx = rng.choice(n_pairs, 64, replace=False)
    for k in idx:
        t, s = oracle.align_optimal(codes1[off1[k]:off1[k + 1]], codes2[off2[k]:off2[k + 1]],
                                    matrix, (-10, -1), True, False, False, 1)
        assert s == res.scores[k] and np.array_equal(t[0], res.trace(k))


def test_cost_model_routes_small_batches_to_general():
    rng = np.random.default_rng(5)
    matrix = BLOSUM().score_matrix()
    codes1, off1, codes2, off2 = _random_batch(rng, 256, 1000, 1400)
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, score_only=True) as b:
        assert b.kernel == "general_i32"
    codes1, off1, codes2, off2 = _random_batch(rng, 40000, 200, 300)
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, score_only=True) as b:
        assert b.kernel == "interseq_s16x2"


def test_pipelined_windows_match_resident_batch():
    """Several windows through both slots of the pipelined path (bt_align_batch)."""
    rng = np.random.default_rng(9)
    n_pairs = 700_000
    codes1, off1, codes2, off2 = _random_batch(rng, n_pairs, 20, 60)
    matrix = BLOSUM().score_matrix()
    one = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix, (-10, -1), True, False)
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False) as batch:
        assert batch.kernel == "interseq_s16x2"
        batch.run()
        res = batch.fetch()
    assert np.array_equal(one.scores, res.scores)
    assert np.array_equal(one.trace_len, res.trace_len)
    assert np.array_equal(one.first, res.first)
    idx = rng.choice(n_pairs, 200, replace=False)
    for k in idx:
        t, sc = oracle.align_optimal(codes1[off1[k]:off1[k + 1]], codes2[off2[k]:off2[k + 1]],
                                     matrix, (-10, -1), True, False, False, 1)
        assert sc == one.scores[k] and np.array_equal(t[0], one.trace(k)) and np.array_equal(t[0], res.trace(k))


@pytest.mark.parametrize("gap", [(-10, -1), (-3, -3), (-7, 0)])
def test_banded_interseq_random_bands(gap, monkeypatch):
    """The banded inter-sequence kernels on ragged inputs: tiny and unequal lengths, bands that
    start left of / end right of the sequences, width-1 bands; every score and trace vs the
    oracle (align_banded, semi-global, max_number=1)."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    from biotite_b200.banded import crop_band
    rng = np.random.default_rng(17)
    c1, c2, o1, o2, bands = [], [], [0], [0], []
    while len(bands) < 300:
        n = int(rng.integers(1, 130))
        m = n + int(rng.integers(0, 90))
        a = rng.integers(0, 4, n).astype(np.uint8)
        b = rng.integers(0, 4, m).astype(np.uint8)
        if rng.random() < 0.6:   # plant a shifted copy so that real alignments exist
            shift = int(rng.integers(0, m - n + 1))
            b[shift:shift + n] = a
            flip = rng.random(m) < 0.1
            b[flip] = rng.integers(0, 4, int(flip.sum()))
        lo = int(rng.integers(-n - 5, m + 5))
        hi = lo + int(rng.integers(0, 70))
        try:
            band = crop_band(n, m, (lo, hi))
        except ValueError:
            continue
        c1.append(a); c2.append(b); o1.append(o1[-1] + n); o2.append(o2[-1] + m); bands.append(band)
    codes1, codes2 = np.concatenate(c1), np.concatenate(c2)
    off1, off2 = np.array(o1, dtype=np.int64), np.array(o2, dtype=np.int64)
    bands = np.array(bands, dtype=np.int64)
    matrix = NUC().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, True, False, bands) as batch:
        assert batch.kernel == "banded_i32"
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, bands=bands)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(len(bands)):
        assert np.array_equal(traces[k], ref_traces[k]), (k, bands[k], off1[k + 1] - off1[k], off2[k + 1] - off2[k])
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, True, False, bands, score_only=True) as batch:
        batch.run()
        assert np.array_equal(batch.fetch().scores, ref_scores)


def test_banded_interseq_c4_shape(monkeypatch):
    """Config-4 shaped pairs (3 kb here, band +-64, NUC, affine) against the oracle."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(23)
    c1, c2, o1, o2 = [], [], [0], [0]
    for k in range(96):
        L = int(rng.integers(2500, 3200))
        a = rng.integers(0, 4, L).astype(np.uint8)
        b = a.copy()
        sub = rng.random(L) < 0.05
        b[sub] = rng.integers(0, 4, int(sub.sum()))
        b = b[rng.random(L) >= 0.006]
        ins = np.sort(rng.choice(len(b), 15, replace=False))
        b = np.insert(b, ins, rng.integers(0, 4, 15).astype(np.uint8))
        if len(b) < len(a):
            a, b = b, a
        c1.append(a); c2.append(b); o1.append(o1[-1] + len(a)); o2.append(o2[-1] + len(b))
    codes1, codes2 = np.concatenate(c1), np.concatenate(c2)
    off1, off2 = np.array(o1, dtype=np.int64), np.array(o2, dtype=np.int64)
    bands = np.array([bt.banded.crop_band(off1[k + 1] - off1[k], off2[k + 1] - off2[k], (-64, 64))
                      for k in range(96)], dtype=np.int64)
    matrix = NUC().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, bands) as batch:
        assert batch.kernel == "banded_i32"
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, (-10, -1), bands=bands)
    assert np.array_equal(res.scores, ref_scores)
    assert res.scores.max() > 10000   # beyond the packed 16-bit kernels' comfort zone
    traces = res.traces()
    for k in range(96):
        assert np.array_equal(traces[k], ref_traces[k]), k


def test_banded_batch_matches_oracle(route):
    rng = np.random.default_rng(3)
    n_pairs = 40
    seqs1, seqs2, bands = [], [], []
    for k in range(n_pairs):
        n = int(rng.integers(300, 700))
        s1 = _random_seq(rng, n)
        code2 = s1.code.copy()
        mut = rng.random(n) < 0.05
        code2[mut] = rng.integers(0, 4, int(mut.sum()))
        keep = rng.random(n) > 0.01
        s2 = bt.NucleotideSequence()
        s2.code = code2[keep]
        if k % 2:
            s1, s2 = s2, s1
        seqs1.append(s1); seqs2.append(s2); bands.append((-16, 16))
    for gap in ((-10, -1), -6):
        for local in (False, True):
            res = bt.align_banded_batch(seqs1, seqs2, NUC(), np.array(bands), gap, local)
            for k in range(n_pairs):
                ref = oracle_align_banded(seqs1[k], seqs2[k], NUC(), bands[k], gap_penalty=gap,
                                          local=local, max_number=1)[0]
                assert res.scores[k] == ref.score
                assert np.array_equal(res.trace(k), ref.trace)


# ---- indexed batches: one-vs-many (config 3) and all-vs-all (config 5) without materialised pairs ----

def _random_proteins(rng, count, lo, hi):
    seqs = []
    for _ in range(count):
        s = bt.ProteinSequence()
        s.code = rng.integers(0, 20, int(rng.integers(lo, hi)))
        seqs.append(s)
    return seqs


def test_indexed_batch_matches_plain_batch(route):
    rng = np.random.default_rng(31)
    pool1 = _random_proteins(rng, 40, 30, 200)
    pool2 = _random_proteins(rng, 70, 30, 200)
    idx1 = rng.integers(0, 40, 1500).astype(np.int32)
    idx2 = rng.integers(0, 70, 1500).astype(np.int32)
    for gap, term, local in (((-10, -1), True, False), ((-10, -1), False, False), ((-11, -1), True, True), (-9, True, False)):
        res = bt.align_indexed_batch(pool1, pool2, idx1, idx2, BLOSUM(), gap, term, local)
        plain = bt.align_optimal_batch([pool1[i] for i in idx1], [pool2[j] for j in idx2], BLOSUM(), gap, term, local)
        assert np.array_equal(res.scores, plain.scores)
        a, b = res.traces(), plain.traces()
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        k = 17
        assert res.alignment(k) == plain.alignment(k)
        only = bt.align_indexed_batch(pool1, pool2, idx1, idx2, BLOSUM(), gap, term, local, score_only=True)
        assert np.array_equal(only.scores, plain.scores)


def test_all_vs_all_scores_semi_global():
    """Config 5 shape at test size: every pair i <= j, terminal_penalty=False, vs the oracle."""
    rng = np.random.default_rng(37)
    seqs = _random_proteins(rng, 60, 40, 260)
    matrix = BLOSUM()
    got = bt.all_vs_all_scores(seqs, matrix, (-10, -1), terminal_penalty=False, chunk_pairs=500)
    assert got.shape == (60, 60) and np.array_equal(got, got.T)
    for i in range(0, 60, 7):
        for j in range(i, 60, 5):
            ref = oracle_align_optimal(seqs[i], seqs[j], matrix, (-10, -1), terminal_penalty=False,
                                       score_only=True)
            assert got[i, j] == ref


# ---- next row (f): guide-tree distances, statistics fused into the GPU traceback ----------------

def _count_gaps(trace, terminal_penalty):
    """Restatement of the reference's _count_gaps (multiple.pyx:403-455) on an (L,2) trace."""
    if not terminal_penalty:
        full = np.nonzero((trace != -1).all(axis=1))[0]
        if full.size == 0:
            return 0, 0
        trace = trace[full[0]:full[-1] + 1]
    opens = exts = 0
    for col in (0, 1):
        gap = trace[:, col] == -1
        prev = np.concatenate(([False], gap[:-1]))
        opens += int(np.count_nonzero(gap & ~prev))
        exts += int(np.count_nonzero(gap & prev))
    return opens, exts


@pytest.mark.parametrize("term", [True, False])
@pytest.mark.parametrize("gap", [(-10, -1), -6])
def test_trace_statistics_mode(term, gap, route):
    rng = np.random.default_rng(41)
    codes1, off1, codes2, off2 = _random_batch(rng, 400, 20, 160)
    matrix = BLOSUM().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, False, stats=True) as batch:
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, False)
    assert np.array_equal(res.scores, ref_scores)
    for k in range(400):
        assert res.trace_len[k] == len(ref_traces[k])
        assert tuple(res.gap_counts[k]) == _count_gaps(ref_traces[k], term), k


def test_feng_doolittle_distances():
    rng = np.random.default_rng(43)
    base = rng.integers(0, 20, 180)
    seqs = []
    for k in range(12):     # a family of related sequences, as align_multiple expects
        code = base.copy()
        flip = rng.random(len(code)) < 0.25
        code[flip] = rng.integers(0, 20, int(flip.sum()))
        code = code[rng.random(len(code)) > 0.05]
        s = bt.ProteinSequence(); s.code = code
        seqs.append(s)
    matrix = BLOSUM()
    for gap, term in (((-10, -1), True), (-8, False)):
        got = bt.feng_doolittle_distances(seqs, matrix, gap, term)
        go, ge = gap if isinstance(gap, tuple) else (gap, gap)
        sm = matrix.score_matrix().astype(np.float64)
        n = len(seqs)
        ali = {(i, j): oracle_align_optimal(seqs[i], seqs[j], matrix, gap, term, max_number=1)[0]
               for i in range(n) for j in range(i + 1)}
        ref = np.zeros((n, n), dtype=np.float32)
        for i in range(n):
            for j in range(i):
                ci = np.bincount(seqs[i].code, minlength=24)
                cj = np.bincount(seqs[j].code, minlength=24)
                s_rand = (ci @ sm @ cj) / ali[i, j].trace.shape[0]
                o, e = _count_gaps(ali[i, j].trace, term)
                s_rand += o * go + e * ge
                s_max = (ali[i, i].score + ali[j, j].score) / 2.0
                ref[i, j] = ref[j, i] = -np.log((ali[i, j].score - s_rand) / (s_max - s_rand))
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)   # float32 path, reference tolerance
        assert np.array_equal(got, got.T) and np.all(np.diag(got) == 0)


# ---- next row (f) rank 2: E-value estimation from one batch of random local alignments ------------

BACKGROUND = np.array([35155, 8669, 24161, 28354, 17367, 33229, 9906, 23161, 25872, 40625, 10101, 20212,
                       23435, 19208, 23105, 32070, 26311, 29012, 5990, 14488, 0, 0, 0, 0]) / 450431


@pytest.mark.parametrize("matrix_name,gap,ref_lam,ref_k", [
    ("BLOSUM62", (-10000, -10000), 0.318, 0.130),
    ("BLOSUM62", (-12, -2), 0.300, 0.090),
    ("BLOSUM62", (-5, -5), 0.131, 0.009),
    ("PAM250", (-16, -1), 0.172, 0.018),
])
def test_evalue_distribution_parameters(matrix_name, gap, ref_lam, ref_k):
    """tests/sequence/align/test_statistics.py:46-74: parameters of Altschul et al."""
    alphabet = bt.ProteinSequence.alphabet
    matrix = bt.SubstitutionMatrix(alphabet, alphabet, matrix_name)
    np.random.seed(0)
    est = bt.EValueEstimator.from_samples(alphabet, matrix, gap, BACKGROUND, 500, 1000)
    assert est.lam == pytest.approx(ref_lam, rel=0.1)
    assert est.k == pytest.approx(ref_k, rel=0.6)
    # the sampled scores are bit-exact, so the oracle estimates the very same parameters
    np.random.seed(0)
    code = np.random.choice(24, size=(1000, 2, 500), p=BACKGROUND / BACKGROUND.sum())
    off = np.arange(1001, dtype=np.int64) * 500
    ref_scores, _ = oracle.batch(code[:, 0, :].reshape(-1), off, code[:, 1, :].reshape(-1), off,
                                 matrix.score_matrix(), gap, True, True, score_only=True)
    lam = np.pi / np.sqrt(6 * np.var(ref_scores.astype(np.int64)))
    assert est.lam == pytest.approx(lam, rel=1e-12)


def test_evalue_score_scaling():
    """test_statistics.py:118-177: scaling scores and penalties by 1000 scales lambda by 1/1000."""
    alphabet = bt.ProteinSequence.alphabet
    matrix = BLOSUM()
    scaled = bt.SubstitutionMatrix(alphabet, alphabet, matrix.score_matrix() * 1000)
    np.random.seed(0)
    a = bt.EValueEstimator.from_samples(alphabet, matrix, (-12, -1), BACKGROUND, 300, 400)
    np.random.seed(0)
    b = bt.EValueEstimator.from_samples(alphabet, scaled, (-12000, -1000), BACKGROUND, 300, 400)
    assert b.lam == pytest.approx(a.lam / 1000, rel=1e-9)
    assert b.k == pytest.approx(a.k, rel=1e-6)
    assert a.log_evalue(50, 300, 300 * 10000) == pytest.approx(b.log_evalue(50000, 300, 300 * 10000), rel=1e-6)


# ---- packed kernels: ragged edge cases ---------------------------------------------------------------

@pytest.mark.parametrize("gap", [(-10, -1), -7])
@pytest.mark.parametrize("local,term", [(False, True), (False, False), (True, True)])
def test_packed_kernels_ragged_lengths(gap, local, term, monkeypatch):
    """Lengths 1..40 next to 1900..2000 (the packed route's limit), a pair count that is not a
    multiple of the group size, all 24 protein symbols incl. B, Z, X and *."""
    if isinstance(gap, int) and not local and not term:
        pytest.skip("linear semi-global has no packed kernel")
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(51)
    n_pairs = 203
    len1 = np.where(rng.random(n_pairs) < 0.85, rng.integers(1, 40, n_pairs), rng.integers(1900, 2001, n_pairs))
    len2 = np.where(rng.random(n_pairs) < 0.85, rng.integers(1, 40, n_pairs), rng.integers(1900, 2001, n_pairs))
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, 24, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, 24, off2[-1]).astype(np.uint8)
    for k in range(0, n_pairs, 4):   # some homologous pairs, including long ones
        n = int(min(len1[k], len2[k]))
        codes2[off2[k]:off2[k] + n] = codes1[off1[k]:off1[k] + n]
    matrix = BLOSUM().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local) as batch:
        assert batch.kernel == "interseq_s16x2"
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(n_pairs):
        assert np.array_equal(traces[k], ref_traces[k]), (k, len1[k], len2[k])


def test_packed_kernels_nucleotide_matrix_and_range_guard(monkeypatch):
    """The 15x15 NUC matrix through the packed kernels; a batch whose scores could leave the
    16-bit range must be refused by the packed route even when it is forced."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(53)
    codes1, off1, codes2, off2 = _random_batch(rng, 130, 50, 400, alphabet=15)
    matrix = NUC().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False) as batch:
        assert batch.kernel == "interseq_s16x2"
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False)
    assert np.array_equal(res.scores, ref_scores)
    assert all(np.array_equal(a, b) for a, b in zip(res.traces(), ref_traces))
    big = (matrix * 100).astype(np.int32)   # 400 x 500 per match > 32767
    with bt.DeviceBatch(codes1, off1, codes2, off2, big, (-1000, -100), True, False) as batch:
        assert batch.kernel == "general_i32"
        batch.run()
        res = batch.fetch()
    ref_scores, _ = oracle.batch(codes1, off1, codes2, off2, big, (-1000, -100), True, False, score_only=True)
    assert np.array_equal(res.scores, ref_scores)


# ---- long sequences on the packed route (IS_MAX_LEN = 8000, memory-aware windows) ---------------

def test_packed_route_long_database_sequences(monkeypatch):
    """Config 3 at its real length clip: queries <= 600 against database sequences up to 5000
    residues, local affine score-only, through the packed 16-bit kernels."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(101)
    queries = _random_proteins(rng, 4, 200, 600)
    lengths = [30, 31, 64, 700, 1999, 2001, 3100, 4999, 5000]
    db = [bt.ProteinSequence() for _ in range(24)]
    for k, seq in enumerate(db):
        n = lengths[k % len(lengths)]
        seq.code = rng.integers(0, 20, n).astype(np.uint8)
        # plant a diverged copy of a query so that the optimum is not trivial
        q = queries[k % 4].code
        at = int(rng.integers(0, max(1, n - len(q))))
        piece = q[:max(0, min(len(q), n - at))].copy()
        flip = rng.random(len(piece)) < 0.2
        piece[flip] = rng.integers(0, 20, int(flip.sum()))
        seq.code[at:at + len(piece)] = piece
    idx1 = np.repeat(np.arange(4, dtype=np.int32), len(db))
    idx2 = np.tile(np.arange(len(db), dtype=np.int32), 4)
    res = bt.align_indexed_batch(queries, db, idx1, idx2, BLOSUM(), (-10, -1), local=True, score_only=True)
    for k in range(len(idx1)):
        ref = oracle_align_optimal(queries[idx1[k]], db[idx2[k]], BLOSUM(), (-10, -1), local=True, score_only=True)
        assert res.scores[k] == ref, k


def test_packed_route_long_traced_pairs(monkeypatch):
    """2500-residue pairs, global affine with traces, packed route: 200 MB of direction nibbles
    per group, windows cut by the direction budget."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(55)
    n_pairs = 64
    len1 = rng.integers(2300, 2600, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    codes1 = rng.integers(0, 20, off1[-1]).astype(np.uint8)
    seqs2 = []
    for k in range(n_pairs):
        a = codes1[off1[k]:off1[k + 1]].copy()
        flip = rng.random(len(a)) < 0.3
        a[flip] = rng.integers(0, 20, int(flip.sum()))
        seqs2.append(np.delete(a, rng.choice(len(a), 40, replace=False)))
    off2 = np.concatenate(([0], np.cumsum([len(s) for s in seqs2]))).astype(np.int64)
    codes2 = np.concatenate(seqs2)
    matrix = BLOSUM().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False) as batch:
        assert batch.kernel == "interseq_s16x2"
        batch.run()
        res = batch.fetch()
    traces = res.traces()
    for k in range(0, n_pairs, 11):
        ref_traces, ref_score = oracle.align_optimal(codes1[off1[k]:off1[k + 1]], codes2[off2[k]:off2[k + 1]],
                                                     matrix, (-10, -1), True, False, False, 1)
        assert res.scores[k] == ref_score and np.array_equal(traces[k], ref_traces[0]), k


def test_all_vs_all_with_sequences_beyond_the_16_bit_range():
    """A few sequences long enough that their mutual alignments leave the 16-bit score range
    (shorter sequence > 2976 residues with BLOSUM62): those pairs take the 32-bit kernel, all
    others stay packed; pairs are oriented shorter-first."""
    rng = np.random.default_rng(77)
    seqs = _random_proteins(rng, 70, 40, 200)
    base = np.full(3300, 18, dtype=np.uint8)        # tryptophan: 11 per matched residue
    base[::50] = rng.integers(0, 20, len(base[::50]))
    for n in (3050, 3300):
        seq = bt.ProteinSequence()
        seq.code = base[:n].copy()
        seqs.append(seq)
    got = bt.all_vs_all_scores(seqs, BLOSUM(), (-10, -1), terminal_penalty=False)
    assert np.array_equal(got, got.T)
    checks = [(70, 71), (70, 70), (3, 71), (0, 70), (5, 9)]
    for i, j in checks:
        ref = oracle_align_optimal(seqs[i], seqs[j], BLOSUM(), (-10, -1), terminal_penalty=False, score_only=True)
        assert got[i, j] == ref, (i, j)
    assert got[70, 71] > 32767      # the shared 3050-residue prefix scores beyond int16


def test_match_mismatch_tuples_take_the_packed_route(monkeypatch):
    """A (match, mismatch) pair over a small alphabet is run as the equivalent matrix, so
    nucleotide batches scored with tuples use the packed kernels; results as the reference's
    tuple scoring (scoring.rs:101-108), computed by the oracle with the tuple itself."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")   # the batch is small: skip the cost model
    rng = np.random.default_rng(91)
    seqs1 = [_random_seq(rng, int(rng.integers(60, 300))) for _ in range(192)]
    seqs2 = []
    for s in seqs1:
        code = s.code.copy()
        flip = rng.random(len(code)) < 0.2
        code[flip] = rng.integers(0, 4, int(flip.sum()))
        t = bt.NucleotideSequence()
        t.code = np.delete(code, rng.choice(len(code), 5, replace=False))
        seqs2.append(t)
    for gap, local, term in (((-8, -2), False, True), (-6, True, True), ((-5, -1), False, False)):
        res = bt.align_optimal_batch(seqs1, seqs2, (5, -4), gap, term, local)
        for k in range(0, 192, 13):
            ref = oracle_align_optimal(seqs1[k], seqs2[k], (5, -4), gap, term, local, max_number=1)[0]
            assert res.scores[k] == ref.score and np.array_equal(res.trace(k), ref.trace), (gap, local, term, k)
    p1, p2 = bt.pack_sequences(seqs1), bt.pack_sequences(seqs2)
    from biotite_b200.batch import _resolve_scoring
    with bt.DeviceBatch(p1.codes, p1.offsets, p2.codes, p2.offsets, _resolve_scoring((5, -4), p1, p2), (-8, -2)) as b:
        assert b.kernel == "interseq_s16x2"


def test_search_scores_matches_oracle(monkeypatch):
    """Config 3 through the chunked, pipelined search API: every query against every database
    sequence (local affine, score only), database chunks smaller than the database."""
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")
    rng = np.random.default_rng(19)
    queries = _random_proteins(rng, 8, 60, 300)
    database = _random_proteins(rng, 50, 30, 900)
    for k in range(0, 50, 3):      # plant diverged copies of query fragments
        q = queries[k % 8].code
        piece = q[:min(len(q), len(database[k].code))].copy()
        flip = rng.random(len(piece)) < 0.25
        piece[flip] = rng.integers(0, 20, int(flip.sum()))
        database[k].code[:len(piece)] = piece
    timings = {}
    got = bt.search_scores(queries, database, BLOSUM(), (-10, -1), chunk_size=16, timings=timings)
    assert got.shape == (8, 50) and timings["cells"] == sum(len(q) for q in queries) * sum(len(d) for d in database)
    for i in range(8):
        for j in range(0, 50, 4):
            assert got[i, j] == oracle_align_optimal(queries[i], database[j], BLOSUM(), (-10, -1), local=True,
                                                     score_only=True), (i, j)
    lo, hi = bt.database_shard([len(d) for d in database], 1, 2)
    part = bt.search_scores(queries, database, BLOSUM(), (-10, -1), chunk_size=16, db_range=(lo, hi))
    assert np.array_equal(part, got[:, lo:hi])


def test_batch_with_outlier_pairs_is_split():
    """A couple of pairs beyond the packed kernels' limits (8000 residues / 16-bit scores) in an
    otherwise ordinary batch: they run as a batch of their own, results keep the batch order."""
    rng = np.random.default_rng(123)
    seqs1 = _random_proteins(rng, 200, 50, 250)
    seqs2 = []
    for s in seqs1:
        code = s.code.copy()
        flip = rng.random(len(code)) < 0.25
        code[flip] = rng.integers(0, 20, int(flip.sum()))
        t = bt.ProteinSequence()
        t.code = np.delete(code, rng.choice(len(code), 4, replace=False))
        seqs2.append(t)
    long1, long2 = bt.ProteinSequence(), bt.ProteinSequence()
    long1.code = rng.integers(0, 20, 8300).astype(np.uint8)
    long2.code = np.delete(long1.code, rng.choice(8300, 60, replace=False))
    rich1, rich2 = bt.ProteinSequence(), bt.ProteinSequence()
    rich1.code = np.full(3100, 18, dtype=np.uint8)
    rich2.code = np.full(3050, 18, dtype=np.uint8)
    seqs1[17], seqs2[17] = long1, long2
    seqs1[120], seqs2[120] = rich1, rich2
    res = bt.align_optimal_batch(seqs1, seqs2, BLOSUM(), (-10, -1), terminal_penalty=False)
    assert isinstance(res, bt.CompositeBatchResult) and len(res) == 200
    assert res.scores[120] > 32767
    for k in (0, 16, 17, 18, 119, 120, 121, 199):
        ref = oracle_align_optimal(seqs1[k], seqs2[k], BLOSUM(), (-10, -1), terminal_penalty=False, max_number=1)[0]
        assert res.scores[k] == ref.score and np.array_equal(res.trace(k), ref.trace), k
        assert res.alignment(k) == ref
    traces = res.traces()
    assert all(len(traces[k]) == res.trace_len[k] for k in range(200))
    only = bt.align_optimal_batch(seqs1, seqs2, BLOSUM(), (-10, -1), terminal_penalty=False, score_only=True)
    assert np.array_equal(only.scores, res.scores)


def test_banded_c4_at_the_real_shape(monkeypatch):
    """BASELINE config 4 at its real shape - 10 kb nucleotide pairs, band +-128, NUC, affine (-10,-1),
    score + trace - on 256 pairs against the oracle (banded.rs:164-255, :598-638)."""
    import sys
    monkeypatch.setenv("B200ALIGN_ROUTE", "packed")   # 256 pairs: the cost model would take the intra-pair kernel
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import bench_configs
    n = 256
    codes1, off1, codes2, off2, bands = bench_configs.make_c4(n, 7)
    matrix = NUC().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, bands) as batch:
        assert batch.kernel == "banded_i32"
        batch.run()
        res = batch.fetch()
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, (-10, -1), bands=bands)
    assert np.array_equal(res.scores, ref_scores)
    assert res.scores.min() > 30000          # far beyond 16-bit lanes
    traces = res.traces()
    for k in range(n):
        assert np.array_equal(traces[k], ref_traces[k]), k


@pytest.mark.parametrize("dtype", [np.uint16, np.uint64])
def test_two_and_eight_byte_codes(dtype):
    """code_bytes = 2 and 8 (util.rs:37-52): the general route reads the codes at their width."""
    rng = np.random.default_rng(9)
    from biotite_b200.pairwise import _native_align_pair
    matrix = BLOSUM().score_matrix()
    for local, terminal in ((False, True), (False, False), (True, True)):
        a = rng.integers(0, 20, 150).astype(dtype)
        b = rng.integers(0, 20, 170).astype(dtype)
        traces, score = _native_align_pair(a, b, matrix, (-10, -1), terminal, local, False, 3)
        ref_traces, ref_score = oracle.align_optimal(a, b, matrix, (-10, -1), terminal, local, False, 3)
        assert score == ref_score and len(traces) == len(ref_traces)
        assert all(np.array_equal(x, y) for x, y in zip(traces, ref_traces))
    # a batch of wide codes
    lens1, lens2 = rng.integers(1, 200, 70), rng.integers(1, 200, 70)
    off1 = np.concatenate(([0], np.cumsum(lens1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(lens2))).astype(np.int64)
    c1, c2 = rng.integers(0, 20, off1[-1]).astype(dtype), rng.integers(0, 20, off2[-1]).astype(dtype)
    res = bt.align_batch_arrays(c1, off1, c2, off2, matrix, (-10, -1), True, False)
    ref_scores, ref_tr = oracle.batch(c1, off1, c2, off2, matrix, (-10, -1), True, False)
    assert np.array_equal(res.scores, ref_scores)
    got = res.traces()
    assert all(np.array_equal(got[k], ref_tr[k]) for k in range(70))


def _leading_gap_cases(count=160):
    """Banded pairs where starting with a gap beats starting with a mismatch (gap -3, NUC mismatch -4):
    the second sequence is the first one with its head replaced, the band starts on the border."""
    rng = np.random.default_rng(17)
    seqs1, seqs2, bands = [], [], []
    for _ in range(count):
        n = int(rng.integers(8, 40))
        s1 = _random_seq(rng, n)
        code2 = s1.code.copy()
        head = int(rng.integers(1, 3))
        code2[:head] = (code2[:head] + 1 + rng.integers(0, 3, head)) % 4      # mismatching head
        code2 = np.concatenate((rng.integers(0, 4, int(rng.integers(0, 3))), code2))
        s2 = bt.NucleotideSequence()
        s2.code = code2
        lo = int(rng.integers(-3, 1))
        seqs1.append(s1); seqs2.append(s2); bands.append((lo, lo + int(rng.integers(3, 9))))
    return seqs1, seqs2, bands


def test_banded_leading_gap_is_printed_as_a_pair(route):
    """trace.rs:61-68: the start cell is not part of the path, so the column of a first gap step is
    printed as the pair in front of the entered cell; re-scoring such a trace gives less than the
    reported score.  The drop-in keeps that (tests/test_oracle_bruteforce.py pins it by enumeration)."""
    seqs1, seqs2, bands = _leading_gap_cases()
    for gap in (-3, (-3, -1)):
        res = bt.align_banded_batch(seqs1, seqs2, NUC(), np.array(bands), gap, False)
        odd = 0
        for k in range(len(seqs1)):
            ref = oracle_align_banded(seqs1[k], seqs2[k], NUC(), bands[k], gap_penalty=gap, local=False, max_number=1)[0]
            assert res.scores[k] == ref.score
            assert np.array_equal(res.trace(k), ref.trace)
            odd += bt.score(ref, NUC(), gap) != ref.score
        assert odd >= 5
    for k in range(0, len(seqs1), 16):
        got = bt.align_banded(seqs1[k], seqs2[k], NUC(), bands[k], gap_penalty=(-3, -1), max_number=10)
        _same(got, oracle_align_banded(seqs1[k], seqs2[k], NUC(), bands[k], gap_penalty=(-3, -1), max_number=10))


@pytest.mark.parametrize("local", [True, False])
@pytest.mark.parametrize("term", [True, False])
@pytest.mark.parametrize("gap", [-5, -10])
def test_linear_penalty_equals_affine_with_equal_open_and_extension(local, term, gap):
    """The reference's own property test (tests/sequence/align/test_pairwise.py:122-166): same score,
    same number of alignments and, when all of them were enumerated, the same alignments."""
    for seed in range(4):
        rng = np.random.default_rng(seed)
        s1, s2 = _random_seq(rng, rng.integers(10, 100)), _random_seq(rng, rng.integers(10, 100))
        ref = bt.align_optimal(s1, s2, NUC(), gap, term, local, 1000)
        test = bt.align_optimal(s1, s2, NUC(), (gap, gap), term, local, 1000)
        assert test[0].score == ref[0].score and len(test) == len(ref)
        if len(test) < 1000:
            assert all(any(np.array_equal(t.trace, r.trace) for r in ref) for t in test)


def test_size_independent_score_properties_at_batch_scale():
    """20 000 ragged protein pairs per mode, more than the oracle checks in reasonable time: properties
    that hold for every pair whatever its size.  (a) swapping the sequences keeps the score (symmetric
    matrix; the 3-state recurrence is symmetric in G1 / G2, pairwise.rs:608-681); (b) a linear penalty
    g scores like the affine penalty (g, g) (the reference's test_pairwise.py:122-166); (c) the score of
    a traced run is the score of the score-only run; (d) local >= semi-global >= global, since each
    search space contains the next and the freed gaps cost <= 0."""
    rng = np.random.default_rng(8)
    n_pairs = 20000
    codes1, off1, codes2, off2 = _random_batch(rng, n_pairs, 30, 420)
    # half of the pairs related (mutated copies), so that scores are not all near the noise floor
    for k in range(0, n_pairs, 2):
        n = min(off1[k + 1] - off1[k], off2[k + 1] - off2[k])
        keep = rng.random(n) > 0.15
        codes2[off2[k]:off2[k] + n][keep] = codes1[off1[k]:off1[k] + n][keep]
    matrix = BLOSUM().score_matrix()

    def scores(c1, o1, c2, o2, gap, term, local, score_only=True):
        return bt.align_batch_arrays(c1, o1, c2, o2, matrix, gap, term, local, score_only=score_only).scores

    by_mode = {}
    for term, local in ((True, False), (False, False), (True, True)):
        s = scores(codes1, off1, codes2, off2, (-10, -1), term, local)
        assert np.array_equal(s, scores(codes2, off2, codes1, off1, (-10, -1), term, local)), (term, local)
        assert np.array_equal(s, scores(codes1, off1, codes2, off2, (-10, -1), term, local, score_only=False)), (term, local)
        lin = scores(codes1, off1, codes2, off2, -6, term, local)
        assert np.array_equal(lin, scores(codes1, off1, codes2, off2, (-6, -6), term, local)), (term, local)
        assert np.array_equal(lin, scores(codes2, off2, codes1, off1, -6, term, local)), (term, local)
        by_mode[(term, local)] = s
    assert np.all(by_mode[(True, True)] >= by_mode[(False, False)])
    assert np.all(by_mode[(False, False)] >= by_mode[(True, False)])
    assert np.all(by_mode[(True, True)] >= 0)


def test_sharded_all_vs_all_block_tasks_on_one_rank():
    """all_vs_all_scores_sharded with the real GPU entry points (triangular and cross generated batches)
    in a one-rank gloo group: three blocks give three diagonal and three cross tasks; the assembled
    matrix must be the one-batch triangle's.  (Two ranks on two GPUs: tools/check_sharded_gpu.py.)"""
    import socket

    import torch.distributed as dist

    from biotite_b200.distributed import all_vs_all_scores_sharded
    rng = np.random.default_rng(23)
    pool = _random_proteins(rng, 240, 20, 400)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        timings = {}
        got = all_vs_all_scores_sharded(pool, BLOSUM(), (-10, -1), terminal_penalty=False, blocks=3, timings=timings)
    finally:
        dist.destroy_process_group()
    ref = bt.all_vs_all_scores(pool, BLOSUM(), (-10, -1), terminal_penalty=False)
    assert timings["tasks"] == 6 and timings["cells"] > 0
    assert np.array_equal(got, ref) and np.array_equal(got, got.T)
