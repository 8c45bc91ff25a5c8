"""
Device-side wire-format emitters (csrc/emit_kernel.cuh, through bt_batch_fetch_cigar /
bt_batch_fetch_gapped) against the host functions that restate the reference
(biotite_b200.cigar.write_alignment_to_cigar, Alignment.get_gapped_sequences) applied to the
oracle's traces.  Bit-exact (strings).
"""
import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle

pytestmark = pytest.mark.gpu

PROTEIN = bt.ProteinSequence.alphabet
NUC = bt.NucleotideSequence.unambiguous_alphabet()


def _sequence(cls_alphabet, code):
    seq = bt.ProteinSequence() if cls_alphabet is PROTEIN else bt.NucleotideSequence(ambiguous=False)
    seq.code = np.asarray(code, dtype=np.uint8)
    return seq


def _batch(rng, n_pairs, lo, hi, n_letters):
    len1, len2 = rng.integers(lo, hi, n_pairs), rng.integers(lo, hi, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, n_letters, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, n_letters, off2[-1]).astype(np.uint8)
    # homologous pairs with substitutions and indels, so that the CIGARs have structure
    for k in range(0, n_pairs, 2):
        src = codes1[off1[k]:off1[k + 1]]
        m = int(len2[k])
        start = int(rng.integers(0, max(1, len(src) // 4)))
        piece = src[start:start + m].copy()
        flip = rng.random(len(piece)) < 0.1
        piece[flip] = rng.integers(0, n_letters, int(flip.sum()))
        if len(piece) > 12:
            cut = int(rng.integers(2, len(piece) - 8))
            piece = np.concatenate((piece[:cut], piece[cut + int(rng.integers(1, 5)):]))
        codes2[off2[k]:off2[k] + len(piece)] = piece
    return codes1, off1, codes2, off2


@pytest.mark.parametrize("route", ["packed", "general"])
@pytest.mark.parametrize("local,term", [(False, True), (False, False), (True, True)])
@pytest.mark.parametrize("gap", [(-10, -1), -8])
def test_cigar_and_gapped_strings_match_host(route, local, term, gap, monkeypatch):
    monkeypatch.setenv("B200ALIGN_ROUTE", route)
    rng = np.random.default_rng(11)
    n_pairs = 192
    codes1, off1, codes2, off2 = _batch(rng, n_pairs, 20, 180, 20)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    _, ref_traces = oracle.batch(codes1, off1, codes2, off2, matrix, gap, term, local)
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, gap, term, local) as batch:
        batch.run()
        variants = {
            (): batch.cigars(),
            ("distinguish",): batch.cigars(distinguish_matches=True),
            ("hard",): batch.cigars(hard_clip=True),
            ("terminal",): batch.cigars(include_terminal_gaps=True),
        }
        arrays = batch.cigars(as_string=False)
        gapped = batch.gapped_sequences(PROTEIN)
    checked = 0
    for k in range(n_pairs):
        s1 = _sequence(PROTEIN, codes1[off1[k]:off1[k + 1]])
        s2 = _sequence(PROTEIN, codes2[off2[k]:off2[k + 1]])
        ali = bt.Alignment([s1, s2], ref_traces[k])
        assert list(gapped[k]) == ali.get_gapped_sequences(), k
        if not np.any(ref_traces[k][:, 1] != -1):
            assert variants[()][k] == ""      # nothing of the segment aligned: the reference raises
            continue
        assert variants[()][k] == bt.write_alignment_to_cigar(ali), k
        assert variants[("distinguish",)][k] == bt.write_alignment_to_cigar(ali, distinguish_matches=True), k
        assert variants[("hard",)][k] == bt.write_alignment_to_cigar(ali, hard_clip=True), k
        assert variants[("terminal",)][k] == bt.write_alignment_to_cigar(ali, include_terminal_gaps=True), k
        assert np.array_equal(arrays[k], bt.write_alignment_to_cigar(ali, as_string=False)), k
        checked += 1
    assert checked > n_pairs // 2


def test_docstring_examples_on_device():
    """cigar.py:272-314 through the batch emitters (pairs replicated to fill a packed group)."""
    ref = bt.NucleotideSequence("TATAAAAGGTTTCCGACCGTAGGTAGCTGA")
    seg = bt.NucleotideSequence("CCCCGGTTTGACCGTATGTAG")
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    res = {}
    for name, kwargs in (("semi", dict(terminal_penalty=False)), ("local", dict(local=True))):
        n = 4
        off1 = np.arange(n + 1, dtype=np.int64) * len(ref)
        off2 = np.arange(n + 1, dtype=np.int64) * len(seg)
        with bt.DeviceBatch(np.tile(ref.code, n), off1, np.tile(seg.code, n), off2, matrix.score_matrix(), -10,
                            kwargs.get("terminal_penalty", True), kwargs.get("local", False)) as batch:
            batch.run()
            res[name] = (batch.cigars(), batch.cigars(distinguish_matches=True),
                         batch.cigars(include_terminal_gaps=True), batch.cigars(hard_clip=True),
                         batch.gapped_sequences(ref.alphabet))
    assert set(res["semi"][0]) == {"9M2D12M"}
    assert set(res["semi"][1]) == {"4X5=2D7=1X4="}
    assert set(res["semi"][2]) == {"3D9M2D12M4D"}
    assert res["semi"][4][0] == ("TATAAAAGGTTTCCGACCGTAGGTAGCTGA", "---CCCCGGTTT--GACCGTATGTAG----")
    assert set(res["local"][0]) == {"4S5M2D12M"}
    assert set(res["local"][3]) == {"4H5M2D12M"}
    assert res["local"][4][0] == ("GGTTTCCGACCGTAGGTAG", "GGTTT--GACCGTATGTAG")
    text = bt.alignments_to_fasta(res["semi"][4][:1], [("ref", "seg")])
    assert text == ">ref\nTATAAAAGGTTTCCGACCGTAGGTAGCTGA\n>seg\n---CCCCGGTTT--GACCGTATGTAG----\n"


def test_banded_batch_emitters():
    """Banded traces (straightened band, swapped pairs) through the emitters."""
    rng = np.random.default_rng(5)
    n_pairs = 48
    codes1, codes2, bands = [], [], []
    for k in range(n_pairs):
        n = int(rng.integers(200, 400))
        a = rng.integers(0, 4, n).astype(np.uint8)
        b = a.copy()
        flip = rng.random(n) < 0.05
        b[flip] = rng.integers(0, 4, int(flip.sum()))
        b = b[rng.random(n) > 0.02]
        codes1.append(a if len(a) <= len(b) else b)     # shorter first, as align_banded arranges it
        codes2.append(b if len(a) <= len(b) else a)
        bands.append((-12, 12))
    off1 = np.concatenate(([0], np.cumsum([len(c) for c in codes1]))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum([len(c) for c in codes2]))).astype(np.int64)
    matrix = bt.SubstitutionMatrix.std_nucleotide_matrix()
    sub = matrix.score_matrix()[:4, :4].copy()
    for local in (False, True):
        with bt.DeviceBatch(np.concatenate(codes1), off1, np.concatenate(codes2), off2, sub, (-10, -1), True, local,
                            bands=np.array(bands)) as batch:
            batch.run()
            result = batch.fetch()
            cigars = batch.cigars(distinguish_matches=True)
            gapped = batch.gapped_sequences(NUC)
        traces = result.traces()
        for k in range(n_pairs):
            s1, s2 = _sequence(NUC, codes1[k]), _sequence(NUC, codes2[k])
            ali = bt.Alignment([s1, s2], traces[k])
            assert list(gapped[k]) == ali.get_gapped_sequences(), k
            assert cigars[k] == bt.write_alignment_to_cigar(ali, distinguish_matches=True), k


def test_emitters_reject_score_only_batches():
    rng = np.random.default_rng(1)
    codes1, off1, codes2, off2 = _batch(rng, 64, 30, 60, 20)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    with bt.DeviceBatch(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, score_only=True) as batch:
        batch.run()
        with pytest.raises(ValueError):
            batch.cigars()
        with pytest.raises(ValueError):
            batch.gapped_sequences(PROTEIN)
