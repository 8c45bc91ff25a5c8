"""Host-side pieces of the batch layer that need no GPU: gathering sub-batches, the 16-bit
score bound used to split batches, and the stitching of part results (CompositeBatchResult).
The part results are built from the CPU oracle in the library's compact trace format and expanded
by the library's host function bt_expand_traces."""
import numpy as np

import biotite_b200 as bt
from biotite_b200.batch import PackedSequences, _packed_shorter_bound, _take
from tests.test_distributed_cpu import _make_batch, _oracle_compute


def test_take_gathers_sequences():
    rng = np.random.default_rng(0)
    lengths = rng.integers(1, 30, 50)
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    codes = rng.integers(0, 20, off[-1]).astype(np.uint8)
    packed = PackedSequences(codes, off, bt.ProteinSequence.alphabet)
    ids = np.array([3, 7, 8, 20, 49, 0])
    sub = _take(packed, ids)
    assert len(sub) == len(ids) and sub.alphabet is packed.alphabet
    for k, i in enumerate(ids):
        assert np.array_equal(sub.codes[sub.offsets[k]:sub.offsets[k + 1]], codes[off[i]:off[i + 1]])
    assert len(_take(packed, np.array([], dtype=np.int64))) == 0


def test_packed_shorter_bound():
    blosum = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    # highest = shorter * 11 + 10 + 11 must stay below 32767 (interseq_eligible)
    bound = _packed_shorter_bound(blosum, (-10, -1))
    assert bound * 11 + 10 + 11 < 32767 <= (bound + 1) * 11 + 10 + 11 + 11
    assert _packed_shorter_bound((5, -4), -6) == 0      # tuples are not bounded here
    nuc = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()
    assert _packed_shorter_bound(nuc, -7) > bound       # smaller best score, longer sequences fit


def test_composite_result_keeps_batch_order():
    codes1, off1, codes2, off2 = _make_batch(seed=3, n_pairs=40)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    p1 = PackedSequences(codes1, off1, bt.ProteinSequence.alphabet)
    p2 = PackedSequences(codes2, off2, bt.ProteinSequence.alphabet)
    whole = _oracle_compute(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, None, False)
    mask = np.zeros(40, dtype=bool)
    mask[[2, 11, 12, 39]] = True
    parts, index = [], []
    for m in (~mask, mask):
        ids = np.nonzero(m)[0]
        q1, q2 = _take(p1, ids), _take(p2, ids)
        parts.append(_oracle_compute(q1.codes, q1.offsets, q2.codes, q2.offsets, matrix, (-10, -1), True, False,
                                     None, False))
        index.append(ids)
    res = bt.CompositeBatchResult(parts, index)
    assert len(res) == 40 and np.array_equal(res.scores, whole.scores)
    assert np.array_equal(res.trace_len, whole.trace_len)
    ref = whole.traces()
    got = res.traces()
    for k in range(40):
        assert np.array_equal(got[k], ref[k]) and np.array_equal(res.trace(k), ref[k]), k


def test_compact_steps_expand_like_step_bytes():
    """bt_expand_traces2 (2-bit steps, int32 lengths) against bt_expand_traces (step bytes) on the
    oracle's traces: the host half of the compact result form needs no GPU."""
    from biotite_b200.batch import BatchResult
    codes1, off1, codes2, off2 = _make_batch(seed=5, n_pairs=60)
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    whole = _oracle_compute(codes1, off1, codes2, off2, matrix, (-10, -1), False, False, None, False)
    n = len(whole.scores)
    ops_off = whole.ops_off
    words_per_pair = (whole.trace_len + 15) // 16
    ops2_off = np.concatenate(([0], np.cumsum(words_per_pair)[:-1])).astype(np.int64) + 3   # any layout will do
    ops2 = np.full(int(ops2_off[-1] + words_per_pair[-1]) + 2, 0xFFFFFFFF, dtype=np.uint32)
    for k in range(n):
        steps = whole.ops[ops_off[k]:ops_off[k] + whole.trace_len[k]].astype(np.uint64)
        for w in range(int(words_per_pair[k])):
            chunk = steps[16 * w:16 * w + 16]
            ops2[ops2_off[k] + w] = np.uint32(int(sum(int(v) << (2 * b) for b, v in enumerate(chunk))))
    compact = BatchResult(whole.scores, whole.trace_len.astype(np.int32), whole.first, None, ops2_off)
    compact.ops2 = ops2
    ref = whole.traces()
    got = compact.traces()
    for k in range(n):
        assert np.array_equal(got[k], ref[k]) and np.array_equal(compact.trace(k), ref[k]), k
    # capacity bound of the one-shot call: depends on the lengths only and always suffices
    from biotite_b200 import _cabi
    lib = _cabi.load()
    cap = lib.bt_ops2_capacity(_cabi.ptr(off1), _cabi.ptr(off2), n)
    assert cap >= int(((np.diff(off1) + np.diff(off2) + 15) // 16).sum())
