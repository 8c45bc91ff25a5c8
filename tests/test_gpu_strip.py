"""The lean intra-pair kernel (csrc/strip_kernel.cuh: affine, full table, first-optimal direction
bytes) against the oracle and against the full-tie-set wavefront kernel it replaces for
max_number = 1 / score-only work (B200ALIGN_GENERAL=full keeps the old kernel).  Reference:
pairwise.rs:608-681 (cell), :845-881 (optimum), trace.rs:59-90 (trace of index 0)."""
import os

import numpy as np
import pytest

import biotite_b200 as bt
from oracle import oracle

pytestmark = pytest.mark.gpu

MATRIX = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()


def _batch(rng, lens1, lens2, alphabet=20):
    off1 = np.concatenate(([0], np.cumsum(lens1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(lens2))).astype(np.int64)
    return (rng.integers(0, alphabet, off1[-1]).astype(np.uint8), off1,
            rng.integers(0, alphabet, off2[-1]).astype(np.uint8), off2)


def _run(codes1, off1, codes2, off2, gap, terminal, local, score_only=False, full=False):
    os.environ["B200ALIGN_ROUTE"] = "general"
    if full:
        os.environ["B200ALIGN_GENERAL"] = "full"
    try:
        with bt.DeviceBatch(codes1, off1, codes2, off2, MATRIX, gap, terminal, local, score_only=score_only) as batch:
            batch.run()
            assert batch.kernel == "general_i32"
            return batch.fetch()
    finally:
        os.environ.pop("B200ALIGN_ROUTE", None)
        os.environ.pop("B200ALIGN_GENERAL", None)


@pytest.mark.parametrize("gap", [(-10, -1), (-1, -1), (0, 0), (-3, 0), (-12, -3)])
@pytest.mark.parametrize("local,terminal", [(False, True), (False, False), (True, True)])
def test_lean_kernel_matches_the_oracle(gap, local, terminal):
    rng = np.random.default_rng(abs(hash((gap, local, terminal))) % 2**32)
    # one lane, one stripe, stripe seams (128-row stripes), several rounds of warps, ragged widths
    lens1 = np.array([1, 2, 3, 4, 5, 31, 127, 128, 129, 255, 256, 257, 300, 640, 1368, 2100, 17, 64])
    lens2 = np.array([1, 5, 2, 9, 4, 33, 140, 128, 3, 256, 255, 700, 301, 100, 1345, 90, 2300, 63])
    # half of the second sequences are mutated copies: long diagonal runs, ties between gap placements
    codes1, off1, codes2, off2 = _batch(rng, lens1, lens2, alphabet=6)
    for k in range(0, len(lens1), 2):
        n = min(lens1[k], lens2[k])
        codes2[off2[k]:off2[k] + n] = codes1[off1[k]:off1[k] + n]
    res = _run(codes1, off1, codes2, off2, gap, terminal, local)
    ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, MATRIX, gap, terminal, local)
    assert np.array_equal(res.scores, ref_scores)
    traces = res.traces()
    for k in range(len(lens1)):
        assert np.array_equal(traces[k], ref_traces[k]), f"trace of pair {k} ({lens1[k]} x {lens2[k]})"
    only = _run(codes1, off1, codes2, off2, gap, terminal, local, score_only=True)
    assert np.array_equal(only.scores, ref_scores)


def test_lean_and_full_kernels_agree_on_random_pairs():
    rng = np.random.default_rng(11)
    lens1, lens2 = rng.integers(1, 700, 200), rng.integers(1, 700, 200)
    codes1, off1, codes2, off2 = _batch(rng, lens1, lens2)
    for local, terminal in ((False, True), (False, False), (True, True)):
        lean = _run(codes1, off1, codes2, off2, (-10, -1), terminal, local)
        full = _run(codes1, off1, codes2, off2, (-10, -1), terminal, local, full=True)
        assert np.array_equal(lean.scores, full.scores) and np.array_equal(lean.trace_len, full.trace_len)
        assert np.array_equal(lean.first, full.first)
        a, _ = lean.traces_flat()
        b, _ = full.traces_flat()
        assert np.array_equal(a, b)


def test_empty_sequences_and_single_pair_api():
    a, b = bt.ProteinSequence("MKTAYIAKQR"), bt.ProteinSequence("")
    matrix = bt.SubstitutionMatrix.std_protein_matrix()
    for x, y in ((a, b), (b, a), (b, b)):
        got = bt.align_optimal(x, y, matrix, (-10, -1), max_number=1)
        traces, score = oracle.align_optimal(x.code, y.code, MATRIX, (-10, -1), True, False, False, 1)
        assert got[0].score == score and np.array_equal(got[0].trace, traces[0])
    rng = np.random.default_rng(3)
    s1, s2 = bt.ProteinSequence(), bt.ProteinSequence()
    s1.code = rng.integers(0, 20, 1368).astype(np.uint8)
    s2.code = np.delete(s1.code, rng.choice(1368, 23, replace=False))
    for local, terminal in ((False, True), (False, False), (True, True)):
        got = bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=terminal, local=local, max_number=1)
        traces, score = oracle.align_optimal(s1.code, s2.code, MATRIX, (-10, -1), terminal, local, False, 1)
        assert got[0].score == score and np.array_equal(got[0].trace, traces[0])
        only = bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=terminal, local=local, score_only=True)
        assert only == score
        # co-optimal enumeration keeps the full tie sets
        many = bt.align_optimal(s1, s2, matrix, (-10, -1), terminal_penalty=terminal, local=local, max_number=5)
        ref_many, _ = oracle.align_optimal(s1.code, s2.code, MATRIX, (-10, -1), terminal, local, False, 5)
        assert len(many) == len(ref_many) and all(np.array_equal(x.trace, t) for x, t in zip(many, ref_many))


def test_shared_memory_just_below_the_opt_in_threshold():
    """Boundary buffers of 48 KiB minus a little: dynamic + static shared memory cross the default
    limit although the dynamic part alone does not (found by the fuzzer: linear gaps, 3000 ragged pairs)."""
    rng = np.random.default_rng(1)
    nuc = bt.SubstitutionMatrix.std_nucleotide_matrix().score_matrix()      # 225 matrix entries
    for max_m in range(1190, 1200):          # 10 warps x (max_m + 8) boundary entries + 225: 49 000 .. 49 400 bytes
        lens1 = np.array([1199, 30, 700]); lens2 = np.array([max_m, 40, 20])
        codes1, off1, codes2, off2 = _batch(rng, lens1, lens2, alphabet=4)
        os.environ["B200ALIGN_ROUTE"] = "general"
        try:
            with bt.DeviceBatch(codes1, off1, codes2, off2, nuc, -9, True, False) as batch:
                batch.run()
                res = batch.fetch()
        finally:
            os.environ.pop("B200ALIGN_ROUTE", None)
        ref_scores, ref_traces = oracle.batch(codes1, off1, codes2, off2, nuc, -9, True, False)
        assert np.array_equal(res.scores, ref_scores)
        assert all(np.array_equal(res.trace(k), ref_traces[k]) for k in range(3))
