"""Shared helpers of the test-suite: golden fixtures and oracle adapters."""

import gzip
import json
from functools import lru_cache
from pathlib import Path

import numpy as np

import biotite_b200 as bt
from oracle import oracle

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"


@lru_cache(maxsize=1)
def golden():
    with gzip.open(GOLDEN_DIR / "legacy_alignments.json.gz", "rt") as f:
        return json.load(f)


@lru_cache(maxsize=1)
def cas9_sequences():
    return [bt.ProteinSequence(e["sequence"]) for e in golden()["cas9"]]


def gap_param(value):
    """JSON lists -> tuples (affine), ints stay ints (linear)."""
    return tuple(value) if isinstance(value, list) else value


def scoring_of(matrix):
    return matrix if isinstance(matrix, tuple) else matrix.score_matrix()


def oracle_align_optimal(seq1, seq2, matrix, gap_penalty=-10, terminal_penalty=True,
                         local=False, max_number=1000, score_only=False):
    """The reference's ``align_optimal`` contract, computed by the CPU oracle."""
    dtype = np.promote_types(seq1.code.dtype, seq2.code.dtype)
    traces, score = oracle.align_optimal(
        np.ascontiguousarray(seq1.code, dtype), np.ascontiguousarray(seq2.code, dtype),
        scoring_of(matrix), gap_penalty, terminal_penalty, local, score_only, max_number)
    if score_only:
        return score
    return [bt.Alignment([seq1, seq2], t, score) for t in traces]


def oracle_align_banded(seq1, seq2, matrix, band, gap_penalty=-10, local=False,
                        max_number=1000, score_only=False):
    traces, score = oracle.banded_wrapper(
        seq1.code, seq2.code, scoring_of(matrix), band, gap_penalty, local, score_only,
        max_number)
    if score_only:
        return score
    return [bt.Alignment([seq1, seq2], t, score) for t in traces]
