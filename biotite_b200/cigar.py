"""
CIGAR strings for alignments: the interface of the reference's
``biotite.sequence.align.cigar`` (``src/biotite/sequence/align/cigar.py``):
:class:`CigarOp`, :func:`write_alignment_to_cigar`, :func:`read_alignment_from_cigar`.

For whole batches the operations are produced on the GPU from the compact traces
(``DeviceBatch.cigars`` -> ``bt_batch_fetch_cigar``, ``csrc/emit_kernel.cuh``); the functions
here are the single-alignment host counterparts with the reference's argument meaning and
error behaviour, plus the decoder of the BAM-encoded device output.
"""

from __future__ import annotations

import enum

import numpy as np

from .alignment import Alignment, get_codes

__all__ = ["CigarOp", "read_alignment_from_cigar", "write_alignment_to_cigar",
           "cigar_strings", "cigar_arrays"]


class CigarOp(enum.IntEnum):
    """CIGAR operations with their BAM numbers (cigar.py:20-35)."""

    MATCH = 0
    INSERTION = 1
    DELETION = 2
    INTRON = 3
    SOFT_CLIP = 4
    HARD_CLIP = 5
    PADDING = 6
    EQUAL = 7
    DIFFERENT = 8
    BACK = 9

    @staticmethod
    def from_cigar_symbol(symbol):
        return _STR_TO_OP[symbol]

    def to_cigar_symbol(self):
        return _OP_TO_STR[self]


_SYMBOLS = "MIDNSHP=XB"
_STR_TO_OP = {c: CigarOp(k) for k, c in enumerate(_SYMBOLS)}
_OP_TO_STR = {v: k for k, v in _STR_TO_OP.items()}
_SYMBOL_BYTES = np.frombuffer(_SYMBOLS.encode(), dtype=np.uint8)


def _parse(cigar):
    ops, count = [], ""
    for char in cigar:
        if char.isdigit():
            count += char
        else:
            ops.append((CigarOp.from_cigar_symbol(char), int(count)))
            count = ""
    return np.array(ops, dtype=int).reshape(-1, 2)


def _format(op_tuples):
    return "".join(f"{int(n)}{_SYMBOLS[int(op)]}" for op, n in op_tuples)


def read_alignment_from_cigar(cigar, position, reference_sequence, segment_sequence):
    """
    Build an :class:`Alignment` from a CIGAR string or ``(operation, length)`` pairs
    (cigar.py:73-193).  `position` is the 0-based position of the first aligned reference
    symbol; soft-clipped segment symbols are skipped, hard-clipped ones are not part of
    `segment_sequence`.
    """
    if isinstance(cigar, str):
        operations = _parse(cigar)
    else:
        operations = np.asarray(cigar, dtype=int)
        if operations.ndim != 2:
            raise ValueError("Expected array with shape (n,2)")
        if operations.shape[1] != 2:
            raise ValueError("Expected (operation, length) pairs")
    sequences = [reference_sequence, segment_sequence]
    if len(operations) == 0:
        return Alignment(sequences, np.zeros((0, 2), dtype=int))
    rows = []
    ref_pos, seg_pos = position, 0
    for op, length in operations:
        op, length = CigarOp(op), int(length)
        if op in (CigarOp.MATCH, CigarOp.EQUAL, CigarOp.DIFFERENT):
            rows.append(np.stack((np.arange(ref_pos, ref_pos + length),
                                  np.arange(seg_pos, seg_pos + length)), axis=-1))
            ref_pos += length
            seg_pos += length
        elif op == CigarOp.INSERTION:
            rows.append(np.stack((np.full(length, -1), np.arange(seg_pos, seg_pos + length)), axis=-1))
            seg_pos += length
        elif op in (CigarOp.DELETION, CigarOp.INTRON):
            rows.append(np.stack((np.arange(ref_pos, ref_pos + length), np.full(length, -1)), axis=-1))
            ref_pos += length
        elif op == CigarOp.SOFT_CLIP:
            seg_pos += length
        elif op == CigarOp.HARD_CLIP:
            pass
        else:
            raise ValueError(f"CIGAR operation {op} is not implemented")
    trace = np.concatenate(rows) if rows else np.zeros((0, 2), dtype=int)
    return Alignment(sequences, trace.astype(np.int64))


def write_alignment_to_cigar(alignment, reference_index=0, segment_index=1, introns=(),
                             distinguish_matches=False, hard_clip=False,
                             include_terminal_gaps=False, as_string=True):
    """
    CIGAR of one :class:`Alignment` (cigar.py:196-392): terminal gaps of the segment are
    dropped unless `include_terminal_gaps`, segment symbols outside the alignment become
    soft or hard clips, reference gaps inside `introns` become ``N``.
    """
    trace = np.asarray(alignment.trace)
    seg_all = trace[:, segment_index]
    aligned = np.nonzero(seg_all != -1)[0]
    # clipped symbols: what precedes the first / follows the last aligned segment symbol
    start_clip = int(seg_all[aligned[0]])
    end_clip = len(alignment.sequences[segment_index]) - int(seg_all[aligned[-1]]) - 1
    if not include_terminal_gaps:
        trace = trace[aligned[0]:aligned[-1] + 1]
    ref, seg = trace[:, reference_index], trace[:, segment_index]
    operations = np.full(len(trace), CigarOp.MATCH, dtype=int)
    insertion, deletion = ref == -1, seg == -1
    if np.any(insertion & deletion):
        raise ValueError("Alignment contains insertion and deletion at the same position")
    operations[insertion] = CigarOp.INSERTION
    operations[deletion] = CigarOp.DELETION
    if introns is not None:
        intron_mask = np.zeros(len(trace), dtype=bool)
        for start, stop in introns:
            if start >= stop:
                raise ValueError("Intron start must be smaller than intron stop")
            if start < 0:
                raise ValueError("Intron start must not be negative")
            intron_mask[(ref >= start) & (ref < stop)] = True
        if np.any(intron_mask & ~deletion):
            raise ValueError("Introns must be within gaps in the reference sequence")
        operations[intron_mask] = CigarOp.INTRON
    if distinguish_matches:
        sub = Alignment(alignment.sequences, trace)
        codes = get_codes(sub)
        equal = codes[reference_index] == codes[segment_index]
        match = operations == CigarOp.MATCH
        operations[equal & match] = CigarOp.EQUAL
        operations[~equal & match] = CigarOp.DIFFERENT
    # run-length encode
    if len(operations):
        starts = np.concatenate(([0], np.nonzero(operations[:-1] != operations[1:])[0] + 1))
        lengths = np.diff(np.append(starts, len(operations)))
        op_tuples = np.stack((operations[starts], lengths), axis=-1)
    else:
        op_tuples = np.zeros((0, 2), dtype=int)
    clip_op = CigarOp.HARD_CLIP if hard_clip else CigarOp.SOFT_CLIP
    parts = []
    if start_clip != 0:
        parts.append(np.array([(clip_op, start_clip)], dtype=int))
    parts.append(op_tuples)
    if end_clip != 0:
        parts.append(np.array([(clip_op, end_clip)], dtype=int))
    op_tuples = np.concatenate(parts)
    return _format(op_tuples) if as_string else op_tuples


# ---- decoding the device output (bt_batch_fetch_cigar) ---------------------------------------
def cigar_arrays(n_ops, cigar, ops_off):
    """Per pair ``(n, 2)`` arrays ``(operation, length)`` from BAM-encoded words."""
    out = []
    for k in range(len(n_ops)):
        words = cigar[ops_off[k]:ops_off[k] + n_ops[k]]
        out.append(np.stack((words & 0xF, words >> 4), axis=-1).astype(int))
    return out


def cigar_strings(n_ops, cigar, ops_off):
    """Per pair CIGAR strings from BAM-encoded words."""
    out = []
    for k in range(len(n_ops)):
        words = cigar[ops_off[k]:ops_off[k] + n_ops[k]]
        out.append("".join(f"{int(w) >> 4}{_SYMBOLS[int(w) & 0xF]}" for w in words))
    return out
