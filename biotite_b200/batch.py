"""
Batched many-pairs entry points (``max_number=1`` semantics, what every in-tree
caller of the reference uses: ``multiple.pyx:328``, ``statistics.py:208``,
``superimpose.py:547``, ``tm.py:374``).

``align_optimal_batch`` / ``align_banded_batch`` take many independent pairs,
run them in one GPU batch through ``bt_batch_*`` of ``libb200align.so`` and
return a :class:`BatchResult` that keeps the traces in the compact form the
GPU produced; ``Alignment`` objects are built lazily because a million Python
objects cost more than the DP itself.
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _cabi
from .alignment import Alignment
from .banded import crop_band
from .matrix import SubstitutionMatrix
from .pairwise import _check_gap_penalty

__all__ = [
    "PackedSequences",
    "pack_sequences",
    "DeviceBatch",
    "BatchResult",
    "align_batch_arrays",
    "compact_out",
    "align_batch_multi",
    "align_indexed_batch",
    "all_vs_all_scores",
    "feng_doolittle_distances",
    "align_optimal_batch",
    "align_banded_batch",
]


class PackedSequences:
    """
    Many sequences in one contiguous code array: sequence ``k`` is
    ``codes[offsets[k]:offsets[k+1]]``.  This is the layout the batch kernels
    read, so packing once and reusing it avoids per-call Python work.
    """

    def __init__(self, codes, offsets, alphabet=None, sequences=None):
        self.codes = np.ascontiguousarray(codes)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.alphabet = alphabet
        self.sequences = sequences

    def __len__(self):
        return len(self.offsets) - 1

    @property
    def lengths(self):
        return np.diff(self.offsets)


def pack_sequences(sequences, dtype=None):
    """Concatenate the codes of `sequences` (an iterable of ``Sequence``)."""
    if isinstance(sequences, PackedSequences):
        return sequences
    sequences = list(sequences)
    if dtype is None:
        dtype = np.uint8
        for s in sequences:
            dtype = np.promote_types(dtype, s.code.dtype)
    lengths = np.fromiter((len(s) for s in sequences), dtype=np.int64, count=len(sequences))
    offsets = np.zeros(len(sequences) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    codes = np.empty(int(offsets[-1]), dtype=dtype)
    for s, a, b in zip(sequences, offsets[:-1], offsets[1:]):
        codes[a:b] = s.code
    alphabet = sequences[0].get_alphabet() if sequences else None
    return PackedSequences(codes, offsets, alphabet, sequences)


def _as_index(idx):
    """Pair index array as contiguous int32; values that do not fit raise instead of wrapping."""
    idx = np.asarray(idx)
    if idx.dtype != np.int32:
        if idx.size and (idx.min() < 0 or idx.max() > np.iinfo(np.int32).max):
            raise ValueError("pair index out of range")
        idx = idx.astype(np.int32)
    return np.ascontiguousarray(idx)


class DeviceBatch:
    """
    Thin handle on a ``BtBatch``: inputs are copied to the GPU on construction
    and stay resident; :meth:`run` launches the fill and traceback kernels and
    returns their device time; :meth:`fetch` copies the results back.
    """

    def __init__(self, codes1, off1, codes2, off2, scoring, gap_penalty,
                 terminal_penalty=True, local=False, bands=None, score_only=False,
                 idx1=None, idx2=None, stats=False):
        """
        Without `idx1` / `idx2` pair ``k`` aligns sequence ``k`` of either side.  With them
        the two sides are sequence *pools* and pair ``k`` aligns ``pool1[idx1[k]]`` with
        ``pool2[idx2[k]]`` (one-vs-many searches, all-vs-all matrices).
        """
        self._lib = _cabi.load()
        self._handle = None
        dtype = np.promote_types(codes1.dtype, codes2.dtype)
        self.codes1 = np.ascontiguousarray(codes1, dtype)
        self.codes2 = np.ascontiguousarray(codes2, dtype)
        self.off1 = np.ascontiguousarray(off1, dtype=np.int64)
        self.off2 = np.ascontiguousarray(off2, dtype=np.int64)
        self.idx1 = None if idx1 is None else _as_index(idx1)
        self.idx2 = None if idx2 is None else _as_index(idx2)
        if self.idx1 is None:
            if len(self.off1) != len(self.off2):
                raise ValueError("Both sides of the batch must hold the same number of sequences")
            self.n_pairs = len(self.off1) - 1
        else:
            if self.idx2 is None or len(self.idx1) != len(self.idx2):
                raise ValueError("idx1 and idx2 must have the same length")
            self.n_pairs = len(self.idx1)
        self.bands = None if bands is None else np.ascontiguousarray(bands, dtype=np.int64)
        if self.bands is not None and self.bands.shape != (self.n_pairs, 2):
            raise ValueError("bands must have shape (n_pairs, 2)")
        self.score_only = bool(score_only)
        self.stats = bool(stats) and not self.score_only
        mode = 1 if self.score_only else (2 if self.stats else 0)
        self._params, self._matrix = _cabi.make_params(
            scoring, gap_penalty, terminal_penalty, local, bands is not None, dtype)
        handle = ctypes.c_void_p(None)
        if self.idx1 is None:
            _cabi.check(self._lib.bt_batch_create(
                ctypes.byref(self._params), _cabi.ptr(self.codes1), _cabi.ptr(self.off1),
                _cabi.ptr(self.codes2), _cabi.ptr(self.off2), self.n_pairs,
                _cabi.ptr(self._matrix), _cabi.ptr(self.bands), mode,
                ctypes.byref(handle)))
        else:
            _cabi.check(self._lib.bt_batch_create_indexed(
                ctypes.byref(self._params), _cabi.ptr(self.codes1), _cabi.ptr(self.off1),
                len(self.off1) - 1, _cabi.ptr(self.codes2), _cabi.ptr(self.off2),
                len(self.off2) - 1, _cabi.ptr(self.idx1), _cabi.ptr(self.idx2), self.n_pairs,
                _cabi.ptr(self._matrix), _cabi.ptr(self.bands), mode,
                ctypes.byref(handle)))
        self._handle = handle

    @classmethod
    def generated(cls, codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True, local=False,
                  kind="cross"):
        """
        Score-only batch whose pairs are enumerated, planned and packed on the GPU
        (``bt_batch_create_generated``): ``kind="cross"`` aligns every sequence of pool 1 with
        every sequence of pool 2 (pair ``i * n2 + j``), ``kind="triangle"`` every pair ``i <= j``
        of one pool (pass it twice; symmetric matrix).  Raises ``ValueError`` for modes outside
        the packed kernels; callers then fall back to index arrays.
        """
        self = cls.__new__(cls)
        self._lib = _cabi.load()
        self._handle = None
        dtype = np.promote_types(codes1.dtype, codes2.dtype)
        self.codes1 = np.ascontiguousarray(codes1, dtype)
        self.codes2 = np.ascontiguousarray(codes2, dtype)
        self.off1 = np.ascontiguousarray(off1, dtype=np.int64)
        self.off2 = np.ascontiguousarray(off2, dtype=np.int64)
        self.idx1 = self.idx2 = self.bands = None
        n1, n2 = len(self.off1) - 1, len(self.off2) - 1
        self.kind = {"cross": 0, "triangle": 1}[kind]
        self.n_pairs = n1 * n2 if self.kind == 0 else n1 * (n1 + 1) // 2
        self.score_only, self.stats = True, False
        self._params, self._matrix = _cabi.make_params(scoring, gap_penalty, terminal_penalty, local, False, dtype)
        handle = ctypes.c_void_p(None)
        _cabi.check(self._lib.bt_batch_create_generated(
            ctypes.byref(self._params), _cabi.ptr(self.codes1), _cabi.ptr(self.off1), n1, _cabi.ptr(self.codes2),
            _cabi.ptr(self.off2), n2, self.kind, _cabi.ptr(self._matrix), 1, ctypes.byref(handle)))
        self._handle = handle
        return self

    def top_k(self, k):
        """``(scores, columns)``, each ``(n1, k)`` int32: the `k` best pool-2 hits of every pool-1
        sequence of a ``kind="cross"`` generated batch after :meth:`run`, selected on the GPU and
        sorted here (higher score first, equal scores by ascending index; ``-1`` pads short rows)."""
        n1 = len(self.off1) - 1
        scores = np.empty((n1, k), dtype=np.int32)
        cols = np.empty((n1, k), dtype=np.int32)
        _cabi.check(self._lib.bt_batch_fetch_topk(self._handle, int(k), _cabi.ptr(scores), _cabi.ptr(cols)))
        order = np.argsort(-scores.astype(np.int64), axis=1, kind="stable")   # rows arrive in ascending column order
        return np.take_along_axis(scores, order, axis=1), np.take_along_axis(cols, order, axis=1)

    def score_matrix(self, out=None):
        """Symmetric ``(n, n)`` score matrix of a ``kind="triangle"`` generated batch after :meth:`run`."""
        n = len(self.off1) - 1
        if out is None:
            out = np.empty((n, n), dtype=np.int32)
        _check_out("out", out, np.int32, (n, n))
        _cabi.check(self._lib.bt_batch_fetch_score_matrix(self._handle, _cabi.ptr(out)))
        return out

    def _ops_offsets(self):
        if self.idx1 is None:
            return self.off1 + self.off2
        lens = np.diff(self.off1)[self.idx1] + np.diff(self.off2)[self.idx2]
        return np.concatenate(([0], np.cumsum(lens)[:-1])).astype(np.int64)

    @property
    def cells(self):
        return int(self._lib.bt_batch_cells(self._handle))

    @property
    def launches(self):
        return int(self._lib.bt_batch_launches(self._handle))

    @property
    def windows(self):
        """Fill launches of the last :meth:`run` (one per window)."""
        return int(self._lib.bt_batch_windows(self._handle))

    @property
    def kernel(self):
        return self._lib.bt_batch_kernel(self._handle).decode()

    def run(self):
        """Launch the kernels; returns the CUDA-event time in milliseconds."""
        ms = ctypes.c_float(0.0)
        _cabi.check(self._lib.bt_batch_run(self._handle, ctypes.byref(ms)))
        return float(ms.value)

    def timings(self):
        """``(fill_ms, traceback_ms)`` of the last :meth:`run` (CUDA events)."""
        fill, back = ctypes.c_float(0.0), ctypes.c_float(0.0)
        _cabi.check(self._lib.bt_batch_timings(self._handle, ctypes.byref(fill), ctypes.byref(back)))
        return float(fill.value), float(back.value)

    def fetch(self, out=None):
        """D2H copy of scores (and compact traces); returns a :class:`BatchResult`."""
        n = self.n_pairs
        if out is None:
            scores = np.empty(n, dtype=np.int32)
            trace_len = first = ops = None
            if not self.score_only:
                trace_len = np.empty(n, dtype=np.int64)
                first = np.empty((n, 2), dtype=np.int32)
                if not self.stats:
                    ops = np.empty(int(self._lib.bt_batch_ops_bytes(self._handle)), dtype=np.uint8)
        else:
            scores, trace_len, first, ops = out
            _check_out("scores", scores, np.int32, (n,))
            if not self.score_only:
                _check_out("trace_len", trace_len, np.int64, (n,))
                _check_out("first", first, np.int32, (n, 2))
                if not self.stats:
                    _check_out("ops", ops, np.uint8, (int(self._lib.bt_batch_ops_bytes(self._handle)),))
        _cabi.check(self._lib.bt_batch_fetch_scores(self._handle, _cabi.ptr(scores)))
        if not self.score_only:
            _cabi.check(self._lib.bt_batch_fetch_traces(
                self._handle, _cabi.ptr(trace_len), _cabi.ptr(first), _cabi.ptr(ops)))
        if self.stats:
            # statistics mode: `first` carries (gap opening, gap extending) column counts
            result = BatchResult(scores, trace_len, None, None, np.zeros(0, dtype=np.int64), self.bands)
            result.gap_counts = first
            return result
        # step-byte offsets are only needed to expand traces (10^8 score-only pairs: skip 1.4 s)
        ops_off = np.zeros(0, dtype=np.int64) if self.score_only else self._ops_offsets()
        return BatchResult(scores, trace_len, first, ops, ops_off, self.bands)

    # ---- wire formats emitted on the device (csrc/emit_kernel.cuh) ---------------------------
    def cigars(self, distinguish_matches=False, hard_clip=False, include_terminal_gaps=False,
               as_string=True, swapped=None):
        """
        CIGAR of every pair of a traced batch after :meth:`run`, produced on the GPU from the
        compact traces: the batched ``write_alignment_to_cigar`` (reference
        ``align/cigar.py:196-392``; reference = first, segment = second sequence).  Returns a
        list of strings, or of ``(n, 2)`` ``(operation, length)`` arrays if not `as_string`.
        """
        from . import cigar as _cigar
        n = self.n_pairs
        total = int(self._lib.bt_batch_ops_bytes(self._handle))
        n_ops = np.empty(n, dtype=np.int32)
        words = np.empty(total, dtype=np.uint32)
        flags = (1 if distinguish_matches else 0) | (2 if hard_clip else 0) | (4 if include_terminal_gaps else 0)
        swapped = None if swapped is None else np.ascontiguousarray(swapped, dtype=np.uint8)
        _cabi.check(self._lib.bt_batch_fetch_cigar(self._handle, flags, _cabi.ptr(swapped),
                                                   _cabi.ptr(n_ops), _cabi.ptr(words)))
        off = self._ops_offsets()
        return (_cigar.cigar_strings if as_string else _cigar.cigar_arrays)(n_ops, words, off)

    def gapped_sequences(self, alphabet1, alphabet2=None, swapped=None, as_bytes=False):
        """
        The two gapped strings of every pair (``Alignment.get_gapped_sequences``, reference
        ``align/alignment.py:233-243``), written by the GPU from the compact traces.  The
        alphabets must consist of single characters.  Returns a list of ``(str, str)``.
        """
        alphabet2 = alphabet1 if alphabet2 is None else alphabet2
        tables = []
        for alphabet in (alphabet1, alphabet2):
            symbols = [str(x) for x in alphabet.get_symbols()]
            if any(len(x) != 1 for x in symbols):
                raise ValueError("The alphabet uses symbols other than single characters")
            tables.append(np.frombuffer("".join(symbols).encode("latin-1"), dtype=np.uint8).copy())
        n = self.n_pairs
        total = int(self._lib.bt_batch_ops_bytes(self._handle))
        n_cols = np.empty(n, dtype=np.int64)
        g1 = np.empty(total, dtype=np.uint8)
        g2 = np.empty(total, dtype=np.uint8)
        swapped = None if swapped is None else np.ascontiguousarray(swapped, dtype=np.uint8)
        _cabi.check(self._lib.bt_batch_fetch_gapped(
            self._handle, _cabi.ptr(tables[0]), len(tables[0]), _cabi.ptr(tables[1]), len(tables[1]),
            _cabi.ptr(swapped), _cabi.ptr(n_cols), _cabi.ptr(g1), _cabi.ptr(g2)))
        off = self._ops_offsets()
        if as_bytes:
            return n_cols, g1, g2, off
        b1, b2 = g1.tobytes(), g2.tobytes()
        return [(b1[off[k]:off[k] + n_cols[k]].decode("latin-1"), b2[off[k]:off[k] + n_cols[k]].decode("latin-1"))
                for k in range(n)]

    def close(self):
        if self._handle is not None:
            self._lib.bt_batch_destroy(self._handle)
            self._handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_out(name, array, dtype, shape):
    """Caller-supplied result arrays go to the C ABI as raw pointers: insist on the exact layout."""
    if not isinstance(array, np.ndarray) or array.dtype != np.dtype(dtype) or array.shape != tuple(shape) \
            or not array.flags.c_contiguous or not array.flags.writeable:
        raise ValueError(f"out: `{name}` must be a writeable C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")


def compact_out(off1, off2, pinned=True, n_devices=1):
    """
    Result arrays ``(scores, trace_len, first, ops2, ops2_off)`` for the compact form of
    :func:`align_batch_arrays`, page-locked by default so that the device-to-host copies of a
    window overlap the kernels of the next one.  Reusable across calls of the same shape.
    """
    lib = _cabi.load()
    off1 = np.ascontiguousarray(off1, dtype=np.int64)
    off2 = np.ascontiguousarray(off2, dtype=np.int64)
    n = len(off1) - 1
    cap = int(lib.bt_ops2_capacity(_cabi.ptr(off1), _cabi.ptr(off2), n)) + (16 * int(n_devices) if int(n_devices) > 1 else 0)
    make = _cabi.pinned_empty if pinned else (lambda shape, dtype: np.empty(shape, dtype))
    return (make((n,), np.int32), make((n,), np.int32), make((n, 2), np.int32), make((cap,), np.uint32),
            make((n,), np.int64))


def align_batch_arrays(codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True,
                       local=False, bands=None, score_only=False, out=None, compact=False):
    """
    One call from host arrays to host results (``bt_align_batch``): H2D copies, kernels
    and D2H copies are pipelined inside the library.  `out` may supply preallocated
    (ideally page-locked, see ``_cabi.pinned_empty``; with pageable arrays every copy blocks
    the host until it is done) result arrays ``(scores, trace_len, first, ops)``.

    `compact` selects the compact trace form (``bt_align_batch_compact``): ``trace_len`` as int32
    and the steps as 2 bits each, so that only about ``L / 4`` bytes per pair leave the device;
    `out` is then ``(scores, trace_len, first, ops2, ops2_off)`` as made by :func:`compact_out`.
    Returns a :class:`BatchResult`.
    """
    lib = _cabi.load()
    dtype = np.promote_types(codes1.dtype, codes2.dtype)
    codes1 = np.ascontiguousarray(codes1, dtype)
    codes2 = np.ascontiguousarray(codes2, dtype)
    off1 = np.ascontiguousarray(off1, dtype=np.int64)
    off2 = np.ascontiguousarray(off2, dtype=np.int64)
    if len(off1) != len(off2):
        raise ValueError("Both sides of the batch must hold the same number of sequences")
    n = len(off1) - 1
    bands = None if bands is None else np.ascontiguousarray(bands, dtype=np.int64)
    params, matrix = _cabi.make_params(scoring, gap_penalty, terminal_penalty, local,
                                       bands is not None, dtype)
    if bands is not None and bands.shape != (n, 2):
        raise ValueError("bands must have shape (n_pairs, 2)")
    if compact:
        if out is None:
            out = compact_out(off1, off2, pinned=False)
        scores, trace_len, first, ops2, ops2_off = out
        cap = int(lib.bt_ops2_capacity(_cabi.ptr(off1), _cabi.ptr(off2), n))
        _check_out("scores", scores, np.int32, (n,))
        if not score_only:
            _check_out("trace_len", trace_len, np.int32, (n,))
            _check_out("first", first, np.int32, (n, 2))
            _check_out("ops2", ops2, np.uint32, (cap,))
            _check_out("ops2_off", ops2_off, np.int64, (n,))
        words = ctypes.c_int64(0)
        _cabi.check(lib.bt_align_batch_compact(
            ctypes.byref(params), _cabi.ptr(codes1), _cabi.ptr(off1), _cabi.ptr(codes2),
            _cabi.ptr(off2), n, _cabi.ptr(matrix), _cabi.ptr(bands), int(bool(score_only)),
            _cabi.ptr(scores), None if score_only else _cabi.ptr(trace_len),
            None if score_only else _cabi.ptr(first), None if score_only else _cabi.ptr(ops2), cap,
            None if score_only else _cabi.ptr(ops2_off), ctypes.byref(words)))
        if score_only:
            return BatchResult(scores, None, None, None, np.zeros(0, dtype=np.int64), bands)
        result = BatchResult(scores, trace_len, first, None, ops2_off, bands)
        result.ops2 = ops2
        result.ops2_words = int(words.value)
        return result
    if out is None:
        scores = np.empty(n, dtype=np.int32)
        trace_len = first = ops = None
        if not score_only:
            trace_len = np.empty(n, dtype=np.int64)
            first = np.empty((n, 2), dtype=np.int32)
            ops = np.empty(int(off1[-1] + off2[-1]), dtype=np.uint8)
    else:
        scores, trace_len, first, ops = out
        _check_out("scores", scores, np.int32, (n,))
        if not score_only:
            _check_out("trace_len", trace_len, np.int64, (n,))
            _check_out("first", first, np.int32, (n, 2))
            _check_out("ops", ops, np.uint8, (int(off1[-1] + off2[-1]),))
    _cabi.check(lib.bt_align_batch(
        ctypes.byref(params), _cabi.ptr(codes1), _cabi.ptr(off1), _cabi.ptr(codes2),
        _cabi.ptr(off2), n, _cabi.ptr(matrix), _cabi.ptr(bands), int(bool(score_only)),
        _cabi.ptr(scores), _cabi.ptr(trace_len), _cabi.ptr(first), _cabi.ptr(ops)))
    return BatchResult(scores, trace_len, first, ops, (off1, off2), bands)


def align_batch_multi(codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True, local=False,
                      bands=None, score_only=False, devices=None, out=None):
    """
    :func:`align_batch_arrays` (compact result form) over several GPUs of the node from this one
    process (``bt_align_batch_multi``): the pairs are cut into contiguous shards of equal DP cells,
    one host thread per device, results written straight into disjoint slices of the result arrays.
    `devices`: list of CUDA device ordinals (default: all).  `out`: arrays from
    ``compact_out(off1, off2, n_devices=max(2, len(devices)))``.  Returns a :class:`BatchResult` whose
    ``shards`` attribute holds the first pair of every device's shard.
    """
    lib = _cabi.load()
    dtype = np.promote_types(codes1.dtype, codes2.dtype)
    codes1 = np.ascontiguousarray(codes1, dtype)
    codes2 = np.ascontiguousarray(codes2, dtype)
    off1 = np.ascontiguousarray(off1, dtype=np.int64)
    off2 = np.ascontiguousarray(off2, dtype=np.int64)
    if len(off1) != len(off2):
        raise ValueError("Both sides of the batch must hold the same number of sequences")
    n = len(off1) - 1
    bands = None if bands is None else np.ascontiguousarray(bands, dtype=np.int64)
    if bands is not None and bands.shape != (n, 2):
        raise ValueError("bands must have shape (n_pairs, 2)")
    if devices is None:
        devices = list(range(lib.bt_device_count()))
    devs = np.ascontiguousarray(devices, dtype=np.int32)
    params, matrix = _cabi.make_params(scoring, gap_penalty, terminal_penalty, local, bands is not None, dtype)
    if out is None:
        out = compact_out(off1, off2, pinned=False, n_devices=max(2, len(devs)))
    scores, trace_len, first, ops2, ops2_off = out
    cap = int(lib.bt_ops2_capacity(_cabi.ptr(off1), _cabi.ptr(off2), n)) + 16 * len(devs)
    _check_out("scores", scores, np.int32, (n,))
    if not score_only:
        _check_out("trace_len", trace_len, np.int32, (n,))
        _check_out("first", first, np.int32, (n, 2))
        if ops2.dtype != np.uint32 or ops2.ndim != 1 or len(ops2) < cap or not ops2.flags.c_contiguous:
            raise ValueError(f"out: `ops2` must be a C-contiguous uint32 array of at least {cap} words")
        _check_out("ops2_off", ops2_off, np.int64, (n,))
    words = ctypes.c_int64(0)
    shards = np.zeros(len(devs) + 1, dtype=np.int64)
    _cabi.check(lib.bt_align_batch_multi(
        _cabi.ptr(devs), len(devs), ctypes.byref(params), _cabi.ptr(codes1), _cabi.ptr(off1), _cabi.ptr(codes2),
        _cabi.ptr(off2), n, _cabi.ptr(matrix), _cabi.ptr(bands), int(bool(score_only)), _cabi.ptr(scores),
        None if score_only else _cabi.ptr(trace_len), None if score_only else _cabi.ptr(first),
        None if score_only else _cabi.ptr(ops2), len(ops2) if not score_only else 0,
        None if score_only else _cabi.ptr(ops2_off), ctypes.byref(words), _cabi.ptr(shards)))
    if score_only:
        result = BatchResult(scores, None, None, None, np.zeros(0, dtype=np.int64), bands)
    else:
        result = BatchResult(scores, trace_len, first, None, ops2_off, bands)
        result.ops2 = ops2
        result.ops2_words = int(words.value)
    result.shards = shards
    return result


class BatchResult:
    """
    Scores and first-optimal traces of a batch.

    Attributes
    ----------
    scores : ndarray, int32, shape=(n,)
    trace_len : ndarray, int64, shape=(n,) or None
        Number of alignment columns per pair.
    first : ndarray, int32, shape=(n,2) or None
        DP-table coordinates of the cell of the first alignment column.
    ops : ndarray, uint8 or None
        Step bytes (0: both advance, 1: gap in sequence 1, 2: gap in sequence 2)
        in end-to-start order; pair ``k`` starts at ``ops_off[k]``.
    """

    def __init__(self, scores, trace_len, first, ops, ops_off, bands=None, swapped=None,
                 seqs1=None, seqs2=None):
        self.scores = scores
        self.trace_len = trace_len
        self.first = first
        self.ops = ops
        # offsets of the step bytes; may be given lazily as a (off1, off2) pair of offset arrays
        self._ops_off = ops_off
        self.bands = bands
        self.swapped = swapped
        self.seqs1 = seqs1
        self.seqs2 = seqs2
        self.gap_counts = None   # (n, 2) gap-open / gap-extension column counts (statistics batches)
        # compact form (bt_align_batch_compact): 2-bit steps, pair k starts at word ops_off[k]
        self.ops2 = None
        self.ops2_words = 0

    def __len__(self):
        return len(self.scores)

    @property
    def ops_off(self):
        if isinstance(self._ops_off, tuple):
            self._ops_off = self._ops_off[0] + self._ops_off[1]
        self._ops_off = np.ascontiguousarray(self._ops_off, dtype=np.int64)
        return self._ops_off

    def traces_flat(self):
        """
        All traces in one ``(sum(L), 2)`` int64 array plus the ``n+1`` row offsets.
        """
        if self.trace_len is None:
            raise ValueError("The batch was computed with score_only=True")
        lib = _cabi.load()
        n = len(self.scores)
        rows_off = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(self.trace_len, out=rows_off[1:])
        out = np.empty((int(rows_off[-1]), 2), dtype=np.int64)
        if self.ops2 is not None:
            _cabi.check(lib.bt_expand_traces2(
                n, _cabi.ptr(self.trace_len), _cabi.ptr(np.ascontiguousarray(self.first)),
                _cabi.ptr(self.ops2), _cabi.ptr(self.ops_off), _cabi.ptr(self.bands),
                _cabi.ptr(rows_off), _cabi.ptr(out)))
        else:
            _cabi.check(lib.bt_expand_traces(
                n, _cabi.ptr(self.trace_len), _cabi.ptr(np.ascontiguousarray(self.first)),
                _cabi.ptr(self.ops), _cabi.ptr(self.ops_off), _cabi.ptr(self.bands),
                _cabi.ptr(rows_off), _cabi.ptr(out)))
        if self.swapped is not None and self.swapped.any():
            flip = np.repeat(self.swapped, self.trace_len)
            out[flip] = out[flip][:, ::-1]
        return out, rows_off

    def traces(self):
        """List of ``(L, 2)`` int64 trace arrays (views into one buffer)."""
        flat, rows_off = self.traces_flat()
        return [flat[rows_off[k]:rows_off[k + 1]] for k in range(len(self.scores))]

    def trace(self, k):
        lib = _cabi.load()
        length = np.ascontiguousarray(self.trace_len[k:k + 1])
        rows_off = np.array([0, int(length[0])], dtype=np.int64)
        out = np.empty((int(length[0]), 2), dtype=np.int64)
        band = None if self.bands is None else np.ascontiguousarray(self.bands[k:k + 1])
        expand, steps = (lib.bt_expand_traces, self.ops) if self.ops2 is None else (lib.bt_expand_traces2, self.ops2)
        _cabi.check(expand(
            1, _cabi.ptr(length), _cabi.ptr(np.ascontiguousarray(self.first[k:k + 1])),
            _cabi.ptr(steps), _cabi.ptr(self.ops_off[k:k + 1]), _cabi.ptr(band),
            _cabi.ptr(rows_off), _cabi.ptr(out)))
        if self.swapped is not None and self.swapped[k]:
            out = np.ascontiguousarray(out[:, ::-1])
        return out

    def alignment(self, k, seq1=None, seq2=None):
        """The :class:`Alignment` of pair `k` (sequences default to the batch inputs)."""
        if seq1 is None:
            seq1 = self.seqs1[k]
        if seq2 is None:
            seq2 = self.seqs2[k]
        return Alignment([seq1, seq2], self.trace(k), int(self.scores[k]))

    def alignments(self):
        traces = self.traces()
        return [
            Alignment([self.seqs1[k], self.seqs2[k]], traces[k], int(self.scores[k]))
            for k in range(len(self.scores))
        ]


def align_indexed_batch(pool1, pool2, idx1, idx2, matrix, gap_penalty=-10, terminal_penalty=True,
                        local=False, score_only=False):
    """
    ``align_optimal(pool1[idx1[k]], pool2[idx2[k]], ..., max_number=1)`` for every ``k``
    without materialising the pairs: the pools (lists of ``Sequence`` or
    :class:`PackedSequences`) are uploaded once.  Returns a :class:`BatchResult`.
    """
    _check_gap_penalty(gap_penalty)
    p1, p2 = pack_sequences(pool1), pack_sequences(pool2)
    scoring = _resolve_scoring(matrix, p1, p2)
    with DeviceBatch(p1.codes, p1.offsets, p2.codes, p2.offsets, scoring, gap_penalty,
                     terminal_penalty, local, None, score_only, idx1=idx1, idx2=idx2) as batch:
        batch.run()
        result = batch.fetch()
    if p1.sequences is not None and p2.sequences is not None:
        result.seqs1 = _IndexedView(p1.sequences, batch.idx1)
        result.seqs2 = _IndexedView(p2.sequences, batch.idx2)
    return result


class _IndexedView:
    def __init__(self, items, index):
        self._items, self._index = items, index

    def __getitem__(self, k):
        return self._items[int(self._index[k])]

    def __len__(self):
        return len(self._index)


def all_vs_all_scores(sequences, matrix, gap_penalty=-10, terminal_penalty=True, local=False,
                      chunk_pairs=1 << 24, timings=None, chunk_limit=2**31 - 1):
    """
    Symmetric ``(N, N)`` int32 matrix of optimal alignment scores of every pair ``i <= j``
    (the quantity the reference's ``align_multiple`` computes pair by pair for its guide
    tree, ``multiple.pyx:322-333``).  The triangle is enumerated in chunks of `chunk_pairs`
    index pairs; the sequences are uploaded once per chunk call, never the pairs.  Protein-sized
    jobs (up to `chunk_limit` pairs, 8-bit codes, a matrix the packed kernels can stage) skip all
    of that: the GPU enumerates the triangle itself (``bt_batch_create_generated``).  `timings`
    (a dict) receives the summed CUDA-event time ``device_ms`` and the ``cells`` of the batches.
    `sequences` may be a list of ``Sequence`` or a :class:`PackedSequences`.
    """
    _check_gap_penalty(gap_penalty)
    pool = sequences if hasattr(sequences, "offsets") else pack_sequences(sequences)
    scoring = _resolve_scoring(matrix, pool, pool)
    if not isinstance(scoring, tuple) and not np.array_equal(scoring, scoring.T):
        raise ValueError("all_vs_all_scores needs a symmetric substitution matrix")
    n = len(pool)
    if not isinstance(scoring, tuple) and pool.codes.dtype == np.uint8 and 0 < n and n * (n + 1) // 2 <= chunk_limit:
        # the whole triangle as one batch, enumerated / sorted / packed on the GPU
        try:
            batch = DeviceBatch.generated(pool.codes, pool.offsets, pool.codes, pool.offsets, scoring, gap_penalty,
                                          terminal_penalty, local, kind="triangle")
        except ValueError:
            batch = None   # a mode outside the packed kernels: index arrays below
        if batch is not None:
            with batch:
                ms = batch.run()
                if timings is not None:
                    timings["device_ms"], timings["cells"], timings["kernel"] = ms, batch.cells, batch.kernel
                return batch.score_matrix()
    out = np.zeros((n, n), dtype=np.int32)
    lengths = pool.lengths
    bound = _packed_shorter_bound(scoring, gap_penalty)
    total_ms, total_cells, kernels = 0.0, 0, set()

    def chunks():
        """Index pairs (first, second) of one batch after the other."""
        rows_done = 0
        while rows_done < n:
            # rows i = rows_done .. stop-1 of the upper triangle (j >= i)
            counts = n - np.arange(rows_done, n)
            stop = rows_done + max(1, int(np.searchsorted(np.cumsum(counts), chunk_pairs, side="right")))
            stop = min(stop, n)
            i_idx = np.repeat(np.arange(rows_done, stop, dtype=np.int32), n - np.arange(rows_done, stop))
            j_idx = np.concatenate([np.arange(i, n, dtype=np.int32) for i in range(rows_done, stop)])
            # The packed 16-bit kernels are chosen per batch from its longest first / second
            # sequence: orient every pair shorter-first (the matrix is symmetric, so is the
            # score) and run the pairs whose shorter sequence keeps the scores inside the 16-bit
            # range separately from the few that need the 32-bit kernel.
            swap = lengths[i_idx] > lengths[j_idx]
            first, second = np.where(swap, j_idx, i_idx), np.where(swap, i_idx, j_idx)
            short = (lengths[first] <= bound) & (lengths[second] <= _PACKED_MAX_LEN)
            for mask in (short, ~short):
                if mask.any():
                    yield first[mask], second[mask]
            rows_done = stop

    def prepare(a, b):
        return a, b, DeviceBatch(pool.codes, pool.offsets, pool.codes, pool.offsets, scoring, gap_penalty,
                                 terminal_penalty, local, None, True, idx1=a, idx2=b)

    # the next batch is built (index arrays, sort, plan, uploads) while the GPU scores this one
    import threading
    source = chunks()
    slot = {}

    device = _cabi.load().bt_get_device()

    def prepare_next(done=None):
        try:
            # a new host thread starts on device 0: take the caller's device
            if device >= 0:
                _cabi.check(_cabi.load().bt_set_device(device))
            if done is not None:       # scores of the previous batch -> both triangles
                a, b, scores = done
                out[a, b] = scores
                out[b, a] = scores
            item = next(source, None)
            slot["next"] = None if item is None else prepare(*item)
        except BaseException as exc:   # re-raised on the consuming thread
            slot["next"] = exc

    prepare_next()
    done = None
    while slot["next"] is not None:
        current = slot.pop("next")
        if isinstance(current, BaseException):
            raise current
        a, b, batch = current
        worker = threading.Thread(target=prepare_next, args=(done,))
        worker.start()
        try:
            total_ms += batch.run()
            total_cells += batch.cells
            kernels.add(batch.kernel)
            done = (a, b, batch.fetch().scores)
        finally:
            batch.close()
            worker.join()
    if isinstance(slot.get("next"), BaseException):
        raise slot["next"]
    if done is not None:
        out[done[0], done[1]] = done[2]
        out[done[1], done[0]] = done[2]
    if timings is not None:
        timings["device_ms"] = total_ms
        timings["cells"] = total_cells
        timings["kernel"] = "+".join(sorted(kernels))
    return out


_PACKED_MAX_LEN = 8000   # IS_MAX_LEN of csrc/interseq_kernel.cuh


def _packed_shorter_bound(scoring, gap_penalty):
    """Longest shorter sequence whose scores provably stay below 2^15 (the `highest` bound of
    interseq_eligible, csrc/interseq_kernel.cuh)."""
    if isinstance(scoring, tuple):
        return 0
    best = max(int(np.max(scoring)), 1)
    gap_open = gap_penalty[0] if isinstance(gap_penalty, tuple) else gap_penalty
    return (32766 + int(gap_open) - best) // best


def feng_doolittle_distances(sequences, matrix, gap_penalty=-10, terminal_penalty=True):
    """
    Pairwise distance matrix of the progressive-alignment guide tree (Feng & Doolittle), as
    the reference's ``align_multiple`` computes it (``multiple.pyx:286-400``): every pair
    ``i >= j`` is aligned with ``max_number=1``; with the alignment score ``S``, its length
    ``L`` and its gap-opening / gap-extension column counts,

        S_max  = (S_ii + S_jj) / 2
        S_rand = sum_ab matrix[a, b] * count_i[a] * count_j[b] / L + n_open * open + n_ext * ext
        D_ij   = -log((S_ij - S_rand) / (S_max - S_rand))

    Scores, lengths and gap counts come from one GPU batch in statistics mode, so no trace
    leaves the device.  Returns a symmetric float32 ``(N, N)`` matrix with a zero diagonal;
    raises ``ValueError`` like the reference if a randomised alignment scores better.
    """
    _check_gap_penalty(gap_penalty)
    pool = pack_sequences(sequences)
    scoring = _resolve_scoring(matrix, pool, pool)
    if isinstance(scoring, tuple):
        raise TypeError("A substitution matrix is required")
    n = len(pool)
    i_idx, j_idx = np.tril_indices(n)          # i >= j, including the diagonal
    i_idx, j_idx = i_idx.astype(np.int32), j_idx.astype(np.int32)
    # The orientation (sequence i first, j second, i >= j) is the reference's and is kept: the
    # first-optimal trace, hence the gap counts, may depend on it.  The library routes a batch
    # as a whole by its longest first / second sequence, so the pairs are run in up to three
    # batches: first sequence inside the 16-bit bound, else second inside it, else neither
    # (the 32-bit kernel; a handful of pairs of very long sequences).
    lengths = pool.lengths
    bound = _packed_shorter_bound(scoring, gap_penalty)
    len_i, len_j = lengths[i_idx], lengths[j_idx]
    fits = (len_i <= _PACKED_MAX_LEN) & (len_j <= _PACKED_MAX_LEN)
    first_short = fits & (len_i <= bound)
    second_short = fits & ~first_short & (len_j <= bound)
    rest = ~(first_short | second_short)

    class _Stats:
        scores = np.zeros(len(i_idx), dtype=np.int32)
        trace_len = np.zeros(len(i_idx), dtype=np.int64)
        gap_counts = np.zeros((len(i_idx), 2), dtype=np.int32)
    res = _Stats()
    for mask in (first_short, second_short, rest):
        if not mask.any():
            continue
        with DeviceBatch(pool.codes, pool.offsets, pool.codes, pool.offsets, scoring, gap_penalty,
                         terminal_penalty, False, None, False, idx1=i_idx[mask], idx2=j_idx[mask],
                         stats=True) as batch:
            batch.run()
            part = batch.fetch()
        res.scores[mask] = part.scores
        res.trace_len[mask] = part.trace_len
        res.gap_counts[mask] = part.gap_counts
    scores = np.zeros((n, n), dtype=np.int32)
    scores[i_idx, j_idx] = res.scores
    gap_open, gap_ext = (gap_penalty if isinstance(gap_penalty, tuple) else (gap_penalty, gap_penalty))
    alphabet_size = scoring.shape[0]
    counts = np.zeros((n, alphabet_size), dtype=np.float64)
    for k in range(n):
        counts[k] = np.bincount(pool.codes[pool.offsets[k]:pool.offsets[k + 1]], minlength=alphabet_size)
    expected = counts @ scoring.astype(np.float64) @ counts.T     # sum_ab M[a,b] c_i[a] c_j[b]
    off = i_idx != j_idx
    a, b = i_idx[off], j_idx[off]
    length = res.trace_len[off].astype(np.float64)
    s_rand = (expected[a, b] / length + res.gap_counts[off, 0] * gap_open
              + res.gap_counts[off, 1] * gap_ext).astype(np.float32)
    s_max = ((scores[a, a].astype(np.float64) + scores[b, b]) / 2.0).astype(np.float32)
    s_ij = scores[a, b].astype(np.float32)
    bad = np.nonzero(s_ij < s_rand)[0]
    if bad.size:
        raise ValueError(
            f"The randomized alignment of sequences {b[bad[0]]} and {a[bad[0]]} "
            f"scores better than the real pairwise alignment, "
            f"cannot calculate proper pairwise distance")
    dist = np.zeros((n, n), dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        d = -np.log((s_ij - s_rand) / (s_max - s_rand)).astype(np.float32)
    dist[a, b] = d
    dist[b, a] = d
    return dist


def _resolve_scoring(matrix, packed1, packed2):
    if isinstance(matrix, tuple):
        match, mismatch = int(matrix[0]), int(matrix[1])
        # A (match, mismatch) pair over small alphabets is the same function as the matrix
        # `match` on the diagonal, `mismatch` elsewhere (scoring.rs:101-108): handing the kernels
        # that matrix makes such batches eligible for the packed 16-bit route.
        sizes = [len(p.alphabet) for p in (packed1, packed2) if p.alphabet is not None]
        if len(sizes) == 2 and max(sizes) <= 64 and packed1.codes.dtype == np.uint8 \
                and packed2.codes.dtype == np.uint8 and -128 <= min(match, mismatch) and max(match, mismatch) <= 127:
            table = np.full((sizes[0], sizes[1]), mismatch, dtype=np.int32)
            np.fill_diagonal(table, match)
            return table
        return (match, mismatch)
    if isinstance(matrix, SubstitutionMatrix):
        for packed, alph in ((packed1, matrix.get_alphabet1()), (packed2, matrix.get_alphabet2())):
            if packed.alphabet is not None and not alph.extends(packed.alphabet):
                raise ValueError("The sequences' alphabets do not fit the matrix")
        return matrix.score_matrix()
    return np.ascontiguousarray(matrix, dtype=np.int32)


def align_optimal_batch(seqs1, seqs2, matrix, gap_penalty=-10, terminal_penalty=True,
                        local=False, score_only=False):
    """
    ``align_optimal(seqs1[k], seqs2[k], ..., max_number=1)`` for every ``k`` in
    one GPU batch.  `seqs1` / `seqs2` are equally long lists of ``Sequence`` or
    :class:`PackedSequences`.  Returns a :class:`BatchResult`.
    """
    _check_gap_penalty(gap_penalty)
    p1, p2 = pack_sequences(seqs1), pack_sequences(seqs2)
    if len(p1) != len(p2):
        raise ValueError("Both lists must hold the same number of sequences")
    scoring = _resolve_scoring(matrix, p1, p2)
    # The library routes a batch as a whole: a few pairs beyond the packed kernels' limits
    # (a titin among 10^6 ordinary proteins) would send everything to the 32-bit kernel.  Such
    # pairs are run as a batch of their own and the results are stitched together.
    len1, len2 = p1.lengths, p2.lengths
    outlier = np.maximum(len1, len2) > _PACKED_MAX_LEN
    if not isinstance(scoring, tuple):
        outlier |= np.minimum(len1, len2) > _packed_shorter_bound(scoring, gap_penalty)
    if len(p1) >= 128 and outlier.any() and outlier.sum() * 8 <= len(p1):
        parts, index = [], []
        for mask in (~outlier, outlier):
            ids = np.nonzero(mask)[0]
            q1, q2 = _take(p1, ids), _take(p2, ids)
            part = align_batch_arrays(q1.codes, q1.offsets, q2.codes, q2.offsets, scoring, gap_penalty,
                                      terminal_penalty, local, None, score_only)
            parts.append(part)
            index.append(ids)
        return CompositeBatchResult(parts, index, p1.sequences, p2.sequences)
    result = align_batch_arrays(p1.codes, p1.offsets, p2.codes, p2.offsets, scoring, gap_penalty,
                                terminal_penalty, local, None, score_only)
    result.seqs1, result.seqs2 = p1.sequences, p2.sequences
    return result


def _take(packed, ids):
    """The sequences `ids` of a :class:`PackedSequences` as a new one (codes gathered)."""
    lengths = packed.lengths[ids]
    offsets = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(lengths, out=offsets[1:])
    # positions of the gathered residues: start of each sequence + running index inside it
    starts = np.repeat(packed.offsets[:-1][ids] - offsets[:-1], lengths)
    codes = packed.codes[starts + np.arange(int(offsets[-1]), dtype=np.int64)]
    return PackedSequences(codes, offsets, packed.alphabet, None)


class CompositeBatchResult:
    """
    Results of a batch that was run in several parts (pairs outside the packed kernels' limits
    separately from the rest); same interface as :class:`BatchResult`, pair numbering of the
    original batch.
    """

    def __init__(self, parts, index, seqs1=None, seqs2=None):
        self.parts, self.index = parts, index
        n = sum(len(ids) for ids in index)
        self.scores = np.empty(n, dtype=np.int32)
        self._part = np.empty(n, dtype=np.int32)
        self._local = np.empty(n, dtype=np.int64)
        for k, (part, ids) in enumerate(zip(parts, index)):
            self.scores[ids] = part.scores
            self._part[ids] = k
            self._local[ids] = np.arange(len(ids))
        self.trace_len = None
        if all(part.trace_len is not None for part in parts):
            self.trace_len = np.empty(n, dtype=np.int64)
            for part, ids in zip(parts, index):
                self.trace_len[ids] = part.trace_len
        self.seqs1, self.seqs2 = seqs1, seqs2

    def __len__(self):
        return len(self.scores)

    def trace(self, k):
        return self.parts[self._part[k]].trace(int(self._local[k]))

    def traces(self):
        out = [None] * len(self)
        for part, ids in zip(self.parts, self.index):
            for local, trace in enumerate(part.traces()):
                out[ids[local]] = trace
        return out

    def alignment(self, k, seq1=None, seq2=None):
        seq1 = self.seqs1[k] if seq1 is None else seq1
        seq2 = self.seqs2[k] if seq2 is None else seq2
        return Alignment([seq1, seq2], self.trace(k), int(self.scores[k]))

    def alignments(self):
        traces = self.traces()
        return [Alignment([self.seqs1[k], self.seqs2[k]], traces[k], int(self.scores[k]))
                for k in range(len(self))]


def align_banded_batch(seqs1, seqs2, matrix, bands, gap_penalty=-10, local=False,
                       score_only=False):
    """
    ``align_banded(seqs1[k], seqs2[k], matrix, bands[k], ..., max_number=1)`` for
    every ``k``.  `bands` is one ``(lower, upper)`` tuple for all pairs or an
    ``(n, 2)`` array.  Pairs whose second sequence is shorter are swapped
    internally exactly as the reference does (``banded.py:215-224``) and
    un-swapped in the returned traces.
    """
    _check_gap_penalty(gap_penalty)
    p1, p2 = pack_sequences(seqs1), pack_sequences(seqs2)
    n = len(p1)
    if n != len(p2):
        raise ValueError("Both lists must hold the same number of sequences")
    bands = np.asarray(bands, dtype=np.int64)
    if bands.ndim == 1:
        bands = np.broadcast_to(bands, (n, 2))
    len1, len2 = p1.lengths, p2.lengths
    swapped = len2 < len1
    scoring = _resolve_scoring(matrix, p1, p2)
    if swapped.any():
        if not swapped.all():
            # mixed orientation: gather into shorter-first order
            dtype = np.promote_types(p1.codes.dtype, p2.codes.dtype)
            a_len = np.where(swapped, len2, len1)
            b_len = np.where(swapped, len1, len2)
            a_off = np.concatenate(([0], np.cumsum(a_len)))
            b_off = np.concatenate(([0], np.cumsum(b_len)))
            a_codes = np.empty(int(a_off[-1]), dtype=dtype)
            b_codes = np.empty(int(b_off[-1]), dtype=dtype)
            for k in range(n):
                s1 = p1.codes[p1.offsets[k]:p1.offsets[k + 1]]
                s2 = p2.codes[p2.offsets[k]:p2.offsets[k + 1]]
                if swapped[k]:
                    s1, s2 = s2, s1
                a_codes[a_off[k]:a_off[k + 1]] = s1
                b_codes[b_off[k]:b_off[k + 1]] = s2
            if not isinstance(scoring, tuple) and not np.array_equal(scoring, scoring.T):
                raise ValueError(
                    "A batch that mixes orientations needs a symmetric substitution matrix")
            q1 = PackedSequences(a_codes, a_off)
            q2 = PackedSequences(b_codes, b_off)
        else:
            q1, q2 = p2, p1
            if not isinstance(scoring, tuple):
                scoring = np.ascontiguousarray(scoring.T)
    else:
        q1, q2 = p1, p2
    eff = np.empty((n, 2), dtype=np.int64)
    l1, l2 = q1.lengths, q2.lengths
    for k in range(n):
        band = (-int(bands[k, 0]), -int(bands[k, 1])) if swapped[k] else (
            int(bands[k, 0]), int(bands[k, 1]))
        eff[k] = crop_band(int(l1[k]), int(l2[k]), band)
    result = align_batch_arrays(q1.codes, q1.offsets, q2.codes, q2.offsets, scoring, gap_penalty,
                                True, local, eff, score_only)
    result.swapped = swapped if swapped.any() else None
    result.seqs1, result.seqs2 = p1.sequences, p2.sequences
    return result
