"""
ctypes binding of ``libb200align.so`` (C ABI in ``include/b200align.h``).

This is the only place the Python layer touches native code; it plays the role
of the reference's PyO3 module ``biotite.rust.sequence.align``
(``src/rust/sequence/align/mod.rs:23-24``).  The library must be present
(``__graft_entry__.build()`` compiles it in-tree): there is no Python or CPU
fallback, a missing library or a missing GPU raises.
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200ALIGN_LIB", _PKG / "libb200align.so"))

BT_OK = 0
BT_ERR_INVALID_ARG = -1
BT_ERR_INVALID_CODE = -2
BT_ERR_CUDA = -3
BT_ERR_OOM = -4
BT_ERR_RANGE = -5

EXPORTED_SYMBOLS = [
    "bt_last_error", "bt_version", "bt_device_count", "bt_set_device", "bt_get_device",
    "bt_align_pair", "bt_tracelist_count", "bt_tracelist_length", "bt_tracelist_copy",
    "bt_tracelist_free", "bt_batch_create", "bt_batch_create_indexed", "bt_batch_ops_bytes", "bt_batch_run", "bt_batch_fetch_scores",
    "bt_batch_fetch_traces", "bt_batch_fetch_cigar", "bt_batch_fetch_gapped", "bt_batch_timings", "bt_microbench", "bt_batch_cells", "bt_batch_launches", "bt_batch_kernel",
    "bt_batch_destroy", "bt_align_batch", "bt_xdrop_regions", "bt_debug_plan", "bt_expand_traces", "bt_pinned_alloc",
    "bt_pinned_free", "bt_pool_trim", "bt_ops2_capacity", "bt_align_batch_compact", "bt_expand_traces2", "bt_align_batch_multi", "bt_batch_windows",
    "bt_batch_create_generated", "bt_batch_fetch_score_matrix", "bt_debug_device_plan", "bt_batch_fetch_topk",
]


class BtParams(ctypes.Structure):
    _fields_ = [
        ("affine", ctypes.c_int32),
        ("gap_open", ctypes.c_int32),
        ("gap_ext", ctypes.c_int32),
        ("local", ctypes.c_int32),
        ("terminal_penalty", ctypes.c_int32),
        ("use_matrix", ctypes.c_int32),
        ("match_score", ctypes.c_int32),
        ("mismatch_score", ctypes.c_int32),
        ("n_rows", ctypes.c_int32),
        ("n_cols", ctypes.c_int32),
        ("banded", ctypes.c_int32),
        ("code_bytes", ctypes.c_int32),
    ]


_lib = None

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32


def load():
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a).  biotite_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.bt_last_error.restype = ctypes.c_char_p
    lib.bt_version.restype = ctypes.c_char_p
    lib.bt_device_count.restype = ctypes.c_int
    lib.bt_set_device.argtypes = [ctypes.c_int]
    lib.bt_get_device.restype = ctypes.c_int
    lib.bt_align_pair.argtypes = [
        ctypes.POINTER(BtParams), _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i64,
        ctypes.POINTER(_i32), ctypes.POINTER(_vp)]
    lib.bt_tracelist_count.restype = _i64
    lib.bt_tracelist_count.argtypes = [_vp]
    lib.bt_tracelist_length.restype = _i64
    lib.bt_tracelist_length.argtypes = [_vp, _i64]
    lib.bt_tracelist_copy.argtypes = [_vp, _i64, _vp]
    lib.bt_tracelist_free.restype = None
    lib.bt_tracelist_free.argtypes = [_vp]
    lib.bt_batch_create.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32,
        ctypes.POINTER(_vp)]
    lib.bt_batch_create_indexed.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i32,
        ctypes.POINTER(_vp)]
    lib.bt_batch_create_generated.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _i64, _vp, _vp, _i64, _i32, _vp, _i32, ctypes.POINTER(_vp)]
    lib.bt_batch_fetch_score_matrix.argtypes = [_vp, _vp]
    lib.bt_batch_fetch_topk.argtypes = [_vp, _i32, _vp, _vp]
    lib.bt_debug_device_plan.argtypes = [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp]
    lib.bt_batch_ops_bytes.restype = _i64
    lib.bt_batch_ops_bytes.argtypes = [_vp]
    lib.bt_batch_run.argtypes = [_vp, ctypes.POINTER(ctypes.c_float)]
    lib.bt_batch_fetch_scores.argtypes = [_vp, _vp]
    lib.bt_batch_fetch_traces.argtypes = [_vp, _vp, _vp, _vp]
    lib.bt_batch_fetch_cigar.argtypes = [_vp, _i32, _vp, _vp, _vp]
    lib.bt_batch_fetch_gapped.argtypes = [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]
    lib.bt_batch_timings.argtypes = [_vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    lib.bt_microbench.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.bt_batch_cells.restype = _i64
    lib.bt_batch_cells.argtypes = [_vp]
    lib.bt_batch_launches.restype = _i64
    lib.bt_batch_launches.argtypes = [_vp]
    lib.bt_batch_windows.restype = _i64
    lib.bt_batch_windows.argtypes = [_vp]
    lib.bt_batch_kernel.restype = ctypes.c_char_p
    lib.bt_batch_kernel.argtypes = [_vp]
    lib.bt_batch_destroy.restype = None
    lib.bt_batch_destroy.argtypes = [_vp]
    lib.bt_align_batch.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp]
    lib.bt_ops2_capacity.restype = _i64
    lib.bt_ops2_capacity.argtypes = [_vp, _vp, _i64]
    lib.bt_align_batch_compact.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp,
        ctypes.POINTER(_i64)]
    lib.bt_expand_traces2.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.bt_align_batch_multi.argtypes = [
        _vp, _i32, ctypes.POINTER(BtParams), _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _vp,
        ctypes.POINTER(_i64), _vp]
    lib.bt_xdrop_regions.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i64, _i32, _i64, _vp, _vp,
        ctypes.POINTER(_vp)]
    lib.bt_debug_plan.argtypes = [
        ctypes.POINTER(BtParams), _vp, _vp, _i64, _i32, _i32, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]
    lib.bt_expand_traces.argtypes = [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    lib.bt_pinned_alloc.restype = _vp
    lib.bt_pinned_alloc.argtypes = [ctypes.c_size_t]
    lib.bt_pinned_free.restype = None
    lib.bt_pinned_free.argtypes = [_vp]
    lib.bt_pool_trim.restype = None
    _lib = lib
    return lib


def check(status: int):
    """Map a BtStatus to the exception type the reference's users would see."""
    if status == BT_OK:
        return
    msg = load().bt_last_error().decode("utf-8", "replace")
    if status in (BT_ERR_INVALID_ARG, BT_ERR_INVALID_CODE):
        raise ValueError(msg)
    if status == BT_ERR_OOM:
        raise MemoryError(msg)
    if status == BT_ERR_RANGE:
        raise OverflowError(msg)
    raise RuntimeError(msg)


def ptr(array):
    return None if array is None else array.ctypes.data_as(_vp)


def make_params(scoring, gap_penalty, terminal_penalty, local, banded, code_dtype):
    """
    Build a :class:`BtParams` from the values the reference hands to its native
    layer: `scoring` is an int32 2-D array or a ``(match, mismatch)`` tuple,
    `gap_penalty` an int or an ``(open, extend)`` tuple.  Returns
    ``(params, matrix_or_None)``; the matrix must be kept alive by the caller.
    """
    P = BtParams()
    if isinstance(gap_penalty, tuple):
        P.affine, P.gap_open, P.gap_ext = 1, int(gap_penalty[0]), int(gap_penalty[1])
    else:
        P.affine, P.gap_open, P.gap_ext = 0, int(gap_penalty), int(gap_penalty)
    P.local = int(bool(local))
    P.terminal_penalty = int(bool(terminal_penalty))
    matrix = None
    if isinstance(scoring, tuple):
        P.use_matrix = 0
        P.match_score, P.mismatch_score = int(scoring[0]), int(scoring[1])
    else:
        matrix = np.ascontiguousarray(scoring, dtype=np.int32)
        if matrix.ndim != 2:
            raise ValueError("The substitution matrix must be two-dimensional")
        P.use_matrix = 1
        P.n_rows, P.n_cols = matrix.shape
    P.banded = int(bool(banded))
    dtype = np.dtype(code_dtype)
    if dtype.kind != "u" or dtype.itemsize not in (1, 2, 4, 8):
        raise TypeError(f"Unsupported dtype: {dtype}")
    P.code_bytes = dtype.itemsize
    return P, matrix


def pinned_empty(shape, dtype):
    """
    A NumPy array backed by page-locked host memory (``cudaHostAlloc``), for
    full-speed asynchronous H2D / D2H copies.  Freed when the array is garbage
    collected.
    """
    lib = load()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(1, count * dtype.itemsize)
    address = lib.bt_pinned_alloc(nbytes)
    if not address:
        raise MemoryError(lib.bt_last_error().decode())
    buf = (ctypes.c_char * nbytes).from_address(address)
    holder = _PinnedHolder(address)
    array = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _PINNED[id(buf)] = holder  # keep alive as long as the ctypes buffer lives
    import weakref
    weakref.finalize(buf, _release, id(buf))
    return array


class _PinnedHolder:
    def __init__(self, address):
        self.address = address


_PINNED = {}


def _release(key):
    holder = _PINNED.pop(key, None)
    if holder is not None and _lib is not None:
        _lib.bt_pinned_free(holder.address)
