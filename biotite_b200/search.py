"""
One-vs-many scoring at database scale (BASELINE config 3: many queries against a sequence
database, local Smith-Waterman, score only).  The reference would call ``align_optimal(q, d,
..., local=True, score_only=True)`` for every (query, database sequence) pair
(``src/biotite/sequence/align/pairwise.py:117-244``); here the database is cut into chunks,
every chunk becomes one indexed GPU batch (no pair is materialised beyond two int32 index
arrays), and the host prepares chunk ``k + 1`` (index arrays, global sort, window plan, uploads)
on a worker thread while the GPU scores chunk ``k``.

Multi-GPU: the *database* is sharded (every rank holds all queries):
:func:`database_shard` gives rank ``r`` of ``world`` its slice with equal residue counts; ranks
score their slices independently, the only exchange is the final gather of the score columns.
"""

from __future__ import annotations

import os
import sys
import threading
import time

import numpy as np

from .batch import DeviceBatch, _check_gap_penalty, _resolve_scoring, pack_sequences

__all__ = ["search_scores", "search_scores_sharded", "database_shard"]


def database_shard(db_lengths, rank, world_size):
    """``(begin, end)`` of rank `rank`'s contiguous database slice; slices hold (nearly) the
    same number of residues, hence the same number of DP cells against any query set."""
    lengths = np.asarray(db_lengths, dtype=np.int64)
    prefix = np.concatenate(([0], np.cumsum(lengths)))
    targets = prefix[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(prefix, targets, side="left")
    bounds = np.maximum.accumulate(np.concatenate(([0], cuts, [len(lengths)])))
    return int(bounds[rank]), int(bounds[rank + 1])


def _merge_top(best, new, k):
    """Stable merge of two ``(scores, indices)`` candidate lists per row: higher score first, equal
    scores in the order given (callers append candidates with ascending database indices)."""
    if best is None:
        return new
    scores = np.concatenate((best[0], new[0]), axis=1)
    index = np.concatenate((best[1], new[1]), axis=1)
    order = np.argsort(-scores.astype(np.int64), axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, order, axis=1), np.take_along_axis(index, order, axis=1)


def search_scores(queries, database, matrix, gap_penalty=(-10, -1), local=True, terminal_penalty=True,
                  chunk_size=100_000, db_range=None, out=None, timings=None, devices=None, top_k=None):
    """
    Optimal alignment score of every query against every database sequence.

    Parameters
    ----------
    queries, database : list of Sequence or PackedSequences
    matrix, gap_penalty, local, terminal_penalty
        As for ``align_optimal``; the default is local affine scoring.
    chunk_size : int
        Database sequences per GPU batch (``len(queries) * chunk_size`` pairs).
    db_range : (int, int), optional
        Score only this slice of the database (see :func:`database_shard`).
    out : ndarray, optional
        Preallocated ``(len(queries), n_db)`` int32 result.
    timings : dict, optional
        Receives ``device_ms`` (sum of the CUDA-event times of the batches) and ``cells``.
    devices : list of int, optional
        Several GPUs of the node from this one process: the database slice is cut into
        :func:`database_shard` parts, one host thread per device scores its part into its own
        columns of the result (no exchange between devices).
    top_k : int, optional
        Keep only the `top_k` best database hits of every query: the selection runs on the GPU
        chunk by chunk (``bt_batch_fetch_topk``), the score matrix is never materialised - a
        10^3 x 10^7 search returns 8 `top_k` bytes per query instead of 40 GB.

    Returns
    -------
    scores : ndarray, shape=(len(queries), n_db), dtype=int32
        Without `top_k`.
    (scores, indices) : ndarrays, shape=(len(queries), top_k), dtype=int32
        With `top_k`: the best scores of every query in descending order (equal scores by
        ascending database index) and the database indices they belong to; rows are padded with
        index -1 when the database slice holds fewer than `top_k` sequences.
    """
    _check_gap_penalty(gap_penalty)
    q = queries if hasattr(queries, "offsets") else pack_sequences(queries)
    d = database if hasattr(database, "offsets") else pack_sequences(database)
    scoring = _resolve_scoring(matrix, q, d)
    begin, end = (0, len(d)) if db_range is None else (int(db_range[0]), int(db_range[1]))
    n_q, n_db = len(q), end - begin
    if top_k is not None:
        top_k = int(top_k)
        if top_k < 1:
            raise ValueError("top_k must be at least 1")
        if devices is not None and len(devices) > 1:
            raise ValueError("top_k with several devices: use search_scores_sharded (one process per GPU)")
        scores = None
    else:
        scores = np.empty((n_q, n_db), dtype=np.int32) if out is None else out
        if scores.shape != (n_q, n_db):
            raise ValueError(f"out must have shape {(n_q, n_db)}")
    if devices is not None and len(devices) > 1:
        from . import _cabi
        lib = _cabi.load()
        world = len(devices)
        parts = [database_shard(d.lengths[begin:end], r, world) for r in range(world)]
        results, errors = [None] * world, [None] * world

        def run(r):
            try:
                _cabi.check(lib.bt_set_device(int(devices[r])))
                lo, hi = begin + parts[r][0], begin + parts[r][1]
                t = {}
                if hi > lo:
                    search_scores(q, d, scoring, gap_penalty, local, terminal_penalty, chunk_size, (lo, hi),
                                  scores[:, lo - begin:hi - begin], t)
                results[r] = t
            except BaseException as exc:   # re-raised on the calling thread
                errors[r] = exc

        threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for exc in errors:
            if exc is not None:
                raise exc
        if timings is not None:
            # the devices work side by side: the batch takes as long as the slowest of them
            timings["device_ms"] = max((t.get("device_ms", 0.0) for t in results), default=0.0)
            timings["cells"] = sum(t.get("cells", 0) for t in results)
            timings["kernel"] = "+".join(sorted({t["kernel"] for t in results if t.get("kernel")}))
        return scores
    # a batch holds at most 2^31 - 1 pairs; 64 database sequences are the smallest useful chunk
    chunk_size = int(max(1, min(chunk_size, (2**31 - 1) // max(n_q, 1))))
    chunks = [(c, min(end, c + chunk_size)) for c in range(begin, end, chunk_size)]
    total_ms, total_cells, kernels = 0.0, 0, set()

    def prepare(lo, hi):
        # database chunk as its own pool: offsets rebased, codes sliced (views, no copies)
        off = d.offsets[lo:hi + 1] - d.offsets[lo]
        codes = d.codes[d.offsets[lo]:d.offsets[hi]]
        if not isinstance(scoring, tuple) and q.codes.dtype == np.uint8 and codes.dtype == np.uint8:
            # the cross product is enumerated, sorted and packed on the GPU: no index arrays
            try:
                return DeviceBatch.generated(q.codes, q.offsets, codes, off, scoring, gap_penalty, terminal_penalty,
                                             local, kind="cross")
            except ValueError:
                pass   # a mode outside the packed kernels
        idx1 = np.repeat(np.arange(n_q, dtype=np.int32), hi - lo)
        idx2 = np.tile(np.arange(hi - lo, dtype=np.int32), n_q)
        return DeviceBatch(q.codes, q.offsets, codes, off, scoring, gap_penalty, terminal_penalty, local,
                           None, True, idx1=idx1, idx2=idx2)

    # one page-locked staging buffer for the chunks' scores (a pageable D2H target costs
    # more than the GPU work of small chunks)
    from . import _cabi
    staging = None if top_k is not None else _cabi.pinned_empty((n_q * min(chunk_size, max(n_db, 1)),), np.int32)
    pending = {}
    best = None

    device = _cabi.load().bt_get_device()

    def prepare_into(k):
        try:
            # a new host thread starts on device 0: take the caller's device
            if device >= 0:
                _cabi.check(_cabi.load().bt_set_device(device))
            pending[k] = prepare(*chunks[k])
        except BaseException as exc:   # re-raised on the consuming thread
            pending[k] = exc

    worker = None
    if chunks:
        prepare_into(0)
    for k, (lo, hi) in enumerate(chunks):
        batch = pending.pop(k)
        if isinstance(batch, BaseException):
            raise batch
        if k + 1 < len(chunks):
            worker = threading.Thread(target=prepare_into, args=(k + 1,))
            worker.start()
        try:
            t0 = time.perf_counter()
            total_ms += batch.run()
            total_cells += batch.cells
            kernels.add(batch.kernel)
            t1 = time.perf_counter()
            if top_k is not None:
                if getattr(batch, "kind", None) == 0:
                    cand_scores, cand_cols = batch.top_k(top_k)
                else:   # index-array batch (a mode outside the packed kernels): select on the host
                    full = batch.fetch().scores.reshape(n_q, hi - lo)
                    kk = min(top_k, hi - lo)
                    cand_cols = np.argsort(-full.astype(np.int64), axis=1, kind="stable")[:, :kk].astype(np.int32)
                    cand_scores = np.take_along_axis(full, cand_cols, axis=1)
                cand_index = np.where(cand_cols >= 0, cand_cols + lo, -1).astype(np.int32)
                best = _merge_top(best, (cand_scores, cand_index), top_k)
            else:
                part = staging[:n_q * (hi - lo)]
                batch.fetch(out=(part, None, None, None))
                scores[:, lo - begin:hi - begin] = part.reshape(n_q, hi - lo)
            t2 = time.perf_counter()
        finally:
            batch.close()
            if worker is not None:
                worker.join()
                worker = None
        if os.environ.get("B200ALIGN_TRACE"):
            print(f"[search] chunk {k}: run {t1 - t0:.2f} s, fetch {t2 - t1:.2f} s, wait for next chunk "
                  f"{time.perf_counter() - t2:.2f} s", file=sys.stderr)
    if timings is not None:
        timings["device_ms"] = total_ms
        timings["cells"] = total_cells
        timings["kernel"] = "+".join(sorted(kernels))
    if top_k is not None:
        if best is None:
            best = (np.zeros((n_q, 0), dtype=np.int32), np.zeros((n_q, 0), dtype=np.int32))
        if best[0].shape[1] < top_k:   # fewer database sequences than top_k: pad
            pad = top_k - best[0].shape[1]
            best = (np.pad(best[0], ((0, 0), (0, pad)), constant_values=np.iinfo(np.int32).min),
                    np.pad(best[1], ((0, 0), (0, pad)), constant_values=-1))
        return best
    return scores


def search_scores_sharded(queries, database, matrix, gap_penalty=(-10, -1), local=True,
                          terminal_penalty=True, chunk_size=100_000, group=None, compute=None, dst=0,
                          timings=None, top_k=None):
    """
    Collective call (one process per GPU, ``torch.distributed``): rank ``r`` scores all queries
    against its :func:`database_shard`; rank `dst` receives the full ``(n_queries, n_db)``
    matrix, the other ranks ``None``.  No collective on the data path, only the final gather.
    `compute` defaults to :func:`search_scores`; the CPU test-suite injects the oracle.
    """
    import os

    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    d = database if hasattr(database, "offsets") else pack_sequences(database)
    lo, hi = database_shard(d.lengths, rank, world)
    if compute is None:
        from . import _cabi
        _cabi.check(_cabi.load().bt_set_device(int(os.environ.get("LOCAL_RANK", rank))))
        compute = search_scores
    if top_k is not None:
        # every rank keeps the best hits of its slice (database indices are global); the exchange
        # is 8 top_k bytes per query and rank, merged on `dst` in rank order = ascending indices
        from .distributed import gather_arrays
        kw = {} if compute is not search_scores else {"timings": timings}
        part = compute(queries, d, matrix, gap_penalty, local, terminal_penalty, chunk_size, (lo, hi), top_k=top_k, **kw)
        n_q = part[0].shape[0]
        both = np.concatenate((part[0].reshape(-1), part[1].reshape(-1))).astype(np.int32)
        parts = gather_arrays(both, dst, group)
        if rank != dst:
            return None
        best = None
        for p in parts:
            best = _merge_top(best, (p[:n_q * top_k].reshape(n_q, top_k), p[n_q * top_k:].reshape(n_q, top_k)), top_k)
        return best
    if compute is search_scores:
        # page-locked result block: it goes back to the device for the gather at copy-engine speed
        from . import _cabi
        q = queries if hasattr(queries, "offsets") else pack_sequences(queries)
        part = _cabi.pinned_empty((len(q), hi - lo), np.int32)
        compute(q, d, matrix, gap_penalty, local, terminal_penalty, chunk_size, (lo, hi), out=part, timings=timings)
    else:
        part = compute(queries, d, matrix, gap_penalty, local, terminal_penalty, chunk_size, (lo, hi))
    # the only exchange: the score columns of every rank's database slice, joined on the device
    from .distributed import gather_columns
    return gather_columns(part, dst, group)
