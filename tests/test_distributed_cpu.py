"""World-size-2 ``gloo`` test of the multi-process path: sharding by DP work and the final
gather.  The compute step is injected (CPU oracle) because there is no GPU here; on the GPU
box the same function runs with the CUDA batch entry point (tests/test_gpu_parity.py)."""

import os
import socket

import numpy as np
import pytest

import biotite_b200 as bt
from biotite_b200.distributed import shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(seed=0, n_pairs=37):
    rng = np.random.default_rng(seed)
    len1 = rng.integers(1, 60, n_pairs)
    len2 = rng.integers(1, 60, n_pairs)
    off1 = np.concatenate(([0], np.cumsum(len1))).astype(np.int64)
    off2 = np.concatenate(([0], np.cumsum(len2))).astype(np.int64)
    codes1 = rng.integers(0, 20, off1[-1]).astype(np.uint8)
    codes2 = rng.integers(0, 20, off2[-1]).astype(np.uint8)
    return codes1, off1, codes2, off2


def _oracle_compute(codes1, off1, codes2, off2, scoring, gap, term, local, bands, score_only):
    """Stand-in for align_batch_arrays with the same result layout."""
    from biotite_b200.batch import BatchResult
    from oracle import oracle
    scores, traces = oracle.batch(codes1, off1, codes2, off2, scoring, gap, term, local, score_only,
                                  n_threads=2)
    n = len(off1) - 1
    ops_off = off1 + off2
    trace_len = np.zeros(n, dtype=np.int64)
    first = np.zeros((n, 2), dtype=np.int32)
    ops = np.zeros(int(ops_off[-1]), dtype=np.uint8)
    for k in range(n):
        t = traces[k]
        trace_len[k] = len(t)
        if len(t) == 0:
            continue
        # first alignment column -> table cell; following rows -> step bytes (end-to-start)
        i = t[0, 0] + 1 if t[0, 0] >= 0 else 0
        j = t[0, 1] + 1 if t[0, 1] >= 0 else 0
        first[k] = (i, j)
        step = np.where(t[:, 0] == -1, 1, np.where(t[:, 1] == -1, 2, 0)).astype(np.uint8)
        ops[ops_off[k]:ops_off[k] + len(t)] = step[::-1]
    return BatchResult(scores, trace_len, first, ops, ops_off)


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from biotite_b200.distributed import align_batch_sharded
    codes1, off1, codes2, off2 = _make_batch()
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    res = align_batch_sharded(codes1, off1, codes2, off2, matrix, (-10, -1), True, False,
                              compute=_oracle_compute)
    if rank == 0:
        flat, rows_off = res.traces_flat()
        np.savez(out_path, scores=res.scores, flat=flat, rows_off=rows_off)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_balance_cells():
    codes1, off1, codes2, off2 = _make_batch(n_pairs=1000)
    for world in (1, 2, 4, 8):
        b = shard_bounds(off1, off2, world)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0)
        cells = np.diff(off1) * np.diff(off2)
        per = np.array([cells[b[r]:b[r + 1]].sum() for r in range(world)])
        assert per.max() <= per.mean() + cells.max()


def test_two_rank_gloo_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "result.npz"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = np.load(out)
    codes1, off1, codes2, off2 = _make_batch()
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    ref = _oracle_compute(codes1, off1, codes2, off2, matrix, (-10, -1), True, False, None, False)
    flat, rows_off = ref.traces_flat()
    assert np.array_equal(got["scores"], ref.scores)
    assert np.array_equal(got["rows_off"], rows_off)
    assert np.array_equal(got["flat"], flat)


# ---- database search (config 3): the database is sharded, every rank holds all queries -----------

def _search_inputs(seed=5, n_q=5, n_db=41):
    rng = np.random.default_rng(seed)
    from biotite_b200.batch import PackedSequences
    def pool(count, lo, hi):
        lengths = rng.integers(lo, hi, count)
        off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
        return PackedSequences(rng.integers(0, 20, off[-1]).astype(np.uint8), off, bt.ProteinSequence.alphabet)
    return pool(n_q, 20, 60), pool(n_db, 5, 120)


def _oracle_search(queries, database, matrix, gap, local, term, chunk_size, db_range):
    from oracle import oracle
    lo, hi = db_range
    out = np.zeros((len(queries), hi - lo), dtype=np.int32)
    for i in range(len(queries)):
        qc = queries.codes[queries.offsets[i]:queries.offsets[i + 1]]
        for j in range(lo, hi):
            dc = database.codes[database.offsets[j]:database.offsets[j + 1]]
            out[i, j - lo] = oracle.align_optimal(qc, dc, matrix.score_matrix(), gap, term, local, True, 1)[1]
    return out


def _search_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queries, database = _search_inputs()
    res = bt.search_scores_sharded(queries, database, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1),
                                   compute=_oracle_search)
    if rank == 0:
        np.save(out_path, res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_database_shard_balances_residues():
    rng = np.random.default_rng(2)
    lengths = np.clip(rng.lognormal(np.log(270), 0.7, 5000), 30, 5000).astype(np.int64)
    for world in (1, 2, 3, 8):
        bounds = [bt.database_shard(lengths, r, world) for r in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == len(lengths)
        assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
        per = np.array([lengths[a:b].sum() for a, b in bounds])
        assert per.max() <= per.mean() + lengths.max()


def test_two_rank_gloo_database_search(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "scores.npy"
    mp.spawn(_search_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    queries, database = _search_inputs()
    ref = _oracle_search(queries, database, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1), True, True,
                         0, (0, len(database)))
    assert np.array_equal(np.load(out), ref)


# ---- per-query top-k over a sharded database: every rank keeps its best hits, rank 0 merges ----------

def _oracle_search_topk(queries, database, matrix, gap, local, term, chunk_size, db_range, top_k=None):
    """Stand-in for search_scores(top_k=...): the slice's scores by the oracle, the selection by numpy."""
    full = _oracle_search(queries, database, matrix, gap, local, term, chunk_size, db_range)
    lo = db_range[0]
    kk = min(top_k, full.shape[1])
    order = np.argsort(-full.astype(np.int64), axis=1, kind="stable")[:, :kk]
    scores = np.take_along_axis(full, order, axis=1)
    index = (order + lo).astype(np.int32)
    pad = top_k - kk
    return (np.pad(scores, ((0, 0), (0, pad)), constant_values=np.iinfo(np.int32).min),
            np.pad(index, ((0, 0), (0, pad)), constant_values=-1))


def _topk_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    queries, database = _search_inputs()
    res = bt.search_scores_sharded(queries, database, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1),
                                   compute=_oracle_search_topk, top_k=30)
    if rank == 0:
        np.savez(out_path, scores=res[0], index=res[1])
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_top_k_merge(tmp_path):
    """Rank order = ascending database indices, so the stable merge keeps equal scores in index order;
    30 hits of a 41-sequence database make every rank's list longer than its slice (padding)."""
    import torch.multiprocessing as mp
    out = tmp_path / "topk.npz"
    mp.spawn(_topk_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    queries, database = _search_inputs()
    ref = _oracle_search(queries, database, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1), True, True,
                         0, (0, len(database)))
    order = np.argsort(-ref.astype(np.int64), axis=1, kind="stable")[:, :30]
    got = np.load(out)
    assert np.array_equal(got["index"], order)
    assert np.array_equal(got["scores"], np.take_along_axis(ref, order, axis=1))


def test_merge_top_is_a_stable_descending_merge():
    from biotite_b200.search import _merge_top
    a = (np.array([[9, 7, 7, 1]], dtype=np.int32), np.array([[3, 0, 5, 2]], dtype=np.int32))
    b = (np.array([[9, 7, 2]], dtype=np.int32), np.array([[11, 10, 12]], dtype=np.int32))
    scores, index = _merge_top(a, b, 5)
    assert scores.tolist() == [[9, 9, 7, 7, 7]] and index.tolist() == [[3, 11, 0, 5, 10]]
    assert _merge_top(None, b, 2)[1].tolist() == [[11, 10, 12]]   # the first list is taken as is


# ---- all-vs-all (config 5): the triangle is cut into block tasks, dealt to the ranks by cost ---------

def test_triangle_tasks_cover_the_triangle_once_and_balance_cost():
    from biotite_b200.distributed import triangle_tasks
    rng = np.random.default_rng(4)
    lengths = np.clip(rng.lognormal(np.log(270), 0.7, 3000), 30, 5000).astype(np.int64)
    for world in (1, 2, 3, 8):
        bounds, tasks = triangle_tasks(lengths, world)
        assert bounds[0] == 0 and bounds[-1] == len(lengths) and np.all(np.diff(bounds) >= 0)
        n_blocks = len(bounds) - 1
        seen = sorted(t for per_rank in tasks for t in per_rank)
        assert seen == [(a, b) for a in range(n_blocks) for b in range(a, n_blocks)]
        res = np.array([lengths[bounds[a]:bounds[a + 1]].sum() for a in range(n_blocks)], dtype=np.float64)
        load = np.array([sum(res[a] * res[b] * (0.5 if a == b else 1.0) for a, b in per_rank) for per_rank in tasks])
        assert load.max() <= 1.15 * load.mean()
    # fewer sequences than blocks, and none at all
    bounds, tasks = triangle_tasks(np.array([5, 7]), 4)
    assert bounds.tolist() == [0, 2] and sorted(t for r in tasks for t in r) == [(0, 0)]
    bounds, tasks = triangle_tasks(np.array([5, 7, 6, 9]), 4, blocks=4)
    assert bounds.tolist() == [0, 2, 3, 4] and len([t for r in tasks for t in r]) == 6
    bounds, tasks = triangle_tasks(np.array([], dtype=np.int64), 2)
    assert bounds.tolist() == [0] and all(len(r) == 0 for r in tasks)


def _oracle_block(pool_a, pool_b, gap=(-10, -1), term=True, local=False):
    """Stand-in for all_vs_all_scores (pool_b None) / search_scores: every pair by the oracle."""
    from oracle import oracle
    matrix = bt.SubstitutionMatrix.std_protein_matrix().score_matrix()
    other = pool_a if pool_b is None else pool_b
    out = np.zeros((len(pool_a), len(other)), dtype=np.int32)
    for i in range(len(pool_a)):
        a = pool_a.codes[pool_a.offsets[i]:pool_a.offsets[i + 1]]
        for j in range(len(other)):
            b = other.codes[other.offsets[j]:other.offsets[j + 1]]
            out[i, j] = oracle.align_optimal(a, b, matrix, gap, term, local, True, 1)[1]
    return out


def _triangle_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from biotite_b200.distributed import all_vs_all_scores_sharded
    _, pool = _search_inputs(seed=9, n_db=23)
    res = all_vs_all_scores_sharded(pool, bt.SubstitutionMatrix.std_protein_matrix(), (-10, -1), compute=_oracle_block)
    if rank == 0:
        np.save(out_path, res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_all_vs_all(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "triangle.npy"
    mp.spawn(_triangle_worker, args=(2, _free_port(), str(out)), nprocs=2, join=True)
    _, pool = _search_inputs(seed=9, n_db=23)
    assert np.array_equal(np.load(out), _oracle_block(pool, None))
