"""Multi-process check on real GPUs (torchrun, NCCL): align_batch_sharded, search_scores_sharded and
all_vs_all_scores_sharded against the same work done by rank 0 alone.  Prints one line on rank 0; exit code 1 on mismatch."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
import biotite_b200 as bt  # noqa: E402
from biotite_b200 import _cabi  # noqa: E402
from biotite_b200.batch import PackedSequences  # noqa: E402
from biotite_b200.distributed import align_batch_sharded, all_vs_all_scores_sharded  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
_cabi.check(_cabi.load().bt_set_device(local))
codes1, off1, codes2, off2 = bench.make_c2(40_000, 3)
matrix = bt.SubstitutionMatrix.std_protein_matrix()
res = align_batch_sharded(codes1, off1, codes2, off2, matrix.score_matrix(), (-10, -1), True, False)
ok = True
if rank == 0:
    ref = bt.align_batch_arrays(codes1, off1, codes2, off2, matrix.score_matrix(), (-10, -1), True, False, compact=True)
    ok &= bool(np.array_equal(res.scores, ref.scores) and np.array_equal(res.trace_len, ref.trace_len))
    a, _ = res.traces_flat()
    b, _ = ref.traces_flat()
    ok &= bool(np.array_equal(a, b))
rng = np.random.default_rng(1)


def pool(count, lo, hi):
    lengths = rng.integers(lo, hi, count)
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    return PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)


queries, database = pool(64, 200, 400), pool(6000, 30, 900)
scores = bt.search_scores_sharded(queries, database, matrix, (-10, -1), chunk_size=2000)
if rank == 0:
    ref = bt.search_scores(queries, database, matrix, (-10, -1), chunk_size=2000)
    ok &= bool(np.array_equal(scores, ref))
    # the same database over the node's GPUs from this one process
    multi = bt.search_scores(queries, database, matrix, (-10, -1), chunk_size=2000, devices=list(range(world)))
    ok &= bool(np.array_equal(multi, ref))
# per-query top-k over the sharded database against the top-k of the full matrix
top = bt.search_scores_sharded(queries, database, matrix, (-10, -1), chunk_size=2000, top_k=25)
if rank == 0:
    order = np.argsort(-ref.astype(np.int64), axis=1, kind="stable")[:, :25]
    ok &= bool(np.array_equal(top[1], order) and np.array_equal(top[0], np.take_along_axis(ref, order, axis=1)))
# all-vs-all (config 5): block tasks of the triangle over the ranks against the one-GPU triangle
import time  # noqa: E402
for n_seq in (6000, 12000):
    lengths = np.clip(np.rint(rng.lognormal(np.log(270), 0.7, n_seq)), 30, 5000).astype(np.int64)
    off = np.concatenate(([0], np.cumsum(lengths))).astype(np.int64)
    proteins = PackedSequences(bench.draw_residues(rng, int(off[-1])), off, bt.ProteinSequence.alphabet)
    wall, phases = {}, {}
    for attempt in ("cold", "warm"):
        dist.barrier()
        t0 = time.perf_counter()
        tri = all_vs_all_scores_sharded(proteins, matrix, (-10, -1), terminal_penalty=False, timings=phases)
        dist.barrier()
        wall[attempt] = time.perf_counter() - t0
    if rank == 0:
        bt.all_vs_all_scores(proteins, matrix, (-10, -1), terminal_penalty=False)
        t0 = time.perf_counter()
        ref = bt.all_vs_all_scores(proteins, matrix, (-10, -1), terminal_penalty=False)
        single = time.perf_counter() - t0
        ok &= bool(np.array_equal(tri, ref))
        print(f"all-vs-all of {n_seq} proteins: {world} ranks {wall['warm']:.3f} s (cold {wall['cold']:.3f} s; rank 0: "
              f"compute {phases['compute_s']:.3f} gather {phases['gather_s']:.3f} assemble {phases['assemble_s']:.3f} s, "
              f"{phases['tasks']} tasks), rank 0 alone {single:.3f} s", flush=True)
    del tri
if rank == 0:
    print(f"sharded check on {world} ranks: {'ok' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
