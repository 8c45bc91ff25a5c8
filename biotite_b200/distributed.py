"""
Multi-GPU use of the batch path with one process per GPU (``torchrun``): the pairs of a batch
are split into contiguous shards of equal DP work, every rank aligns its shard on its own
device, and the only communication is the final gather of scores and compact traces - a
``torch.distributed.gather`` of plain int32 / uint32 tensors (NCCL over NVLink when the group's
backend is NCCL, gloo on the CPU), no pickling.  Pairs are independent (the reference's
``align_optimal_generic`` takes one pair, ``src/rust/sequence/align/pairwise.rs:161``), so there
is no collective on the data path.  Within ONE process use ``align_batch_multi`` instead
(``bt_align_batch_multi``: one host thread per device, results straight into the caller's arrays).
"""

from __future__ import annotations

import os

import numpy as np

from .batch import BatchResult, align_batch_arrays

__all__ = ["shard_bounds", "gather_arrays", "gather_columns", "align_batch_sharded", "triangle_tasks",
           "all_vs_all_scores_sharded", "bind_to_gpu_numa_node"]


def shard_bounds(off1, off2, world_size):
    """
    Split pairs ``0..n`` into `world_size` contiguous ranges whose DP cell counts
    (``len1*len2`` summed) are as equal as a contiguous split allows.
    Returns an int64 array of ``world_size + 1`` boundaries.
    """
    off1 = np.asarray(off1, dtype=np.int64)
    off2 = np.asarray(off2, dtype=np.int64)
    cells = np.diff(off1) * np.diff(off2)
    prefix = np.concatenate(([0], np.cumsum(cells)))
    targets = prefix[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(prefix, targets, side="left")
    bounds = np.concatenate(([0], cuts, [len(cells)])).astype(np.int64)
    return np.maximum.accumulate(bounds)


def gather_arrays(array, dst=0, group=None):
    """
    Gather one 1-D NumPy array per rank on rank `dst` (other ranks get ``None``): the lengths
    first, then the payload padded to the longest one, as tensors on the device the group's
    backend moves (CUDA for NCCL, CPU for gloo).  Returns the list of per-rank arrays.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    array = np.ascontiguousarray(array).reshape(-1)
    dtype = array.dtype
    # torch has no uint32 arithmetic, but a gather only moves bytes: reinterpret as int32
    wire = array.view(np.int32) if dtype == np.uint32 else array
    size = torch.tensor([len(wire)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(t.item()) for t in sizes]
    longest = max(sizes) if sizes else 0
    padded = torch.zeros(max(longest, 1), dtype=torch.from_numpy(wire[:0]).dtype, device=device)
    if len(wire):
        padded[:len(wire)] = torch.from_numpy(wire).to(device)
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    parts = [t[:n].cpu().numpy() for t, n in zip(bucket, sizes)]
    return [p.view(np.uint32) if dtype == np.uint32 else p for p in parts]


def gather_columns(part, dst=0, group=None):
    """
    Gather the column blocks of a row-major matrix: every rank holds ``part`` of shape
    ``(rows, width_r)`` (ideally page-locked, ``_cabi.pinned_empty``); rank `dst` receives the
    ``(rows, sum(width_r))`` matrix, the others ``None``.  The blocks travel as device tensors
    (NVLink under NCCL), are joined on the device and leave it through one page-locked copy.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    nccl = dist.get_backend(group) == "nccl"
    device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    part = np.ascontiguousarray(part, dtype=np.int32)
    rows = part.shape[0]
    width = torch.tensor([part.shape[1]], dtype=torch.int64, device=device)
    widths = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(widths, width, group=group)
    widths = [int(t.item()) for t in widths]
    widest = max(max(widths), 1)
    padded = torch.zeros(rows * widest, dtype=torch.int32, device=device)
    if part.size:
        padded[:part.size] = torch.from_numpy(part.reshape(-1)).to(device, non_blocking=True)
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst, group=group)
    if rank != dst:
        return None
    joined = torch.cat([t[:rows * w].view(rows, w) for t, w in zip(bucket, widths)], dim=1)
    if not nccl:
        return joined.numpy()
    from . import _cabi
    out = _cabi.pinned_empty((rows, sum(widths)), np.int32)
    torch.from_numpy(out).copy_(joined, non_blocking=True)
    torch.cuda.synchronize()
    return out


def align_batch_sharded(codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True,
                        local=False, score_only=False, group=None, compute=None, dst=0):
    """
    Collective call: every rank passes the same batch (or at least its own shard's codes),
    aligns shard ``rank`` and rank `dst` receives the :class:`BatchResult` of the whole
    batch (other ranks get ``None``).  `compute` defaults to the GPU batch entry point in its
    compact result form; the CPU test-suite injects a stand-in with the same result layout.
    """
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    off1 = np.ascontiguousarray(off1, dtype=np.int64)
    off2 = np.ascontiguousarray(off2, dtype=np.int64)
    bounds = shard_bounds(off1, off2, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    l_off1 = off1[lo:hi + 1] - off1[lo]
    l_off2 = off2[lo:hi + 1] - off2[lo]
    l_codes1 = codes1[off1[lo]:off1[hi]]
    l_codes2 = codes2[off2[lo]:off2[hi]]
    if compute is None:
        from . import _cabi
        lib = _cabi.load()
        _cabi.check(lib.bt_set_device(int(os.environ.get("LOCAL_RANK", rank))))

        def compute(*args):
            return align_batch_arrays(*args, compact=True)
    part = compute(l_codes1, l_off1, l_codes2, l_off2, scoring, gap_penalty, terminal_penalty, local, None,
                   score_only)
    scores = gather_arrays(part.scores.astype(np.int32, copy=False), dst, group)
    if score_only:
        return None if rank != dst else BatchResult(np.concatenate(scores), None, None, None, np.zeros(0, dtype=np.int64))
    # compact traces: int32 lengths, first cells and the used 2-bit step words of every shard
    if part.ops2 is None:
        # byte form (injected compute): pack it like the device does, 16 steps per word
        words_per_pair = (part.trace_len.astype(np.int64) + 15) // 16
        ops2_off = np.concatenate(([0], np.cumsum(words_per_pair)[:-1])).astype(np.int64)
        ops2 = np.zeros(int(words_per_pair.sum()), dtype=np.uint32)
        ops_off = part.ops_off
        for k in range(len(part.scores)):
            steps = part.ops[ops_off[k]:ops_off[k] + part.trace_len[k]].astype(np.uint64)
            for w in range(int(words_per_pair[k])):
                chunk = steps[16 * w:16 * w + 16]
                ops2[ops2_off[k] + w] = np.uint32(int(sum(int(v) << (2 * b) for b, v in enumerate(chunk))))
        used = ops2
    else:
        # the shard's words are contiguous per window but windows leave gaps: send pair by pair order
        words_per_pair = (part.trace_len.astype(np.int64) + 15) // 16
        take = np.repeat(part.ops_off, words_per_pair) + (np.arange(int(words_per_pair.sum()))
                                                          - np.repeat(np.cumsum(words_per_pair) - words_per_pair, words_per_pair))
        used = part.ops2[take]
    lens = gather_arrays(part.trace_len.astype(np.int32), dst, group)
    firsts = gather_arrays(np.ascontiguousarray(part.first, dtype=np.int32), dst, group)
    words = gather_arrays(used, dst, group)
    if rank != dst:
        return None
    trace_len = np.concatenate(lens)
    per_pair = (trace_len.astype(np.int64) + 15) // 16
    ops2_off = np.concatenate(([0], np.cumsum(per_pair)[:-1])).astype(np.int64)
    result = BatchResult(np.concatenate(scores), trace_len, np.concatenate(firsts).reshape(-1, 2), None, ops2_off)
    result.ops2 = np.concatenate(words) if len(words) else np.zeros(0, dtype=np.uint32)
    if len(result.ops2) == 0:
        result.ops2 = np.zeros(1, dtype=np.uint32)
    result.ops2_words = len(result.ops2)
    return result


def triangle_tasks(lengths, world_size, blocks=None):
    """
    Cut the upper triangle of an all-vs-all job into block tasks and deal them to the ranks.

    The sequences are split into `blocks` contiguous blocks of (nearly) equal residue counts
    (default ``world_size``: with equal blocks a diagonal task costs half an off-diagonal one and
    the ``W (W + 1) / 2`` tasks deal out evenly, ``W / 2`` units per rank); task ``(a, b)``, ``a <= b``, scores block `a` against
    block `b` (``a == b``: the block's own triangle) and costs about the product of the residue
    counts (half of it on the diagonal).  Tasks go to the ranks longest first, each to the least
    loaded rank (deterministic: every rank computes the same assignment).  Returns
    ``(bounds, tasks)`` with ``bounds`` the block boundaries (sequence indices; empty blocks are dropped) and ``tasks[r]`` the list
    of ``(a, b)`` of rank `r`.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    n_blocks = int(blocks) if blocks else int(world_size)
    n_blocks = max(1, min(n_blocks, max(len(lengths), 1)))
    prefix = np.concatenate(([0], np.cumsum(lengths)))
    cuts = np.searchsorted(prefix, prefix[-1] * np.arange(1, n_blocks) / n_blocks, side="left")
    bounds = np.unique(np.concatenate(([0], cuts, [len(lengths)]))).astype(np.int64)   # no empty blocks
    n_blocks = len(bounds) - 1
    residues = np.diff(prefix[bounds])
    todo = []
    for a in range(n_blocks):
        for b in range(a, n_blocks):
            cost = float(residues[a]) * float(residues[b]) * (0.5 if a == b else 1.0)
            todo.append((-cost, a, b))
    todo.sort()
    load = [0.0] * int(world_size)
    tasks = [[] for _ in range(int(world_size))]
    for neg_cost, a, b in todo:
        r = min(range(int(world_size)), key=lambda k: (load[k], k))
        tasks[r].append((a, b))
        load[r] -= neg_cost
    return bounds, tasks


def all_vs_all_scores_sharded(sequences, matrix, gap_penalty=-10, terminal_penalty=True, local=False,
                              group=None, dst=0, blocks=None, compute=None, timings=None):
    """
    Collective :func:`~biotite_b200.all_vs_all_scores` (one process per GPU): the triangle of the
    guide-tree loop (``multiple.pyx:322-333``) is cut into block tasks (:func:`triangle_tasks`),
    every rank scores its tasks - a block's own triangle as a triangular generated batch, two
    different blocks as a cross batch, all enumerated on the GPU - and rank `dst` receives the
    symmetric ``(N, N)`` int32 matrix (the others ``None``).  No collective on the data path; the
    only exchange is the final gather of the blocks.  Every rank holds all sequences (a few MB).
    `compute(pool_a, pool_b_or_None)` defaults to the GPU entry points; the CPU test-suite injects
    the oracle.  `timings` receives ``device_ms`` and ``cells`` of this rank and the wall clock of its
    three phases (``compute_s``, ``gather_s``, ``assemble_s``).
    """
    import torch.distributed as dist

    from .batch import all_vs_all_scores, pack_sequences, PackedSequences
    from .search import search_scores

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pool = sequences if hasattr(sequences, "offsets") else pack_sequences(sequences)
    n = len(pool)
    bounds, tasks = triangle_tasks(pool.lengths, world, blocks)

    def block(a):
        lo, hi = int(bounds[a]), int(bounds[a + 1])
        return PackedSequences(pool.codes[pool.offsets[lo]:pool.offsets[hi]], pool.offsets[lo:hi + 1] - pool.offsets[lo],
                               pool.alphabet)

    if compute is None:
        from . import _cabi
        _cabi.check(_cabi.load().bt_set_device(int(os.environ.get("LOCAL_RANK", rank))))

        def compute(pa, pb):
            t = {}
            if pb is None:
                out = all_vs_all_scores(pa, matrix, gap_penalty, terminal_penalty, local, timings=t)
            else:
                out = search_scores(pa, pb, matrix, gap_penalty, local, terminal_penalty, timings=t)
            if timings is not None:
                timings["device_ms"] = timings.get("device_ms", 0.0) + t.get("device_ms", 0.0)
                timings["cells"] = timings.get("cells", 0) + t.get("cells", 0)
            return out
    import time

    import torch
    nccl = dist.get_backend(group) == "nccl"
    device = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
    size_of = lambda t: int(bounds[t[0] + 1] - bounds[t[0]]) * int(bounds[t[1] + 1] - bounds[t[1]])   # noqa: E731
    longest = max(1, max(sum(size_of(t) for t in per_rank) for per_rank in tasks))
    t0 = time.perf_counter()
    # every block goes to the device as soon as it is scored (no joined host copy)
    padded = torch.zeros(longest, dtype=torch.int32, device=device)
    at = 0
    for a, b in tasks[rank]:
        pa = block(a)
        res = compute(pa, None) if a == b else compute(pa, block(b))
        res = np.ascontiguousarray(res, dtype=np.int32).reshape(-1)
        if res.size != size_of((a, b)):
            raise ValueError(f"block ({a}, {b}) has {res.size} scores, expected {size_of((a, b))}")
        padded[at:at + res.size].copy_(torch.from_numpy(res), non_blocking=True)
        at += res.size
    t1 = time.perf_counter()
    # the only exchange: the blocks of every rank, as device tensors (NVLink under NCCL)
    bucket = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bucket, dst=dst, group=group)
    t2 = time.perf_counter()
    out = None
    if rank == dst:
        # both triangles are written on the device; the matrix leaves it through one page-locked copy
        joined = torch.zeros((n, n), dtype=torch.int32, device=device)
        for r in range(world):
            at = 0
            for a, b in tasks[r]:
                ra, rb = int(bounds[a + 1] - bounds[a]), int(bounds[b + 1] - bounds[b])
                blk = bucket[r][at:at + ra * rb].view(ra, rb)
                at += ra * rb
                joined[bounds[a]:bounds[a + 1], bounds[b]:bounds[b + 1]] = blk
                if a != b:
                    joined[bounds[b]:bounds[b + 1], bounds[a]:bounds[a + 1]] = blk.t()
        if nccl:
            from . import _cabi
            out = _cabi.pinned_empty((n, n), np.int32)
            torch.from_numpy(out).copy_(joined, non_blocking=True)
            torch.cuda.synchronize()
        else:
            out = joined.numpy()
    if timings is not None:
        timings["compute_s"], timings["gather_s"], timings["assemble_s"] = t1 - t0, t2 - t1, time.perf_counter() - t2
        timings["tasks"] = len(tasks[rank])
    return out


def bind_to_gpu_numa_node(local_rank):
    """
    Restrict this process to the CPUs next to GPU `local_rank` (its PCIe root's NUMA node), so
    that page-locked staging buffers are allocated in, and the planning threads run on, the
    memory the GPU's copy engines reach without crossing the socket interconnect.  One process
    per GPU otherwise shares the first node's memory bandwidth.  Returns the CPU set, or
    ``None`` when the topology cannot be read (nothing is changed then).
    """
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                              str(local_rank)], capture_output=True, text=True, timeout=20).stdout.strip()
        if not bus:
            return None
        # sysfs uses a 4-digit PCI domain, nvidia-smi prints 8 digits
        domain, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{domain[-4:].lower()}:{rest.lower()}/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except (OSError, ValueError, subprocess.SubprocessError):
        return None
