"""
Multi-GPU use of the batch path: one process per GPU (``torchrun``), the pairs of a batch
are split into contiguous shards of equal DP work, every rank aligns its shard on its own
device, and the only communication is the final gather of scores and compact traces.
Pairs are independent (the reference's ``align_optimal_generic`` takes one pair,
``src/rust/sequence/align/pairwise.rs:161``), so there is no collective on the data path.
"""

from __future__ import annotations

import os

import numpy as np

from .batch import BatchResult, align_batch_arrays

__all__ = ["shard_bounds", "align_batch_sharded", "bind_to_gpu_numa_node"]


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


def align_batch_sharded(codes1, off1, codes2, off2, scoring, gap_penalty, terminal_penalty=True,
                        local=False, score_only=False, group=None, compute=None, dst=0):
    """
    Collective call: every rank passes the same batch (or at least its own shard's codes),
    aligns shard ``rank`` and rank `dst` receives the :class:`BatchResult` of the whole
    batch (other ranks get ``None``).  `compute` defaults to the GPU batch entry point; the
    CPU test-suite injects the oracle here.
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
        compute = align_batch_arrays
    local_result = compute(l_codes1, l_off1, l_codes2, l_off2, scoring, gap_penalty,
                           terminal_penalty, local, None, score_only)
    payload = (lo, hi, local_result.scores, local_result.trace_len, local_result.first,
               local_result.ops)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    gathered.sort(key=lambda item: item[0])
    scores = np.concatenate([g[2] for g in gathered])
    if score_only:
        return BatchResult(scores, None, None, None, off1 + off2)
    trace_len = np.concatenate([g[3] for g in gathered])
    first = np.concatenate([g[4] for g in gathered])
    ops = np.concatenate([g[5] for g in gathered])
    return BatchResult(scores, trace_len, first, ops, off1 + off2)


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
