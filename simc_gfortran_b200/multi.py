"""Multi-GPU end of run: ONE collective over the accumulator blocks, then an exact integer fold.

Events are sharded by try-index range (one process per GPU); the only exchange of the whole path is the
combination of the exact accumulators when a run ends.  Sums of 128-bit fixed-point numbers and min / max of range
keys do not fit one element-wise NCCL reduction, so the ranks all-gather their blocks (a few KB each) and every rank
folds them itself:

* on the GPU (`allreduce_device`): `torch.distributed.all_gather_into_tensor` (NCCL over NVLink) straight out of
  the handle's device accumulators, on the handle's stream, followed by the library's `k_reduce_gathered` kernel
  (`simc_b200_reduce_gathered`); `Simc.fetch` then returns the total on every rank;
* on the host (`allreduce_accum`, used with gloo in the CPU tests and for accumulators already fetched): all-gather of
  the raw `simc_accum` bytes and `simc_b200_accum_merge`.

`run_until_successes` is the `ngen > 0` rule of the reference ("stop at the try that yields the ngen-th success",
simc.f:346-350) over ranks: rounds of one chunk per rank, a gather of the per-chunk success counts after each round,
and a bisection inside the chunk that holds the ngen-th success -- every try is reproducible from (seed, try index),
so the result is the single-process one whatever the number of ranks.
"""
from __future__ import annotations

import ctypes as C

from .lib import Accum, accum_merge


def _group_device():
    """Where the default group's collectives take their tensors: the current CUDA device under NCCL, the host otherwise."""
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _gather_bytes(buf: bytes):
    """all_gather of equal-length byte strings."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = _group_device()
    mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(dev)
    out = torch.empty(world * len(buf), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine)
    raw = out.cpu().numpy().tobytes()
    return [raw[r * len(buf):(r + 1) * len(buf)] for r in range(world)]


def allreduce_accum(acc: Accum, device=None) -> Accum:
    """Total of every rank's HOST accumulators: one all-gather of the raw records + the exact host fold.
    `device` is accepted for compatibility and ignored (host tensors travel through the default group)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return Accum.from_buffer_copy(bytes(acc))
    parts = _gather_bytes(bytes(acc))
    total = Accum.from_buffer_copy(parts[0])
    for raw in parts[1:]:
        accum_merge(total, Accum.from_buffer_copy(raw))
    return total


class _DeviceWords:
    """A device pointer dressed as a CUDA array (int64 words) so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3, "strides": None}


def allreduce_device(sim, scratch: dict | None = None):
    """Folds the DEVICE accumulators of all ranks into every rank's own block, asynchronously on the handle's stream:
    one NCCL all-gather + the library's reduce kernel.  Call `sim.fetch(acc)` afterwards for the total."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return
    ptr, n = sim.device_accum()
    dev = torch.device("cuda", torch.cuda.current_device())
    local = torch.as_tensor(_DeviceWords(ptr, n), device=dev)
    scratch = scratch if scratch is not None else sim.__dict__.setdefault("_multi_scratch", {})
    g = scratch.get("gathered")
    if g is None or g.numel() != world * n:
        g = scratch["gathered"] = torch.empty(world * n, dtype=torch.int64, device=dev)
    with torch.cuda.stream(torch.cuda.ExternalStream(sim.stream)):
        dist.all_gather_into_tensor(g, local)
        sim.reduce_gathered(g.data_ptr(), world)


def _gather_ints(vals):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [list(vals)]
    dev = _group_device()
    mine = torch.tensor(list(vals), dtype=torch.int64, device=dev)
    out = torch.empty(world * len(vals), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine)
    return out.cpu().view(world, len(vals)).tolist()


def run_until_successes(run_range, new_accum, ngen: int, chunk: int):
    """The reference's `ngen > 0` loop over ranks.  run_range(first_try, n) -> Accum of exactly those tries (host
    accumulators; e.g. lambda f, n: sim.run(f, n, seed, sim.accum_clear())); new_accum() -> an empty Accum.
    Returns (total over all ranks, number of tries = index of the try that gave the ngen-th success + 1); every rank
    returns the same."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    assert ngen > 0 and chunk > 0
    mine = new_accum()
    done = 0                                     # successes of all finished rounds, all ranks
    k = 0
    while True:
        first = (k * world + rank) * chunk
        a = run_range(first, chunk)
        counts = [c[0] for c in _gather_ints([a.nsuccess])]
        if done + sum(counts) < ngen:
            accum_merge(mine, a)
            done += sum(counts)
            k += 1
            continue
        # the ngen-th success sits in the chunk of the first rank whose running total reaches ngen
        before, holder = done, 0
        for r, c in enumerate(counts):
            if before + c >= ngen:
                holder = r
                break
            before += c
        need = ngen - before
        cut = 0
        if rank < holder:
            accum_merge(mine, a)
        elif rank == holder:
            lo, hi, best = 0, chunk, a                      # smallest n with nsuccess(first, n) >= need
            while hi - lo > 1:
                mid = (lo + hi) // 2
                t = run_range(first, mid)
                if t.nsuccess >= need:
                    hi, best = mid, t
                else:
                    lo = mid
            accum_merge(mine, best)
            cut = first + hi
        cut = max(c[0] for c in _gather_ints([cut]))
        return allreduce_accum(mine), cut
