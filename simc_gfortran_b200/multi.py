"""Multi-GPU reduction of the loop's accumulators: ONE NCCL all-reduce of integers.

Events are sharded by try-index range (one process per GPU); the only exchange of the whole path
is the sum of the exact accumulators at the end of a run.  128-bit sums are split
into 32-bit limbs held in int64 so that the element-wise NCCL sum cannot lose a carry; min/max
ranges are reduced as order-preserving integer keys.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import Accum, Fixed128

_MASK32 = (1 << 32) - 1


def _fixed_list(acc: Accum):
    out = [acc.wtcontribute, acc.sum_sigcc]
    out += [acc.sumerr[i] for i in range(8)] + [acc.sumerr2[i] for i in range(8)]
    out += [acc.hist_w[k][b] for k in range(6) for b in range(50)]
    return out


def _key(d: float) -> int:
    i = int(np.float64(d).view(np.int64))
    return i if i >= 0 else i ^ 0x7FFFFFFFFFFFFFFF


def _unkey(k: int) -> float:
    i = k if k >= 0 else k ^ 0x7FFFFFFFFFFFFFFF
    return float(np.int64(i).view(np.float64))


def pack(acc: Accum):
    """-> (sums int64[], mins int64[], maxs int64[])"""
    sums = [acc.ntried, acc.nsuccess, acc.ncontribute, acc.npasscuts, acc.ncontribute_no_rad_proton]
    for f in _fixed_list(acc):
        v = ((int(f.hi) << 64) + int(f.lo)) & ((1 << 128) - 1)      # two's complement, 128 bit
        sums += [(v >> (32 * j)) & _MASK32 for j in range(4)]
    sums += list(np.ctypeslib.as_array(acc.hist_n).ravel())
    sums += list(np.ctypeslib.as_array(acc.stop).ravel())
    sums += list(np.ctypeslib.as_array(acc.transp_calls).ravel())
    sums.append(acc.unsupported)
    mins = [_key(acc.contrib[i].lo) for i in range(32)] + [_key(acc.slop[i].lo) for i in range(8)]
    maxs = [_key(acc.contrib[i].hi) for i in range(32)] + [_key(acc.slop[i].hi) for i in range(8)]
    return (np.array(sums, dtype=np.int64), np.array(mins, dtype=np.int64), np.array(maxs, dtype=np.int64))


def unpack(acc: Accum, sums, mins, maxs) -> Accum:
    out = Accum.from_buffer_copy(bytes(acc))
    it = iter(int(x) for x in sums)
    out.ntried, out.nsuccess, out.ncontribute, out.npasscuts, out.ncontribute_no_rad_proton = (next(it) for _ in range(5))
    for f in _fixed_list(out):
        v = sum(next(it) << (32 * j) for j in range(4)) & ((1 << 128) - 1)   # limb sums carry here
        if v >= 1 << 127:
            v -= 1 << 128
        f.lo = v & ((1 << 64) - 1)
        f.hi = v >> 64
    for name in ("hist_n", "stop", "transp_calls"):
        arr = np.ctypeslib.as_array(getattr(out, name))
        flat = np.array([next(it) for _ in range(arr.size)], dtype=np.int64).reshape(arr.shape)
        arr[...] = flat
    out.unsupported = next(it)
    for i in range(32):
        out.contrib[i].lo, out.contrib[i].hi = _unkey(int(mins[i])), _unkey(int(maxs[i]))
    for i in range(8):
        out.slop[i].lo, out.slop[i].hi = _unkey(int(mins[32 + i])), _unkey(int(maxs[32 + i]))
    return out


def allreduce_accum(acc: Accum, device=None) -> Accum:
    """Sum/min/max of every rank's accumulators (torch.distributed: NCCL on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    sums, mins, maxs = pack(acc)
    dev = device if device is not None else "cpu"
    ts, tmin, tmax = (torch.from_numpy(x).to(dev) for x in (sums, mins, maxs))
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    return unpack(acc, ts.cpu().numpy(), tmin.cpu().numpy(), tmax.cpu().numpy())
