"""Sharding the batch of windows over the GPUs of one box.

Windows are independent: every operator, dot product, step length, prox and dual update is
per window (ADMM.py:347-356), so each rank solves a contiguous slice of the batch with NO
collective on the data path.  The only cross-window quantities are the diagnostics — batch-wide
norms and the batch mean of ``x - x_old`` (ADMM.py:612-637) — which every rank produces as
partial sums; they are added up once per solve with a single small all-reduce
(``n_outer * (12 + T*N)`` doubles).  An all-gather of ``x`` is optional.

One process per GPU (torchrun); ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is
only plumbing here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int):
    """Contiguous, balanced slices: the first ``batch % world`` ranks get one extra window."""
    base, extra = divmod(batch, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def reduce_diagnostics(diag, dx_sum, batch, group=None, device=None):
    """Sum the per-rank partial sums (and the window counts) over the group.

    ``diag`` ``(n_outer, 12)`` and ``dx_sum`` ``(n_outer, T, N)`` are float64 numpy arrays of SUMS
    over the local windows.  Returns global ``(diag, dx_sum, batch)``.  With the NCCL backend the
    buffer takes a round trip through ``device``.
    """
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return diag, dx_sum, batch
    flat = torch.from_numpy(np.concatenate([diag.reshape(-1), dx_sum.reshape(-1), [float(batch)]]))
    if dist.get_backend(group) == "nccl":
        buf = flat.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        flat = buf.cpu()
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out = flat.numpy()
    nd, nx = diag.size, dx_sum.size
    return out[:nd].reshape(diag.shape), out[nd:nd + nx].reshape(dx_sum.shape), int(round(out[-1]))


def solve_sharded(blk, y, *, mask=None, group=None, gather=False, print_info=False, y_is_global=True):
    """Run ``blk.combined_loop`` on this rank's slice of the batch.

    ``y_is_global``: ``y`` holds the whole batch on every rank (the slice is taken here);
    otherwise ``y`` is already the local shard.  The residual / regulariser lists of ``blk`` receive
    the GLOBAL diagnostics (identical on every rank); ``alpha_*`` / ``beta_*`` hold the local
    windows' coefficients.  Returns the local ``x`` slice, or the full ``x`` on every rank when
    ``gather``.

    Fixed iteration counts only (``CG_tol <= 0`` and ``ADMM_tol <= 0``): "windows are independent"
    holds for the arithmetic, but the reference's stop tests are batch-global — the CG test is a
    max over the batch (ADMM.py:360) and the outer test uses whole-batch norms (ADMM.py:645) —
    so a shard that tested only its own windows would run different iteration counts than the
    unsharded call.  Tolerance mode is therefore single-device.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world > 1 and (float(blk.CG_tol) > 0 or float(blk.ADMM_tol) > 0):
        raise ValueError("solve_sharded needs fixed iteration counts (CG_tol <= 0 and ADMM_tol <= 0): the reference's "
                         "stop tests are batch-global (ADMM.py:360, 645), so tolerance mode cannot be sharded")
    if y_is_global:
        lo, hi = shard_bounds(y.size(0), world)[rank]
        y = y[lo:hi]
        mask = None if mask is None else mask[lo:hi]
    prev = blk.diag_reduce
    blk.diag_reduce = lambda d, s, b: reduce_diagnostics(d, s, b, group=group, device=blk.device)
    try:
        x = blk.combined_loop(y, mask=mask, print_info=print_info and rank == 0)
    finally:
        blk.diag_reduce = prev
    if not gather or world == 1:
        return x
    # optional final gather (the only data-path collective; NVLink makes 966 MB at B = 65536 a matter of ms).
    # Shards may differ by one window: every rank pads to the largest shard, the pads are dropped afterwards.
    xd = x.to(blk.device).contiguous()
    sizes = torch.zeros(world, dtype=torch.int64, device=xd.device if dist.get_backend(group) == "nccl" else "cpu")
    sizes[rank] = xd.size(0)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    sizes = sizes.tolist()
    top = max(sizes)
    if xd.size(0) < top:
        xd = torch.cat([xd, xd.new_zeros((top - xd.size(0),) + tuple(xd.shape[1:]))])
    full = torch.empty((world * top,) + tuple(xd.shape[1:]), dtype=xd.dtype, device=xd.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(full, xd, group=group)
    else:
        parts = [torch.empty_like(xd) for _ in range(world)]
        dist.all_gather(parts, xd, group=group)
        full = torch.cat(parts)
    if min(sizes) != top:
        full = torch.cat([full[r * top:r * top + n] for r, n in enumerate(sizes)])
    return full.to(x.device)
