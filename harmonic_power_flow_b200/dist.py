"""Multi-GPU: scenarios shard across ranks with NO data-path collective; one gather of
flags and results at the end.  One process per GPU (torchrun), NCCL backend on
the GPU box, gloo in the CPU tests of the host-side logic."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_size(B: int, world: int) -> int:
    return (B + world - 1) // world


def shard_bounds(B: int, world: int, rank: int):
    """Contiguous block of scenario indices of `rank`: scenario s lives on rank s // ceil(B/world)
    so that concatenating the ranks' outputs in rank order is the identity permutation."""
    per = shard_size(B, world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def shard_inputs(rank, world, *arrays):
    """Slice batch-innermost arrays (last axis = scenario) to this rank's block."""
    B = arrays[0].shape[-1]
    lo, hi = shard_bounds(B, world, rank)
    return [None if a is None else a[..., lo:hi] for a in arrays]


def gather_rank_major(t: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather a batch-innermost tensor [..., B_local] into a RANK-MAJOR stack
    [world, ..., ceil(B/world)] with one collective and no re-layout copy (the tail of the
    last shard is padding when world does not divide B).  Scenario s is element
    [s // per, ..., s % per]; ``unshard`` gives the [..., B] view-copy when it is wanted."""
    world = dist.get_world_size(group)
    per = shard_size(B, world)
    is_c = t.is_complex()
    x = torch.view_as_real(t) if is_c else t            # NCCL has no complex dtypes
    if is_c:
        x = x.movedim(-1, 0)                            # keep the batch axis innermost: [2, ..., B_local]
    x = x.contiguous()
    if x.shape[-1] != per:
        x = torch.cat([x, x.new_zeros(tuple(x.shape[:-1]) + (per - x.shape[-1],))], -1)
    flat = x.new_empty((world * x.shape[0],) + tuple(x.shape[1:]))     # concatenation along dim 0
    dist.all_gather_into_tensor(flat, x, group=group)
    out = flat.view((world,) + tuple(x.shape))
    if is_c:
        out = torch.view_as_complex(out.movedim(1, -1).contiguous())
    return out


def unshard(stacked: torch.Tensor, B: int) -> torch.Tensor:
    """[world, ..., per] rank-major stack -> [..., B] (copies)."""
    world, per = stacked.shape[0], stacked.shape[-1]
    return stacked.movedim(0, -2).reshape(tuple(stacked.shape[1:-1]) + (world * per,))[..., :B].contiguous()


def gather_last_axis(t: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather a batch-innermost tensor [..., B_local] into [..., B] on every rank."""
    return unshard(gather_rank_major(t, B, group), B)


def gather_result(res, B: int, group=None, rank_major: bool = False) -> dict:
    """The single collective step of the path: gather every field of a BatchResult (flags
    first, then results).  rank_major=True keeps the [world, ..., per] stacks (no copies)."""
    out = {}
    f = gather_rank_major if rank_major else gather_last_axis
    for k in ("status", "n_iter_f", "n_iter_h", "err_h", "V_m", "V_a", "I_inj"):
        v = getattr(res, k)
        out[k] = None if v is None else f(v, B, group)
    return out


def gather_slab(slab: torch.Tensor, group=None) -> torch.Tensor:
    """The final gather as ONE collective: every rank contributes its contiguous result slab
    (solver.alloc_result_slab; equal shard sizes), the result is the rank-major [world, nbytes]
    stack; rank r's fields are ``solver.result_from_slab(out[r], ...)`` views."""
    world = dist.get_world_size(group)
    out = slab.new_empty(world * slab.numel())             # concatenation along dim 0 (gloo and NCCL)
    dist.all_gather_into_tensor(out, slab, group=group)
    return out.view(world, slab.numel())


def gather_slab_to_root(slab: torch.Tensor, flag_offset: int, root: int = 0, out=None, group=None,
                        async_op: bool = False):
    """The final gather as the north star words it: convergence FLAGS first (the tail of the result
    slab from ``flag_offset`` on - err_h, n_iter_f, n_iter_h, status: 20 bytes per scenario), then the
    RESULTS (V_m, V_a, I_inj), each as one gather TO THE ROOT RANK only (NCCL: grouped send/recv -
    rank r sends its block once, nobody but the root receives; an all-gather would deliver
    world x the batch to every rank).  Returns ``(stack, works)``: ``stack`` is the rank-major
    [world, nbytes] uint8 tensor on the root (``solver.result_from_slab(stack[r], ...)`` gives rank
    r's fields as views) and None elsewhere; ``works`` are the two async handles (async_op=True)
    - the caller waits on them before reading ``stack`` or reusing ``slab``."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nbytes = slab.numel()
    stack = None
    if rank == root:
        stack = out if out is not None else slab.new_empty((world, nbytes))
        assert stack.shape == (world, nbytes)
    works = []
    for lo, hi in ((flag_offset, nbytes), (0, flag_offset)):
        if hi <= lo:
            continue
        dst = [stack[r, lo:hi] for r in range(world)] if rank == root else None
        w = dist.gather(slab[lo:hi], dst, dst=root, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return stack, works
