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


def gather_last_axis(t: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather a batch-innermost tensor [..., B_local] into [..., B] on every rank.
    Shards are padded to ceil(B/world) so one fixed-size collective serves ragged tails."""
    world = dist.get_world_size(group)
    per = shard_size(B, world)
    lead = t.shape[:-1]
    is_c = t.is_complex()
    x = torch.view_as_real(t) if is_c else t            # NCCL has no complex dtypes
    x = x.movedim(len(lead), 0).contiguous()            # [B_local, ...]
    pad = per - x.shape[0]
    if pad:
        x = torch.cat([x, x.new_zeros((pad,) + tuple(x.shape[1:]))], 0)
    out = x.new_empty((world * per,) + tuple(x.shape[1:]))
    dist.all_gather_into_tensor(out, x, group=group)
    out = out[:B].movedim(0, len(lead))
    if is_c:
        out = torch.view_as_complex(out.contiguous())
    return out.contiguous()


def gather_result(res, B: int, group=None) -> dict:
    """Gather every field of a BatchResult (flags first, then results) -> dict of [.., B] tensors."""
    out = {}
    for k in ("status", "n_iter_f", "n_iter_h", "err_h", "V_m", "V_a", "I_inj"):
        v = getattr(res, k)
        out[k] = None if v is None else gather_last_axis(v, B, group)
    return out
