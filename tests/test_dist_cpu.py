"""world_size-2 gloo test of the multi-GPU host logic (sharding + the final gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harmonic_power_flow_b200 import dist as hdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        H, n = 3, 4
        g = torch.Generator().manual_seed(1234)
        full_V = torch.rand((H, n, B), dtype=torch.float64, generator=g)
        full_I = torch.complex(torch.rand((2, H, B), dtype=torch.float64, generator=g),
                               torch.rand((2, H, B), dtype=torch.float64, generator=g))
        full_st = torch.randint(0, 4, (B,), dtype=torch.int32, generator=g)
        V, I, st = hdist.shard_inputs(rank, world, full_V, full_I, full_st)
        lo, hi = hdist.shard_bounds(B, world, rank)
        assert V.shape[-1] == hi - lo
        gV = hdist.gather_last_axis(V.contiguous(), B)
        gI = hdist.gather_last_axis(I.contiguous(), B)
        gs = hdist.gather_last_axis(st.contiguous(), B)
        ok = torch.equal(gV, full_V) and torch.equal(gI, full_I) and torch.equal(gs, full_st)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 37, 1])
def test_gather_is_identity_permutation_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]


def _slab_worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from harmonic_power_flow_b200 import solver
        H, n, nq = 3, 4, 2
        lay, total = solver.result_slab_layout(H, n, nq, B)
        slab = torch.zeros(total, dtype=torch.uint8)
        res = solver.result_from_slab(slab, H, n, nq, B)
        g = torch.Generator().manual_seed(100 + rank)
        res.V_m.copy_(torch.rand((H, n, B), dtype=torch.float64, generator=g))
        res.I_inj.copy_(torch.complex(torch.rand((nq, H, B), dtype=torch.float64, generator=g),
                                      torch.rand((nq, H, B), dtype=torch.float64, generator=g)))
        res.status.copy_(torch.randint(0, 4, (B,), dtype=torch.int32, generator=g))
        out = hdist.gather_slab(slab)
        ok = out.shape == (world, total)
        for r in range(world):                       # every rank's fields come back bit for bit
            gr = torch.Generator().manual_seed(100 + r)
            want_V = torch.rand((H, n, B), dtype=torch.float64, generator=gr)
            want_I = torch.complex(torch.rand((nq, H, B), dtype=torch.float64, generator=gr),
                                   torch.rand((nq, H, B), dtype=torch.float64, generator=gr))
            want_s = torch.randint(0, 4, (B,), dtype=torch.int32, generator=gr)
            rr = solver.result_from_slab(out[r], H, n, nq, B)
            ok = ok and torch.equal(rr.V_m, want_V) and torch.equal(rr.I_inj, want_I) and torch.equal(rr.status, want_s)
        # the gather to the root rank only (flags first, then results): same bytes, one receiver
        flag_off = lay["err_h"][0]
        for async_op in (False, True):
            stack, works = hdist.gather_slab_to_root(slab, flag_off, root=0, async_op=async_op)
            for w in works:
                w.wait()
            ok = ok and ((stack is None) == (rank != 0))
            if rank == 0:
                ok = ok and torch.equal(stack, out)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_result_slab_single_collective_world2():
    """The result fields as views into one contiguous slab: one all_gather returns every rank's
    V, I and flags bit for bit (what the multi-GPU bench uses for its final gather)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(2))
    assert got == [(0, True), (1, True)]
