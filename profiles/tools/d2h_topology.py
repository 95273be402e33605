"""Host-path probe of a multi-GPU box (torchrun, one rank per GPU): D2H copy rate of one result slab
(69.5 MB, pinned host memory) per rank ALONE, with ALL ranks copying at once, and in pairs - shows whether the
end-to-end step of N ranks is bounded by a shared host path (PCIe switch / root complex / NUMA node) rather
than by anything the solver does.  env PROBE_AFFINITY=1: bind each rank to the CPUs next to its GPU first."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
aff = "default"
if os.environ.get("PROBE_AFFINITY"):
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    a = sorted(os.sched_getaffinity(0)); aff = "%d CPUs %d..%d" % (len(a), a[0], a[-1])
torch.cuda.set_device(local)
dist.init_process_group("gloo")
NB = 69468160
d = torch.empty(NB, dtype=torch.uint8, device="cuda")
h = torch.empty(NB, dtype=torch.uint8).pin_memory()
hin = torch.empty(17825792, dtype=torch.uint8).pin_memory()
din = torch.empty(17825792, dtype=torch.uint8, device="cuda")
h.copy_(d, non_blocking=True); torch.cuda.synchronize()


def rate(active, reps=6, both=False):
    """GB/s of this rank's D2H copy (max-time over reps excluded: median) when the ranks in `active` copy together."""
    out = []
    for _ in range(reps):
        dist.barrier()
        if rank in active:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            h.copy_(d, non_blocking=True)
            if both:
                din.copy_(hin, non_blocking=True)
            torch.cuda.synchronize()
            out.append(time.perf_counter() - t0)
    if not out:
        return 0.0
    out.sort()
    return NB / out[len(out) // 2] / 1e9


rows = []
for r in range(world):
    rows.append(("rank %d alone" % r, [r]))
rows.append(("all %d ranks" % world, list(range(world))))
for a_ in range(world):
    for b_ in range(a_ + 1, world):
        rows.append(("pair %d+%d" % (a_, b_), [a_, b_]))
res = []
for name, act in rows:
    v = torch.tensor([rate(act)], dtype=torch.float64)
    allv = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allv, v)
    res.append((name, [float(x) for x in allv]))
affs = [None] * world
dist.all_gather_object(affs, aff)
if rank == 0:
    print("affinity per rank:", affs)
    for name, vals in res:
        act = [v for v in vals if v > 0]
        print("%-14s D2H GB/s per rank: %s   aggregate %.1f" % (name, " ".join("%5.1f" % v if v > 0 else "    -" for v in vals), sum(act)))
dist.destroy_process_group()
