import os, sys, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import torch, torch.distributed as dist, numpy as np
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios, dist as hdist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
sol = BatchSolver(net, local)
B = 65536
P, Q, I_N = scenarios.make_batch(net, B, "tight", seed0=rank * B, exact_prefix=8)
res = sol.solve(*sol.prepare(P, Q, I_N))
for _ in range(3): g = hdist.gather_result(res, B * world, rank_major=True)
torch.cuda.synchronize(); dist.barrier()
for name in ("all",) + ("status", "err_h", "V_m", "V_a", "I_inj"):
    t0 = time.perf_counter()
    for _ in range(10):
        if name == "all": g = hdist.gather_result(res, B * world, rank_major=True)
        else: g = hdist.gather_rank_major(getattr(res, name), B * world)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
    if rank == 0: print("gather %-7s %.3f ms" % (name, t * 1e3), flush=True)
dist.destroy_process_group()
