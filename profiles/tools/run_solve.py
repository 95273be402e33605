"""One configuration-3 solve (net3, coupled, H <= 25) on cuda:0 - the target of the ncu captures and
of quick A/B timings.  usage: run_solve.py [B] [reps]   env: HPF_HARM_KERNEL=tile, HPF_SOLVE=dense,
HPF_SAME=1 (every scenario identical: no lane ever refills alone - isolates the refill cost)."""
import os
import sys
import tempfile

R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios

net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dense = os.environ.get("HPF_SOLVE") == "dense"
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight")
if os.environ.get("HPF_SAME"):
    P, Q, I_N = (np.ascontiguousarray(np.repeat(x[..., :1], B, -1)) for x in (P, Q, I_N))
dP, dQ, dI = sol.prepare(P, Q, I_N)
r = sol.solve(dP, dQ, dI, dense=dense)
torch.cuda.synchronize()
sol.set_profiling(True)
best = [1e9, 1e9]
for _ in range(reps):
    r = sol.solve(dP, dQ, dI, out=r, dense=dense)
    k = sol.last_kernel_ms()
    best = [min(best[0], k[0]), min(best[1], k[1])]
it = r.n_iter_h.double()
print("B=%d fund %.4f ms harm %.4f ms (%.1f M solves/s) mean it %.2f max it %d conv %d checksum %.12e" % (
    B, best[0], best[1], B / (best[0] + best[1]) / 1e3, it.mean().item(), int(it.max().item()),
    int((r.status == 0).sum().item()), float(r.V_m.sum())))
