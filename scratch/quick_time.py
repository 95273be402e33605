import sys, os, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
import tempfile
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
sol = BatchSolver(net)
print(sol.struct_info())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B in (65536, 262144):
    P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=64)
    P, Q, I_N = sol.prepare(P, Q, I_N)
    for dense in (False, True):
        if dense and B > 65536: continue
        r = sol.solve(P, Q, I_N, dense=dense); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0.record(); r = sol.solve(P, Q, I_N, dense=dense, out=r); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        it = r.n_iter_h.double().mean().item()
        print("B=%d dense=%s  %.3f ms  %.0f solves/s  mean n_iter_h %.2f max %d conv %d" % (
            B, dense, ms, B / ms * 1e3, it, int(r.n_iter_h.max()), int((r.status == 0).sum())))
    if B == 65536:
        a = sol.solve(P, Q, I_N, dense=False).to_host(); b = sol.solve(P, Q, I_N, dense=True).to_host()
        print("structured vs dense: iteration mismatches %d / %d ; median rel V diff %.2e" % (
            (a["n_iter_h"] != b["n_iter_h"]).sum(), B,
            np.median(np.abs(a["V_m"] - b["V_m"]).reshape(-1, B).max(0))))
