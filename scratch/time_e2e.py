import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = 65536
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=64)
hP = torch.as_tensor(P).pin_memory().numpy(); hQ = torch.as_tensor(Q).pin_memory().numpy(); hI = torch.as_tensor(I_N).pin_memory().numpy()
for _ in range(3): r = sol.solve_host(hP, hQ, hI)
ts = []
for _ in range(10):
    t = time.perf_counter(); r = sol.solve_host(hP, hQ, hI); ts.append(time.perf_counter() - t)
print("chunks=%s  e2e best %.3f ms  median %.3f ms  -> %.1f M solves/s" % (os.environ.get("HPF_HOST_CHUNKS", "default"), min(ts) * 1e3, sorted(ts)[5] * 1e3, B / sorted(ts)[5] / 1e6))
r = sol.solve_host(hP, hQ, hI, keep=True); keep = r["device"]
ts = []
for _ in range(10):
    t = time.perf_counter(); r = sol.solve_host(hP, hQ, hI, keep=keep); ts.append(time.perf_counter() - t)
print("with keep: e2e best %.3f ms  median %.3f ms" % (min(ts) * 1e3, sorted(ts)[5] * 1e3))
