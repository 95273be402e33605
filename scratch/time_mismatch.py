"""Time the standalone mismatch kernel (lane vs tile variant) on config 3, B=65536."""
import sys, os, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=64)
P, Q, I_N = sol.prepare(P, Q, I_N)
raw = sol.solve(P, Q, I_N, raw=True, max_iter_h=3, want_I_inj=False)
N = 2 * net.n * net.H - 1 - net.c
dev = P.device
sets = [(raw.V_m.clone(), raw.V_a.clone(), P.clone(), Q.clone(), I_N.clone(),
         (torch.empty((N, B), dtype=torch.float64, device=dev), torch.empty(B, dtype=torch.float64, device=dev)))
        for _ in range(3)]
fns = [(lambda s_=s_: sol.mismatch(s_[0], s_[1], s_[2], s_[3], s_[4], out=s_[5])) for s_ in sets]
for fn in fns: fn()
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(12): fns[r % 3]()
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 12)
print("eager launches: %.2f us per call" % (best * 1e3))
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for fn in fns: fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        for r in range(12): fns[r % 3]()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 12)
by = 16 * net.n * net.H + 8 * N + 16 * net.q * net.H + 16 * (net.m - 1) + 8
print("variant=%s B=%d  %.2f us  %.0f GB/s  (%.1f%% of 6544.7)" % (
    "tile" if os.environ.get("HPF_MISMATCH_TILE") == "1" else "lane", B, best * 1e3, by * B / best / 1e6,
    by * B / best / 1e6 / 65.447))
f, err = sets[0][5]
print("checksum", float(f.abs().sum()), float(err.max()))
