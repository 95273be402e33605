"""hpf_solve_host end to end (pinned host buffers, config 3): sweep of chunk plans ($HPF_HOST_PLAN, sizes in
scenarios, the last one repeats) and compute-stream counts ($HPF_HOST_STREAMS) in ONE process (the library
reads both variables on every call).  Every plan must return the bits of the first one."""
import os, sys, tempfile, time
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch, helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
net, _, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
sol = BatchSolver(net)
P, Q, I_N = scenarios.make_batch(net, B, "tight")
hP, hQ, hI = (torch.as_tensor(x).pin_memory().numpy() for x in (P, Q, I_N))
PLANS = [
    "8192",
    "4096,8192,12288,16384",
    "4096,8192,16384",
    "4096,12288,16384",
    "6144,10240,16384",
    "4096,8192,12288,16384,24576",
    "8192,16384",
    "4096,8192,16384,32768",
    "2048,6144,8192,16384",
    "4096,12288,24576",
    "3072,9216,16384",
    "4096,8192,12288",
    "16384",
]
if len(sys.argv) > 2:
    PLANS = sys.argv[2:]
ref = None
rows = []
for streams in (2, 3):
    for plan in PLANS:
        os.environ["HPF_HOST_PLAN"] = plan
        os.environ["HPF_HOST_STREAMS"] = str(streams)
        for _ in range(2):
            r = sol.solve_host(hP, hQ, hI)
        ts = []
        for _ in range(9):
            t0 = time.perf_counter(); r = sol.solve_host(hP, hQ, hI); ts.append(time.perf_counter() - t0)
        ts.sort()
        snap = tuple(np.array(r[k], copy=True) for k in ("V_m", "V_a", "I_inj", "n_iter_h", "err_h", "status"))
        if ref is None:
            ref = snap
        same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, snap))
        rows.append((ts[len(ts) // 2], ts[0], streams, plan, same))
        print("streams=%d plan=%-44s median %.3f ms  best %.3f ms  (%.1f M solves/s)  bits-equal %s" % (
            streams, plan, ts[len(ts) // 2] * 1e3, ts[0] * 1e3, B / ts[len(ts) // 2] / 1e6, same), flush=True)
rows.sort()
print("== best by median:")
for r in rows[:6]:
    print("  %.3f ms (best %.3f)  streams=%d plan=%s bits-equal %s" % (r[0] * 1e3, r[1] * 1e3, r[2], r[3], r[4]))
