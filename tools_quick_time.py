import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver, scenarios
import tempfile
net, st, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
sol = BatchSolver(net)
for B in (4096, 65536):
    P, Q, I_N = scenarios.make_batch(net, B, "tight")
    P, Q, I_N = sol.prepare(P, Q, I_N)
    r = sol.solve(P, Q, I_N); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = sol.solve(P, Q, I_N); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = r.n_iter_h.double().mean().item()
    print("B=%d  %.2f ms  %.0f solves/s  mean n_iter_h %.2f  conv %d  us/NR-iter(per scenario-slot) %.3f" % (
        B, ms, B / ms * 1e3, it, int((r.status == 0).sum()), ms * 1e3 / (B * it)))
    Vm, Va = r.V_m, r.V_a
    for name, fn in (("mismatch", lambda: sol.mismatch(Vm, Va, P, Q, I_N)),):
        fn(); torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        print("   ", name, "%.3f ms" % e0.elapsed_time(e1))
    if B <= 4096:
        J = sol.jacobian(Vm, Va); torch.cuda.synchronize()
        e0.record(); sol.jacobian(Vm, Va, out=J); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("    jacobian %.3f ms  %.1f GB/s" % (ms, B * sol.N * sol.N * 8 / ms / 1e6))
        f, err = sol.mismatch(Vm, Va, P, Q, I_N)
        sol.lu_solve(J, f); torch.cuda.synchronize()
        e0.record(); sol.lu_solve(J, f); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("    lu_solve %.3f ms  %.2f us per LU slot, %.2f TFLOP/s" % (ms, ms * 1e3 / B, B * (2 / 3 * sol.N ** 3 + 2 * sol.N ** 2) / ms / 1e9))
