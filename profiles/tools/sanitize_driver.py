"""Small workloads through every persistent / warp-specialised kernel, meant to run under
compute-sanitizer (memcheck, racecheck, synccheck):

    compute-sanitizer --tool racecheck python profiles/tools/sanitize_driver.py [mode ...]

Batches are tiny and the grid is capped ($HPF_MAX_CTAS=2) so that lanes / CTAs are refilled from
the work queue; one ragged batch (B not a multiple of 32).  Results are checked against an
un-instrumented expectation only for convergence - the parity tests do the numerics."""
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


def run(tag, net, B, env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    sol = BatchSolver(net)
    for k in (env or {}):
        del os.environ[k]
    P, Q, I_N = scenarios.make_batch(net, B, "wide")
    r = sol.solve(P, Q, I_N, **kw).to_host()
    torch.cuda.synchronize()
    sol.close()
    print("%-34s B=%4d converged %d/%d n_iter_h %d..%d" % (tag, B, int((r["status"] == 0).sum()), B,
                                                          r["n_iter_h"].min(), r["n_iter_h"].max()), flush=True)


modes = sys.argv[1:] or ["warp", "tile", "dense", "cta", "blocked", "gstate", "standalone"]
tmp = tempfile.mkdtemp()
net3, _, _ = helpers.packed_from_files("net3", 25, True, tmp)
cap = {"HPF_MAX_CTAS": "2"}
if "warp" in modes:
    run("harm_hw_kernel (net3)", net3, 150, cap, history=True)
    net2, _, _ = helpers.packed_from_files("net2ev", 19, False, tempfile.mkdtemp())
    run("harm_hw_kernel (net2ev, uncoupled)", net2, 77, cap)
if "tile" in modes:
    run("harm_tile_kernel (net3)", net3, 150, dict(cap, HPF_HARM_KERNEL="tile"), history=True)
    run("harm_tile_kernel<DynDims> (net3)", net3, 45, dict(cap, HPF_NO_SPECIALISE="1"))
if "dense" in modes:
    run("solve_kernel<0> dense smem LU", net3, 21, cap, dense=True, history=True)
if "cta" in modes or "blocked" in modes:
    net1, _, _ = helpers.packed_from_files("net1", 25, True, tempfile.mkdtemp(), julia_schema=True)
    if "cta" in modes:
        run("harm_cta_kernel variant 2 (net1)", net1, 5, cap)
    if "blocked" in modes:
        run("solve_kernel<1> blocked LU (net1)", net1, 3, cap, dense=True)
if "gstate" in modes:
    syn, _ = helpers.synthetic_packed("radial", tempfile.mkdtemp(), h_max=25, n=40, load_scale=0.02)
    run("harm_cta_kernel variant 3 (gstate)", syn, 7, dict(cap, HPF_STRUCT_VARIANT="3"))
    syn2, _ = helpers.synthetic_packed("meshed", tempfile.mkdtemp(), h_max=25, n=70, load_scale=0.02)
    run("variant 3 + multi-CTA operator setup", syn2, 4, dict(cap, HPF_STRUCT_VARIANT="3"))
if "standalone" in modes:
    sol = BatchSolver(net3)
    P, Q, I_N = scenarios.make_batch(net3, 70, "tight")
    raw = sol.solve(P, Q, I_N, raw=True, max_iter_h=3)
    f, err = sol.mismatch(raw.V_m, raw.V_a, P, Q, I_N)
    J = sol.jacobian(raw.V_m, raw.V_a)
    dx, info = sol.lu_solve(J, f)
    d2 = sol.newton_step(raw.V_m, raw.V_a, P, Q, I_N)
    thd = sol.thd(raw.V_m)
    torch.cuda.synchronize()
    print("standalone kernels: |dx - dx_struct| max %.2e" % float((dx - d2).abs().max()), flush=True)
    sol.close()
