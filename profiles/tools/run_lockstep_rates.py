"""Iteration-count agreement of the lock-step path, the per-CTA path and the oracle with a LAPACK step
against the oracle with the reference's SuperLU step, on one synthetic network (all scenarios).
usage: run_lockstep_rates.py kind n B"""
import os, sys, tempfile, pathlib
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "oracle")):
    sys.path.insert(0, p)
import numpy as np, torch
import helpers, oracle_pool
from harmonic_power_flow_b200 import BatchSolver, scenarios
kind, n, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
os.environ["HPF_STRUCT_VARIANT"] = "3"
net, _ = helpers.synthetic_packed(kind, pathlib.Path(tempfile.mkdtemp()), h_max=25, n=n, load_scale=0.02)
P, Q, I_N = scenarios.make_batch(net, B, "tight")
os.environ["HPF_LOCKSTEP"] = "0"; cta = BatchSolver(net); rc = cta.solve(P, Q, I_N).to_host(); cta.close()
os.environ["HPF_LOCKSTEP"] = "1"; ls = BatchSolver(net); rl = ls.solve(P, Q, I_N).to_host(); ls.close()
pool = oracle_pool.OraclePool(helpers.oracle_net(net))
o = pool.solve(P, Q, I_N, "superlu"); ol = pool.solve(P, Q, I_N, "lapack"); pool.close()
def rate(a): return int((a != o["n_iter_h"]).sum())
Vo = helpers.phasor(o["V_m"], o["V_a"])
def dv(r):
    V = helpers.phasor(r["V_m"], r["V_a"]); same = r["n_iter_h"] == o["n_iter_h"]
    d = np.abs(V - Vo).max(axis=(0, 1)) / np.abs(Vo).max(axis=(0, 1))
    return float(np.median(d[same])), float(d[same].max())
print("%s n=%d B=%d: iteration-count mismatches vs oracle(SuperLU): lock-step %d, per-CTA %d, oracle(LAPACK) %d; "
      "lock-step vs per-CTA %d; phasor diff (same count) median/max: lock-step %.1e/%.1e per-CTA %.1e/%.1e" % (
      kind, n, B, rate(rl["n_iter_h"]), rate(rc["n_iter_h"]), rate(ol["n_iter_h"]),
      int((rl["n_iter_h"] != rc["n_iter_h"]).sum()), *dv(rl), *dv(rc)))
