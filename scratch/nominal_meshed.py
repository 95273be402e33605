import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver
net, st = helpers.synthetic_packed("meshed", tempfile.mkdtemp(), h_max=25, n=1000, load_scale=0.002)
sol = BatchSolver(net)
r = sol.solve(net.P[:, None].copy(), net.Q[:, None].copy(), net.I_N[:, :, None].copy()).to_host()
print("GPU nominal meshed-1000: n_iter_f=%d n_iter_h=%d err_h=%.3e status=%d  Vm fund min/max %.16g %.16g  Vm h3 max %.16g" % (
    r["n_iter_f"][0], r["n_iter_h"][0], r["err_h"][0], r["status"][0], r["V_m"][0, :, 0].min(), r["V_m"][0, :, 0].max(), r["V_m"][1, :, 0].max()))
print("oracle (98 min on one core):  n_iter_f=2 n_iter_h=26 err_h=4.085e-06 status=0  Vm fund min/max 1.0 1.1649555668183993 Vm h3 max 0.5791799507662769")
