import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np
import helpers, hpf_oracle as O
kind = sys.argv[1]; n = int(sys.argv[2]); scale = float(sys.argv[3]); hmax = int(sys.argv[4]) if len(sys.argv) > 4 else 25
kw = dict(n=n, load_scale=scale)
if len(sys.argv) > 5: kw["n_pv"] = int(sys.argv[5])
net, st = helpers.synthetic_packed(kind, tempfile.mkdtemp(), h_max=hmax, **kw)
on = helpers.oracle_net(net)
print("n=%d m=%d c=%d H=%d N=%d L=%d" % (net.n, net.m, net.c, net.H, net.N, len(net.R)))
t = time.time(); Y = O.build_admittance_matrices(on); print("Y %.1fs" % (time.time() - t))
t = time.time()
r = O.hpf(on, Y=Y, solver=os.environ.get("SOLVER", "superlu"))
print("hpf %.1fs n_iter_f=%d n_iter_h=%d err_h=%.3e status=%d" % (time.time() - t, r["n_iter_f"], r["n_iter_h"], r["err_h"], r["status"]))
print("err hist", ["%.2e" % e for e in r["err_h_hist"]])
print("Vm fund min/max", r["V_m"][0].min(), r["V_m"][0].max(), "Vm h3 max", r["V_m"][1].max())
