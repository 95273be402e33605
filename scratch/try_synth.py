import sys, os, time, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers, hpf_oracle as O
from harmonic_power_flow_b200 import BatchSolver, scenarios
kind = sys.argv[1]; n = int(sys.argv[2]); scale = float(sys.argv[3]); hmax = int(sys.argv[4]); B = int(sys.argv[5]); nchk = int(sys.argv[6])
net, st = helpers.synthetic_packed(kind, tempfile.mkdtemp(), h_max=hmax, n=n, load_scale=scale)
print("n=%d m=%d c=%d H=%d N=%d nx=%d nZ=%d" % (net.n, net.m, net.c, net.H, net.N, 2 * net.m - 1 - net.c, net.n * net.H - net.m), flush=True)
sol = BatchSolver(net)
t = time.time(); info = sol.struct_info(); torch.cuda.synchronize(); print("struct_info", info, "%.2fs" % (time.time() - t), flush=True)
P, Q, I_N = scenarios.make_batch(net, B, "tight", exact_prefix=min(B, 64))
dP, dQ, dI = sol.prepare(P, Q, I_N)
t = time.time(); r = sol.solve(dP, dQ, dI); torch.cuda.synchronize(); t1 = time.time() - t
t = time.time(); r = sol.solve(dP, dQ, dI); torch.cuda.synchronize(); t2 = time.time() - t
res = r.to_host()
print("solve %.3fs (first %.3fs)  %.1f solves/s  status counts %s  n_iter_f %s n_iter_h min/mean/max %d/%.1f/%d" % (
    t2, t1, B / t2, np.bincount(res["status"], minlength=4), np.unique(res["n_iter_f"]), res["n_iter_h"].min(), res["n_iter_h"].mean(), res["n_iter_h"].max()), flush=True)
on = helpers.oracle_net(net)
Y = O.build_admittance_matrices(on)
for b in range(nchk):
    t = time.time()
    o = O.hpf(on, P=P[:, b], Q=Q[:, b], I_N=I_N[:, :, b], Y=Y)
    Vo = o["V_m"] * np.exp(1j * o["V_a"]); Vg = res["V_m"][:, :, b] * np.exp(1j * res["V_a"][:, :, b])
    print("scen %d oracle %.1fs it_f %d/%d it_h %d/%d err %.2e/%.2e  max|dV|/max|V| %.2e  I_inj rel %.2e" % (
        b, time.time() - t, o["n_iter_f"], res["n_iter_f"][b], o["n_iter_h"], res["n_iter_h"][b], o["err_h"], res["err_h"][b],
        np.abs(Vo - Vg).max() / np.abs(Vo).max(), np.abs(o["I_inj"] - res["I_inj"][:, :, b]).max() / np.abs(o["I_inj"]).max()), flush=True)
