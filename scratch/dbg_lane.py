import sys, os, tempfile
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests")); sys.path.insert(0, os.path.join(R, "oracle"))
import numpy as np, torch
import helpers
from harmonic_power_flow_b200 import BatchSolver
net, _, _ = helpers.packed_from_files("net3", 25, True, tempfile.mkdtemp())
B = 333
rng = np.random.default_rng(11)
Vm = rng.uniform(-1.2, 1.2, (net.H, net.n, B))
Va = rng.uniform(-50.0, 50.0, (net.H, net.n, B))
Va[:, :, 100:140] *= 1e4
Va[3, 1, 7] = 105615.0
Va[4, 2, 8] = -105614.99
P = rng.uniform(-2, 2, (net.n, B)); Q = rng.uniform(-2, 2, (net.n, B))
I_N = rng.normal(size=(net.q, net.H, B)) + 1j * rng.normal(size=(net.q, net.H, B))
Vm[1, 1, 5] = np.nan
Va[2, 0, 6] = np.inf
Vm[0, 2, 9] = np.inf
lane = BatchSolver(net)
f1, e1, i1 = lane.mismatch(Vm, Va, P, Q, I_N, want_I_inj=True)
os.environ["HPF_MISMATCH_TILE"] = "1"
tile = BatchSolver(net)
f2, e2, i2 = tile.mismatch(Vm, Va, P, Q, I_N, want_I_inj=True)
f1, e1, i1, f2, e2, i2 = [t.cpu().numpy() for t in (f1, e1, i1, f2, e2, i2)]
print("lane", e1[[5, 6, 9]], "tile", e2[[5, 6, 9]])
print("lane f nan rows lane5", np.isnan(f1[:, 5]).sum(), "tile", np.isnan(f2[:, 5]).sum())
print("lane f nan rows lane6", np.isnan(f1[:, 6]).sum(), "tile", np.isnan(f2[:, 6]).sum())
bad = np.zeros(B, bool); bad[[5, 6, 9]] = True
scale = np.abs(f2[:, ~bad]).max(0)
r = (np.abs(f1[:, ~bad] - f2[:, ~bad]).max(0) / scale)
print("max rel f", r.max(), "argmax", np.argmax(r), "e", np.abs(e1[~bad] - e2[~bad]).max() / np.abs(e2[~bad]).max(),
      "inj", np.abs(i1[:, :, ~bad] - i2[:, :, ~bad]).max() / np.abs(i2[:, :, ~bad]).max())
print("rel in big-angle lanes", r[97:137].max(), "others", np.delete(r, np.s_[97:137]).max())
