"""Generate tests/golden/ybus_trafo.npz: Y(h) with transformer branches computed by the REFERENCE's
own function - ``build_admittance_matrices`` of "Fundamental Power Flow/pi_trafo_pf_test.py"
(FPF:99-149), whose text is read at run time and exec-ed (nothing is copied; the rest of that
script does not run under current pandas, so only this function is executed) - on (a) the
script's own 3-bus example (FPF:47-58,64-68, its 1 MW / 230 V bases FPF:529-543) and (b) a seeded
6-bus case with parallel branches and several taps / phase shifts.  Run where /root/reference is
mounted:  python oracle/make_golden_trafo.py"""
import os
import re

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(os.environ.get("HPF_REFERENCE_ROOT", "/root/reference"), "Fundamental Power Flow",
                   "pi_trafo_pf_test.py")
src = open(REF).read()
m = re.search(r"^def build_admittance_matrices\(.*?(?=^def init_voltages)", src, re.S | re.M)
g = {"np": np, "pd": pd}
exec(compile(m.group(0), REF, "exec"), g)
ref_build = g["build_admittance_matrices"]


def run(X_sh, frm, to, R, X, G, B, tau, ph, harmonics):
    buses = pd.DataFrame({"X_sh": np.asarray(X_sh, float)})
    lines = pd.DataFrame({"fromID": np.asarray(frm, float), "toID": np.asarray(to, float), "R": R, "X": X,
                          "G": G, "B": B, "tau": tau, "phase_shift": ph})
    Y = ref_build(buses, lines, harmonics)
    return np.asarray(Y.to_numpy(), dtype=complex).reshape(len(harmonics), len(X_sh), len(X_sh))


out = {}
# (a) the script's own example
base_imp = 1 / ((1000000 / 230) / 230)
base_adm = 1 / base_imp
a = dict(X_sh=np.array([0.005, 0, 0]) / base_imp, frm=[1, 2], to=[2, 3],
         R=np.array([0.01425, 0.0642]) / base_imp, X=np.array([0.05828, 0.083]) / base_imp,
         G=np.array([0.00338, 0]) / base_adm, B=np.array([0.0, 0.00001]) / base_adm,
         tau=np.array([0.95, 1.0]), ph=np.array([150.0, 0.0]), harmonics=[1, 3, 5, 7, 9])
# (b) seeded 6-bus case
rng = np.random.default_rng(20260003)
L = 9
frm = np.array([1, 2, 3, 4, 5, 1, 2, 2, 6]); to = np.array([2, 3, 4, 5, 6, 3, 5, 3, 1])
b = dict(X_sh=np.array([3e-5, 0, 0, 2e-4, 0, 0]), frm=frm, to=to, R=rng.uniform(1e-3, 6e-3, L),
         X=rng.uniform(3e-3, 3e-2, L), G=rng.uniform(0, 1e-2, L), B=rng.uniform(0, 1e-3, L),
         tau=rng.uniform(0.9, 1.1, L), ph=rng.choice([0.0, 30.0, -30.0, 150.0], L), harmonics=list(range(1, 26, 2)))
for tag, c in (("a", a), ("b", b)):
    Y = run(**c)
    for k, v in c.items():
        out[tag + "_" + k] = np.asarray(v)
    out[tag + "_Y"] = Y
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ybus_trafo.npz"), **out)
print("wrote ybus_trafo.npz", out["a_Y"].shape, out["b_Y"].shape)
