"""Generate tests/golden/ne_extract_<name>.npz: inputs (the measurement arrays the product's own
loader reads from the reference's simulation files) and outputs of the REFERENCE's script
``Circuit Simulation/NE_from_sim.py`` executed by ``ref_exec_ne.py``.  Run where /root/reference
is mounted:  python oracle/make_golden_ne.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_exec_ne as R                                    # noqa: E402
from harmonic_power_flow_b200 import ne_from_sim           # noqa: E402

for name in ("smps", "circuit_sim"):
    mat = os.path.join(R.REF_ROOT, "Circuit Simulation", name + ".mat")
    g = R.run(mat)
    sim = ne_from_sim.load_simulation(mat)
    assert list(sim.freq) == list(g["freq"])
    out = os.path.join(ROOT, "tests", "golden", "ne_extract_%s.npz" % name)
    np.savez_compressed(out, freq=sim.freq, Vf=sim.Vf, Vh=sim.Vh, I_f=sim.I_f, I_h=sim.I_h,
                        Y_N_c=np.asarray(g["Y_N_c"], dtype=complex), I_N_c=np.asarray(g["I_N_c"], dtype=complex),
                        Y_N_uc=np.asarray(g["Y_N_uc"], dtype=complex), I_N_uc=np.asarray(g["I_N_uc"], dtype=complex),
                        ref_stdout=g["__stdout__"])
    print("wrote", out, "N =", len(sim.freq))
