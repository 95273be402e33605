"""The oracle port (oracle/hpf_oracle.py) against the REFERENCE ITSELF, executed live
(oracle/ref_exec.py: hcne_generalized.py exec-ed through the in-memory compat patch).

The committed fixtures pin the port to the reference as it ran when they were generated; this
test re-checks it on the box it runs on (another numpy / scipy build could move either side).
CPU only; auto-skipped where the reference checkout is absent (the GPU box)."""
import os

import numpy as np
import pytest

import hpf_oracle as O
import ref_exec as R
from conftest import GOLDEN

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not R.available(), reason="reference checkout not present (HPF_REFERENCE_ROOT)")]


@pytest.fixture(scope="module")
def session():
    import make_golden as MG
    s = MG.session("net3", 25, True)
    yield s, MG
    s.close()


def _phasor(Vm, Va):
    return Vm * np.exp(1j * Va)


def test_port_matches_live_reference_nominal(session):
    """hpf() of the reference (HG:511-560) vs the port on net3 / coupled / H <= 25: identical
    iteration counts, every intermediate of the first iteration and the result to 1e-12."""
    s, _ = session
    ref = s.solve(trace=True)
    net = O.net_from_golden(GOLDEN, "net3", 25, True)
    assert np.abs(O.build_admittance_matrices(net) - s.Y_all()).max() == 0.0
    o = O.hpf(net, trace=True)
    assert (o["n_iter_f"], o["n_iter_h"]) == (ref["n_iter_f"], ref["n_iter_h"]) == (2, 11)
    close = lambda a, b, tol: np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * np.abs(np.asarray(b)).max()
    assert close(o["trace"]["f"][0], ref["f_h"][0], 1e-12)       # (entries cancel to ~1e-10 next to ones of ~300)
    assert close(o["trace"]["J"][0], ref["J_h"][0], 1e-12)
    assert np.allclose(o["err_h_hist"], ref["err_h_hist"], rtol=1e-6)
    V, Vr = _phasor(o["V_m"], o["V_a"]), _phasor(ref["V_m"], ref["V_a"])
    assert (np.abs(V - Vr) / np.abs(Vr)).max() < 1e-12
    assert np.abs(o["THD"] - ref["THD"]).max() < 1e-12
    assert close(o["J_last"], ref["J_last"], 1e-9)               # HG:536-549: the returned J


@pytest.mark.parametrize("dense", [False, True])
def test_port_matches_live_reference_on_seeded_scenarios(session, dense):
    """Eight seeded scenarios of BASELINE config 3 (scenarios.draw_scenario = the generator of the
    golden sets and the bench), SuperLU step and LAPACK step: the port follows the reference's own
    round-off path - identical counts, phasors to 1e-10."""
    from harmonic_power_flow_b200 import scenarios
    s, MG = session
    net = O.net_from_golden(GOLDEN, "net3", 25, True)
    Y = O.build_admittance_matrices(net)
    comps = [s.base_buses.loc[i, "component"] for i in range(s.m, s.n)]
    for seed in range(8):
        P, Q, I_N = scenarios.draw_scenario(seed, net.P, net.Q, net.I_N, "tight")
        Pm, Qm, k = MG.draw_scenario(seed, net.P, net.Q, net.n - net.m, net.H, MG.SPREADS["tight"])
        assert np.array_equal(P, Pm) and np.array_equal(Q, Qm)
        ref = s.solve(P=P, Q=Q, I_N_scale={comp: k[j] for j, comp in enumerate(comps)}, dense=dense)
        assert np.abs(ref["I_N"] - I_N).max() < 1e-15
        o = O.hpf(net, P=P, Q=Q, I_N=I_N, Y=Y, solver="lapack" if dense else "superlu")
        assert (o["n_iter_f"], o["n_iter_h"]) == (ref["n_iter_f"], ref["n_iter_h"]), seed
        V, Vr = _phasor(o["V_m"], o["V_a"]), _phasor(ref["V_m"], ref["V_a"])
        assert (np.abs(V - Vr) / np.abs(Vr)).max() < 1e-10, seed
