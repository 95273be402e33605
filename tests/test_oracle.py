"""The CPU oracle (oracle/hpf_oracle.py) against the fixtures generated from the
reference itself (oracle/make_golden.py -> tests/golden/).  CPU only."""
import glob
import os

import numpy as np
import pytest

import hpf_oracle as O
from conftest import GOLDEN

CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))
SETS = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "scen_*.npz")))

# SURVEY 4.3 known answers (reference executed in the build container)
KNOWN = {
    "net2_uc_h51": (2, 13, 1.825408570965692e-07),
    "net2_c_h51": (2, 21, 2.419218e-06),
    "net2_c_h19": (2, 13, 9.251213128891322e-10),
    "net3_c_h25": (2, 11, 1.880007705068465e-09),
    "net3_uc_h25": (2, 13, 1.825408570965692e-07),
    "net1_c_h51": (3, 23, 4.614292281601706e-11),
    "net1_uc_h51": (3, 14, 3.683057e-06),
    "net1_c_h25": (3, 18, 1.773201e-07),
}


def _phasor(Vm, Va):
    return Vm * np.exp(1j * Va)


def test_fixture_set_complete():
    assert set(KNOWN) <= set(CASES)
    assert {"net3_c_h25_tight", "net3_c_h25_wide", "net2ev_c_h19_tight"} <= set(SETS)


@pytest.mark.parametrize("name", sorted(KNOWN))
def test_golden_matches_survey_known_answers(name):
    d = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    nf, nh, err = KNOWN[name]
    assert int(d["n_iter_f"]) == nf and int(d["n_iter_h"]) == nh
    assert float(d["err_h"]) == pytest.approx(err, rel=1e-6)


def test_golden_spot_values_net3():
    # SURVEY 4.3 spot values, bus idx 3, (V_m, V_a) at h = 1, 3, 5
    d = np.load(os.path.join(GOLDEN, "case_net3_c_h25.npz"))
    want = [(1.080476936943083, 6.276070773595553), (1.776585711850423e-01, 4.355151820201283),
            (2.823989096210703e-01, 1.899824023109353)]
    for h, (vm, va) in enumerate(want):
        assert d["V_m"][h, 3] == pytest.approx(vm, rel=1e-12)
        assert d["V_a"][h, 3] == pytest.approx(va, rel=1e-12)
    assert d["V_m"].sum() == pytest.approx(8.036212889972195, rel=1e-12)
    assert d["THD"][3, 0] == pytest.approx(5.850602937069314e-01, rel=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_oracle_nominal_case(name):
    d = np.load(os.path.join(GOLDEN, "case_%s.npz" % name))
    net = O.net_from_golden(GOLDEN, str(d["net"]), int(d["h_max"]), bool(d["coupled"]))
    assert (net.n, net.m, net.c) == (int(d["n"]), int(d["m"]), int(d["c"]))
    # data contract: p.u. Norton equivalents identical to what HG:301-308 produced
    assert np.array_equal(net.I_N, d["I_N"]) and np.array_equal(net.Y_N, d["Y_N"])
    assert np.array_equal(net.P, d["P"]) and np.array_equal(net.Q, d["Q"])
    Y = O.build_admittance_matrices(net)
    assert np.abs(Y - d["Y_all"]).max() <= 1e-13 * np.abs(d["Y_all"]).max()
    o = O.hpf(net, Y=Y, trace=True)
    assert o["n_iter_f"] == int(d["n_iter_f"])
    assert o["n_iter_h"] == int(d["n_iter_h"])
    Vf = _phasor(*o["V_fund"])
    assert np.abs(Vf - _phasor(d["V_fund_m"], d["V_fund_a"])).max() < 1e-12
    f0 = o["trace"]["f"][0]
    assert np.abs(f0 - d["f0"]).max() <= 1e-13 * np.abs(d["f0"]).max()
    J0 = o["trace"]["J"][0]
    if "J0" in d:
        J0g = d["J0"]
    else:
        J0g = np.zeros_like(J0)
        J0g[d["J0_rows"], d["J0_cols"]] = d["J0_vals"]
    assert J0.shape == (net.N, net.N)
    assert np.abs(J0 - J0g).max() <= 1e-13 * np.abs(J0g).max()
    # one Newton step: limited by cond(J), not by the restatement
    assert np.abs(o["trace"]["x"][0] - d["x1"]).max() <= 1e-8 * np.abs(d["x1"]).max()
    # converged result: within the reference's own SuperLU-vs-LAPACK noise floor
    V, Vg = _phasor(o["V_m"], o["V_a"]), _phasor(d["V_m"], d["V_a"])
    assert np.abs(V - Vg).max() < 1e-8
    assert np.abs(o["THD"] - d["THD"]).max() < 1e-8
    assert o["status"] == 0


@pytest.mark.parametrize("name", SETS)
def test_oracle_scenario_sets(name):
    d = np.load(os.path.join(GOLDEN, "scen_%s.npz" % name))
    net = O.net_from_golden(GOLDEN, str(d["net"]), int(d["h_max"]), bool(d["coupled"]))
    Y = O.build_admittance_matrices(net)
    S = min(len(d["seed"]), 24)
    mism, rel = 0, []
    for s in range(S):
        o = O.hpf(net, P=d["P"][s], Q=d["Q"][s], I_N=d["I_N"][s], Y=Y)
        assert o["n_iter_f"] == int(d["n_iter_f"][s])
        mism += o["n_iter_h"] != int(d["n_iter_h"][s])
        V, Vg = _phasor(o["V_m"], o["V_a"]), _phasor(d["V_m"][s], d["V_a"][s])
        rel.append((np.abs(V - Vg) / np.abs(Vg)).max())
    rel = np.array(rel)
    # the reference disagrees with ITSELF (SuperLU vs LAPACK step) at this level
    Vl = _phasor(d["V_m_lapack"][:S], d["V_a_lapack"][:S])
    Vg = _phasor(d["V_m"][:S], d["V_a"][:S])
    floor = (np.abs(Vl - Vg) / np.abs(Vg)).reshape(S, -1).max(axis=1)
    floor_mism = int((d["n_iter_h"][:S] != d["n_iter_h_lapack"][:S]).sum())
    assert mism <= floor_mism + 1
    assert (rel > 1e-9).sum() <= (floor > 1e-9).sum() + 2
    assert np.median(rel) < 1e-11


# ---------------------------------------------------------------- Norton-equivalent extraction (next-1)
@pytest.mark.parametrize("name", ["smps", "circuit_sim"])
def test_ne_oracle_matches_reference_script(name):
    """oracle/ne_oracle.py against the outputs of the reference's own NE_from_sim.py on the
    simulation files that ship with the reference (fixtures from oracle/make_golden_ne.py)."""
    import ne_oracle as NO
    d = np.load(os.path.join(GOLDEN, "ne_extract_%s.npz" % name))
    assert "passed consistency test" in str(d["ref_stdout"])
    Yuc, Iuc = NO.uncoupled(d["Vf"], d["Vh"], d["I_f"], d["I_h"])
    Yc, Ic = NO.coupled(d["Vf"], d["Vh"], d["I_f"], d["I_h"])
    for got, want in ((Yuc, d["Y_N_uc"]), (Iuc, d["I_N_uc"]), (Yc, d["Y_N_c"]), (Ic, d["I_N_c"])):
        assert np.abs(got - want).max() <= 1e-14 * np.abs(want).max()
    # the extracted model reproduces every measurement it was built from (NE:176-186)
    N = len(d["freq"])
    for k in range(N - 1):
        v = np.zeros(N, dtype=complex); v[0] = d["Vf"][0]; v[k + 1] = d["Vh"][0][k]
        assert np.abs(NO.forward(Yc, Ic, v) - d["I_h"][0][k]).max() <= 1e-9 * np.abs(d["I_h"][0][k]).max()


# ---------------------------------------------------------------- transformer branches in Y(h) (next-3)
@pytest.mark.parametrize("tag", ["a", "b"])
def test_trafo_ybus_oracle_matches_reference_function(tag):
    """oracle build_admittance_matrices_trafo against the outputs of the reference's own
    FPF/pi_trafo_pf_test.py:build_admittance_matrices (fixtures from oracle/make_golden_trafo.py)."""
    d = np.load(os.path.join(GOLDEN, "ybus_trafo.npz"))
    n = len(d[tag + "_X_sh"])
    net = O.Net(n=n, m=n, c=1, harmonics=d[tag + "_harmonics"], line_from=d[tag + "_frm"], line_to=d[tag + "_to"],
                R=d[tag + "_R"], X=d[tag + "_X"], G=d[tag + "_G"], B=d[tag + "_B"], X_sh=d[tag + "_X_sh"],
                P=np.zeros(n), Q=np.zeros(n))
    Y = O.build_admittance_matrices_trafo(net, d[tag + "_tau"], d[tag + "_ph"])
    assert np.array_equal(Y, d[tag + "_Y"])
    # tau = 1, phi = 0 and no to-side shunt match reduces to the plain-line model where both apply
    Y1 = O.build_admittance_matrices_trafo(net, np.ones_like(d[tag + "_tau"]), np.zeros_like(d[tag + "_ph"]))
    Y0 = O.build_admittance_matrices(net)
    off = ~np.eye(n, dtype=bool)
    assert np.abs(Y1[:, off] - Y0[:, off]).max() <= 1e-15 * np.abs(Y0).max()
