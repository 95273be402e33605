"""Generate the committed golden fixtures under tests/golden/ from the REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs the read-only
reference checkout at /root/reference); the GPU box and the test-suite use the
fixtures this script wrote.  Usage::

    python oracle/make_golden.py            # all fixtures (~10 min, 8 processes)
    python oracle/make_golden.py --quick    # skip net1 and use fewer scenarios

What is stored (all float64/complex128, compressed npz + one JSON):

* ``networks.json``  the networks of the reference's CSV files as plain tables in SI
  units (net1 schema-adapted: ``X_shunt`` -> ``X_sh``, ``G = B = 0``; ``net2ev`` =
  net2 with bus 3 turned into a second nonlinear bus).
* ``ne_devices.npz``  Norton-equivalent tables in SI units: ``smps`` parsed from the
  reference's ``Circuit Simulation/smps_NE.csv`` exactly like HG:291-299, and the
  synthetic ``ev`` device (seed 20260001 perturbation).
* ``case_<name>.npz``  nominal solves: Y(h), fundamental result, first
  mismatch / Jacobian / state update, final V, iteration counts, error histories, THD.
* ``scen_<set>.npz``  randomised scenario sets: inputs (P, Q, I_N p.u.) and the
  reference's results with its own SuperLU step AND with a LAPACK step (the
  oracle-vs-oracle noise floor of SURVEY 7.3).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time
from multiprocessing import Pool

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_exec as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
EV_SEED = 20260001


# ---------------------------------------------------------------------------- inputs
def read_ne_csv_like_reference(path):
    """HG:291-299: index (Parameter, Frequency), int columns, complex(val.strip('()'))."""
    df = pd.read_csv(path, index_col=["Parameter", "Frequency"])
    df.columns = df.columns.astype(int)
    df = df.apply(lambda col: col.apply(lambda v: complex(v.strip("()"))))
    freqs = np.array(df.columns, dtype=np.int64)
    return dict(freqs=freqs,
                Y_N_c=df.loc["Y_N_c"].to_numpy(dtype=complex),
                Y_N_c_rowfreq=np.array(df.loc["Y_N_c"].index, dtype=np.int64),
                I_N_c=df.loc["I_N_c"].to_numpy(dtype=complex).ravel(),
                Y_N_uc=df.loc["Y_N_uc"].to_numpy(dtype=complex).ravel(),
                I_N_uc=df.loc["I_N_uc"].to_numpy(dtype=complex).ravel())


def perturb_device(dev, seed):
    """Synthetic second device: every entry x (1 + 0.1 N(0,1)) e^{j 0.1 N(0,1)}."""
    rng = np.random.default_rng(seed)
    out = dict(freqs=dev["freqs"], Y_N_c_rowfreq=dev["Y_N_c_rowfreq"])
    for k in ("Y_N_c", "I_N_c", "Y_N_uc", "I_N_uc"):
        a = dev[k]
        out[k] = a * (1 + 0.1 * rng.standard_normal(a.shape)) * \
            np.exp(1j * 0.1 * rng.standard_normal(a.shape))
    return out


def write_ne_csv(dev, path):
    """Same table the reference's extractor exports (CS/NE_from_sim.py:195-209)."""
    freqs = [int(f) for f in dev["freqs"]]
    idx = pd.MultiIndex.from_arrays(
        [len(freqs) * ["Y_N_c"] + ["I_N_c", "Y_N_uc", "I_N_uc"], freqs + [0, 0, 0]],
        names=["Parameter", "Frequency"])
    NE = pd.DataFrame(np.zeros((len(freqs) + 3, len(freqs))), index=idx, dtype=complex,
                      columns=freqs)
    NE.iloc[:len(freqs), :] = dev["Y_N_c"]
    NE.iloc[len(freqs), :] = dev["I_N_c"]
    NE.iloc[len(freqs) + 1, :] = dev["Y_N_uc"]
    NE.iloc[len(freqs) + 2, :] = dev["I_N_uc"]
    NE.to_csv(path)


def read_net(name):
    b = pd.read_csv(os.path.join(R.NET_DIR, name + "_buses.csv"), delimiter=";")
    l = pd.read_csv(os.path.join(R.NET_DIR, name + "_lines.csv"), delimiter=";")
    if "X_shunt" in b.columns:                      # net1 / Julia schema
        b = b.rename(columns={"X_shunt": "X_sh"})
    for col in ("G", "B"):
        if col not in l.columns:
            l[col] = 0
    b = b[["ID", "type", "component", "S", "P", "Q", "X_sh"]]
    l = l[["ID", "fromID", "toID", "R", "X", "G", "B"]]
    return b, l


def net_tables():
    nets = {}
    for name in ("net1", "net2", "net3"):
        nets[name] = read_net(name)
    b, l = read_net("net2")
    b = b.copy()
    b.loc[2, ["type", "component", "P", "Q"]] = ["nonlinear", "ev", 150, 50]
    nets["net2ev"] = (b, l)
    return nets


def tables_to_json(nets):
    out = {}
    for name, (b, l) in nets.items():
        out[name] = dict(
            buses={c: b[c].tolist() for c in b.columns},
            lines={c: l[c].tolist() for c in l.columns})
    return out


# ---------------------------------------------------------------------------- sessions
_SESS = {}
_WORK = None


def workdir():
    global _WORK
    if _WORK is None:
        _WORK = tempfile.mkdtemp(prefix="hpf_golden_")
        nets = net_tables()
        for name, (b, l) in nets.items():
            b.to_csv(os.path.join(_WORK, name + "_buses.csv"), sep=";", index=False)
            l.to_csv(os.path.join(_WORK, name + "_lines.csv"), sep=";", index=False)
        smps = read_ne_csv_like_reference(os.path.join(R.NE_DIR, "smps_NE.csv"))
        write_ne_csv(perturb_device(smps, EV_SEED), os.path.join(_WORK, "ev_NE.csv"))
    return _WORK


def session(net, h_max, coupled):
    key = (net, h_max, coupled)
    if key not in _SESS:
        w = workdir()
        smps = os.path.join(R.NE_DIR, "smps_NE.csv")
        ne = {"SMPS": smps, "smps": smps, "ev": os.path.join(w, "ev_NE.csv")}
        _SESS[key] = R.RefSession(os.path.join(w, net + "_buses.csv"),
                                  os.path.join(w, net + "_lines.csv"), h_max, coupled, ne)
    return _SESS[key]


# ---------------------------------------------------------------------------- nominal cases
CASES = [
    # name, net, h_max, coupled, store dense J0
    ("net2_uc_h51", "net2", 51, False, True),
    ("net2_c_h51", "net2", 51, True, True),
    ("net2_c_h19", "net2", 19, True, True),
    ("net3_c_h25", "net3", 25, True, True),
    ("net3_uc_h25", "net3", 25, False, True),
    ("net2ev_c_h19", "net2ev", 19, True, True),
    ("net2ev_uc_h19", "net2ev", 19, False, True),
    ("net3_c_h5", "net3", 5, True, True),
    ("net1_c_h25", "net1", 25, True, False),
    ("net1_uc_h51", "net1", 51, False, False),
    ("net1_c_h51", "net1", 51, True, False),
]


def run_case(args):
    name, net, h_max, coupled, store_J = args
    t0 = time.time()
    s = session(net, h_max, coupled)
    o = s.solve(trace=True)
    I_N, Y_N = s.ne_arrays()
    d = dict(net=net, h_max=h_max, coupled=coupled, n=s.n, m=s.m, c=s.c,
             harmonics=np.array(s.harmonics), Y_all=s.Y_all(), I_N=I_N, Y_N=Y_N,
             P=s.base_buses["P"].to_numpy(dtype=float), Q=s.base_buses["Q"].to_numpy(dtype=float),
             V_fund_m=o["V_fund"][:, :, 0], V_fund_a=o["V_fund"][:, :, 1],
             V_m=o["V_m"], V_a=o["V_a"], err_h=o["err_h"], n_iter_h=o["n_iter_h"],
             n_iter_f=o["n_iter_f"], err_f_hist=o["err_f_hist"], err_h_hist=o["err_h_hist"],
             THD=o["THD"], f0=o["f_h"][0], x1=o["x_h"][0], f1=o["f_h"][1])
    if store_J:
        d["J0"] = o["J_h"][0]
        d["J_last"] = o["J_last"]
    else:                       # big nets: keep J0 sparse (COO) to stay small
        J0 = o["J_h"][0]
        r, cidx = np.nonzero(J0)
        d["J0_rows"], d["J0_cols"], d["J0_vals"] = r.astype(np.int32), cidx.astype(np.int32), J0[r, cidx]
    np.savez_compressed(os.path.join(OUT, "case_%s.npz" % name), **d)
    return "%s: n_iter_f=%d n_iter_h=%d err_h=%.6e (%.1fs)" % (
        name, o["n_iter_f"], o["n_iter_h"], o["err_h"], time.time() - t0)


# ---------------------------------------------------------------------------- scenarios
def draw_scenario(seed, P0, Q0, q, H, spread):
    """rng = default_rng(seed); P factors, Q factors, |k|, angle(k)."""
    dp, dm, da = spread
    rng = np.random.default_rng(seed)
    P = P0 * rng.uniform(1 - dp, 1 + dp, len(P0))
    Q = Q0 * rng.uniform(1 - dp, 1 + dp, len(Q0))
    mag = rng.uniform(1 - dm, 1 + dm, (q, H))
    ang = rng.uniform(-da, da, (q, H))
    return P, Q, mag * np.exp(1j * ang)

SPREADS = {"tight": (0.1, 0.05, 0.02), "wide": (0.5, 0.2, 0.1)}

SCEN_SETS = [
    # name, net, h_max, coupled, spread, number of seeds
    ("net3_c_h25_tight", "net3", 25, True, "tight", 64),
    ("net3_c_h25_wide", "net3", 25, True, "wide", 40),
    ("net2ev_c_h19_tight", "net2ev", 19, True, "tight", 32),
    ("net3_uc_h25_tight", "net3", 25, False, "tight", 16),
]


def run_scenario(args):
    net, h_max, coupled, spread, seed = args
    s = session(net, h_max, coupled)
    H = len(s.harmonics)
    q = s.n - s.m
    P0 = s.base_buses["P"].to_numpy(dtype=float)
    Q0 = s.base_buses["Q"].to_numpy(dtype=float)
    P, Q, k = draw_scenario(seed, P0, Q0, q, H, SPREADS[spread])
    comps = [s.base_buses.loc[i, "component"] for i in range(s.m, s.n)]
    scale = {comp: k[j] for j, comp in enumerate(comps)}
    res = {}
    for variant, dense in (("superlu", False), ("lapack", True)):
        o = s.solve(P=P, Q=Q, I_N_scale=scale, dense=dense)
        res[variant] = o
    o = res["superlu"]
    return dict(seed=seed, P=P, Q=Q, I_N=o["I_N"],
                V_m=o["V_m"], V_a=o["V_a"], n_iter_f=o["n_iter_f"], n_iter_h=o["n_iter_h"],
                err_h=o["err_h"],
                V_m_lapack=res["lapack"]["V_m"], V_a_lapack=res["lapack"]["V_a"],
                n_iter_h_lapack=res["lapack"]["n_iter_h"], err_h_lapack=res["lapack"]["err_h"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"

    nets = net_tables()
    with open(os.path.join(OUT, "networks.json"), "w") as f:
        json.dump(tables_to_json(nets), f, indent=1)
    smps = read_ne_csv_like_reference(os.path.join(R.NE_DIR, "smps_NE.csv"))
    ev = perturb_device(smps, EV_SEED)
    dev = {}
    for name, d in (("smps", smps), ("ev", ev)):
        for k, v in d.items():
            dev[name + "__" + k] = v
    np.savez_compressed(os.path.join(OUT, "ne_devices.npz"), **dev)

    cases = [c for c in CASES if not (a.quick and c[1] == "net1")]
    sets = SCEN_SETS
    with Pool(a.procs) as pool:
        case_async = pool.map_async(run_case, cases, chunksize=1)
        scen_async = []
        for name, net, h_max, coupled, spread, count in sets:
            if a.quick:
                count = min(count, 8)
            jobs = [(net, h_max, coupled, spread, s) for s in range(count)]
            scen_async.append((name, net, h_max, coupled, spread,
                               pool.map_async(run_scenario, jobs, chunksize=2)))
        for line in case_async.get():
            print(line, flush=True)
        for name, net, h_max, coupled, spread, r in scen_async:
            rows = r.get()
            d = {k: np.array([row[k] for row in rows]) for k in rows[0]}
            d.update(net=net, h_max=h_max, coupled=coupled, spread=np.array(SPREADS[spread]))
            np.savez_compressed(os.path.join(OUT, "scen_%s.npz" % name), **d)
            mism = int((d["n_iter_h"] != d["n_iter_h_lapack"]).sum())
            print("%s: %d scenarios, n_iter_h hist %s, superlu-vs-lapack iteration mismatches %d" % (
                name, len(rows), dict(zip(*np.unique(d["n_iter_h"], return_counts=True))), mism),
                flush=True)

    import scipy
    manifest = dict(generated_by="oracle/make_golden.py", reference="pweigmann/harmonic-power-flow "
                    "Harmonic Power Flow/hcne_generalized.py (exec-ed unmodified except the compat "
                    "patch described in oracle/ref_exec.py)",
                    python=sys.version.split()[0], numpy=np.__version__, scipy=scipy.__version__,
                    pandas=pd.__version__, ev_seed=EV_SEED, quick=a.quick)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
