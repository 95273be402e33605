"""Data contract of the solve path: network CSVs, Norton-equivalent CSVs, per-unit system.

Mirrors the ingest stage of the reference (HG = "Harmonic Power Flow/hcne_generalized.py"):
``init_lines_from_csv`` HG:45-61, ``init_buses_from_csv`` HG:77-94, ``init_network``
HG:113-128, ``import_Norton_Equivalents`` HG:278-310, bases HG:578-593.  Accepted on top of
the reference: the net1/Julia bus schema (``X_shunt`` instead of
``X_sh``; no ``G``/``B`` line columns), a trailing ``V_nom`` column (ignored like HG:93) and a
case-insensitive ``<component>_NE.csv`` lookup in a configurable directory.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import pandas as pd


@dataclass
class Settings:
    """The module-level constants of HG:578-585 and the loop parameters of HG:244,511."""
    BASE_POWER: float = 1000.0      # W
    BASE_VOLTAGE: float = 400.0     # V
    H_MAX: int = 51
    NET_FREQ: int = 50
    thresh_f: float = 1e-6
    max_iter_f: int = 30
    thresh_h: float = 1e-4
    max_iter_h: int = 50
    # HG:289-290 hard-codes this directory; overridable here or with $HPF_NE_DIR
    ne_dir: str = field(default_factory=lambda: os.environ.get(
        "HPF_NE_DIR", os.path.expanduser("~/Git/harmonic-power-flow/Circuit Simulation/")))
    device: int = 0

    @property
    def HARMONICS(self):
        return [h for h in range(1, self.H_MAX + 1, 2)]                  # HG:584

    @property
    def HARMONICS_FREQ(self):
        return [self.NET_FREQ * i for i in self.HARMONICS]               # HG:585

    @property
    def base_current(self):
        return self.BASE_POWER / self.BASE_VOLTAGE                       # HG:591

    @property
    def base_admittance(self):
        return self.base_current / self.BASE_VOLTAGE                     # HG:592

    @property
    def base_impedance(self):
        return 1 / self.base_admittance                                  # HG:593


def init_lines_from_csv(filename, settings: Settings):
    """``ID;fromID;toID;R;X[;G;B]`` in Ohm / Siemens -> p.u. DataFrame (HG:45-61)."""
    df = pd.read_csv(filename, delimiter=";")
    for col in ("G", "B"):
        if col not in df.columns:
            df[col] = 0.0
    for col in ("R", "X"):
        df[col] = df[col].astype(float) / settings.base_impedance
    for col in ("G", "B"):
        df[col] = df[col].astype(float) / settings.base_admittance
    return df


def init_buses_from_csv(filename, settings: Settings):
    """``ID;type;component;S;P;Q;X_sh[;V_nom]`` in W / Ohm -> p.u. DataFrame (HG:77-94)."""
    df = pd.read_csv(filename, delimiter=";")
    if "X_sh" not in df.columns and "X_shunt" in df.columns:
        df = df.rename(columns={"X_shunt": "X_sh"})
    for col in ("S", "P", "Q"):
        df[col] = df[col].astype(float) / settings.BASE_POWER
    df["X_sh"] = df["X_sh"].astype(float) / settings.base_impedance
    return df


def sort_network(buses, lines):
    """Reorder the buses into the order the solver relies on - slack, PV.., PQ.., nonlinear..
    (HG:83; the reference leaves the sorting as a TODO, HG:114) - keeping the file order inside
    each class, renumber the IDs 1..n and remap the line end points.
    -> (buses, lines, order) with order[k] = original row of the bus now at position k."""
    rank = {"slack": 0, "PV": 1, "PQ": 2, "nonlinear": 3}
    unknown = [t for t in buses["type"] if t not in rank]
    if unknown:
        raise ValueError("unknown bus type(s) %s" % sorted(set(map(str, unknown))))
    if int((buses["type"] == "slack").sum()) != 1:
        raise ValueError("exactly one slack bus is required")
    order = np.array(sorted(range(len(buses)), key=lambda i: (rank[buses["type"].iloc[i]], i)))
    old_ids = buses["ID"].to_numpy()[order]
    new_id = {int(o): k + 1 for k, o in enumerate(old_ids)}
    b2 = buses.iloc[order].reset_index(drop=True).copy()
    b2["ID"] = np.arange(1, len(b2) + 1)
    l2 = lines.copy()
    l2["fromID"] = [new_id[int(v)] for v in lines["fromID"]]
    l2["toID"] = [new_id[int(v)] for v in lines["toID"]]
    return b2, l2, order


def init_network(filename_buses, filename_lines, settings: Settings, auto_sort=False):
    """-> (buses, lines, m, n, c) with m = index of the first nonlinear bus (HG:113-128).
    auto_sort=True first brings the buses into the required order (``sort_network``); the default
    keeps the reference's behaviour and refuses an unsorted file."""
    buses = init_buses_from_csv(filename_buses, settings)
    lines = init_lines_from_csv(filename_lines, settings)
    if auto_sort:
        buses, lines, _ = sort_network(buses, lines)
    nl = buses.index[buses["type"] == "nonlinear"]
    m = int(min(nl)) if len(nl) > 0 else len(buses)
    n = len(buses)
    c = int((buses.type == "PV").sum()) + 1
    check_bus_order(buses, m, c)
    return buses, lines, m, n, c


def check_bus_order(buses, m, c):
    """The solver relies on the order slack, PV.., PQ.., nonlinear.. (HG:83)."""
    types = list(buses["type"])
    want = ["slack"] + ["PV"] * (c - 1) + ["PQ"] * (m - c) + ["nonlinear"] * (len(types) - m)
    if types != want:
        raise ValueError("buses must be ordered slack, PV.., PQ.., nonlinear.. (HG:83); got %s" % types)


def find_ne_file(component, ne_dir):
    """``<component>_NE.csv`` (the reference's text format) or ``<component>_NE.npz`` (binary, see
    ``write_ne_npz``; the reference leaves a faster format as a TODO, HG:282), case-insensitive."""
    names = sorted(os.listdir(ne_dir))
    for ext in (".csv", ".npz"):
        want = (component + "_NE" + ext).lower()
        for name in names:
            if name.lower() == want:
                return os.path.join(ne_dir, name)
    raise FileNotFoundError("no Norton-equivalent file for component %r in %s" % (component, ne_dir))


def write_ne_npz(path, freqs, Y_N_c, I_N_c, Y_N_uc, I_N_uc):
    """Binary Norton-equivalent file (SI units): exact complex128 values, no text parsing (a 50-order
    coupled equivalent takes 22 ms to parse from CSV, HG:291-299, and < 1 ms from this file)."""
    np.savez(path, freqs=np.asarray(freqs, dtype=np.int64), Y_N_c=np.asarray(Y_N_c, dtype=np.complex128),
             I_N_c=np.asarray(I_N_c, dtype=np.complex128), Y_N_uc=np.asarray(Y_N_uc, dtype=np.complex128),
             I_N_uc=np.asarray(I_N_uc, dtype=np.complex128))


def read_ne_npz(path):
    """-> the same (Parameter, Frequency) x frequency table ``read_ne_csv`` returns."""
    d = np.load(path)
    freqs = [int(f) for f in d["freqs"]]
    idx = pd.MultiIndex.from_arrays([["Y_N_c"] * len(freqs) + ["I_N_c", "Y_N_uc", "I_N_uc"], freqs + [0, 0, 0]],
                                    names=["Parameter", "Frequency"])
    data = np.vstack([d["Y_N_c"], d["I_N_c"][None, :], d["Y_N_uc"][None, :], d["I_N_uc"][None, :]])
    return pd.DataFrame(data, index=idx, columns=pd.Index(freqs, dtype=int))


def read_ne(path):
    return read_ne_npz(path) if path.lower().endswith(".npz") else read_ne_csv(path)


def read_ne_csv(path):
    """Parse an ``<device>_NE.csv`` (format of CS/NE_from_sim.py:195-209): header
    ``Parameter,Frequency,<f1>,<f2>..``; rows ``Y_N_c,<f_row>``.., ``I_N_c,0``, ``Y_N_uc,0``,
    ``I_N_uc,0``; cells are Python complex reprs ``(a+bj)`` (HG:291-299)."""
    df = pd.read_csv(path, index_col=["Parameter", "Frequency"])
    df.columns = df.columns.astype(int)
    return df.apply(lambda col: col.apply(lambda val: complex(val.strip("()"))))


def write_ne_csv(path, freqs, Y_N_c, I_N_c, Y_N_uc, I_N_uc):
    """Write the table ``read_ne_csv`` reads (SI units), same text format as the reference's."""
    freqs = [int(f) for f in freqs]
    with open(path, "w") as f:
        f.write("Parameter,Frequency," + ",".join(str(x) for x in freqs) + "\n")
        for r, fr in enumerate(freqs):
            f.write("Y_N_c,%d," % fr + ",".join(repr(complex(v)) for v in Y_N_c[r]) + "\n")
        for name, vec in (("I_N_c", I_N_c), ("Y_N_uc", Y_N_uc), ("I_N_uc", I_N_uc)):
            f.write("%s,0," % name + ",".join(repr(complex(v)) for v in vec) + "\n")


def import_Norton_Equivalents(buses, coupled, settings: Settings):
    """dict component -> [I_N (1 x H), Y_N (H x H coupled | 1 x H)] in p.u. (HG:278-310)."""
    NE = {}
    freqs = settings.HARMONICS_FREQ
    for device in buses.component[buses.type == "nonlinear"].unique():
        tab = read_ne(find_ne_file(device, settings.ne_dir))
        missing = [f for f in freqs if f not in tab.columns]
        if missing:   # the reference leaves this as a TODO (HG:295) and fails with a KeyError
            raise KeyError("%s_NE.csv lacks the frequencies %s" % (device, missing))
        tab = tab[freqs]
        if coupled:
            I_N = tab.loc["I_N_c"] / settings.base_current
            Y_N = tab.loc[("Y_N_c", freqs), freqs] / settings.base_admittance
        else:
            I_N = tab.loc["I_N_uc"] / settings.base_current
            Y_N = tab.loc["Y_N_uc"] / settings.base_admittance
        NE[device] = [I_N, Y_N]
    return NE


@dataclass
class PackedNet:
    """Everything the C ABI needs, as contiguous numpy arrays (p.u.)."""
    n: int
    m: int
    c: int
    harmonics: np.ndarray
    from_id: np.ndarray
    to_id: np.ndarray
    R: np.ndarray
    X: np.ndarray
    G: np.ndarray
    B: np.ndarray
    X_sh: np.ndarray
    P: np.ndarray
    Q: np.ndarray
    coupled: bool = True
    devices: list = field(default_factory=list)         # device names, index = device id
    dev_of_nl_bus: np.ndarray = None                    # (q,)
    Y_N: np.ndarray = None                              # (n_dev, H, H) | (n_dev, H) complex
    I_N: np.ndarray = None                              # (q, H) complex, nominal

    @property
    def H(self):
        return len(self.harmonics)

    @property
    def q(self):
        return self.n - self.m

    @property
    def N(self):
        return 2 * self.n * self.H - 1 - self.c

    @property
    def Nf(self):
        return 2 * self.n - 1 - self.c


def pack_network(buses, lines, m, n, c, harmonics, NE=None, coupled=True) -> PackedNet:
    f64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    net = PackedNet(n=int(n), m=int(m), c=int(c), harmonics=i32(harmonics),
                    from_id=i32(lines["fromID"]), to_id=i32(lines["toID"]),
                    R=f64(lines["R"]), X=f64(lines["X"]), G=f64(lines["G"]), B=f64(lines["B"]),
                    X_sh=f64(buses["X_sh"]), P=f64(buses["P"]), Q=f64(buses["Q"]), coupled=bool(coupled))
    if NE is not None:
        attach_devices(net, buses, NE, coupled)
    return net


def attach_devices(net: PackedNet, buses, NE, coupled):
    H = net.H
    comps = [buses.loc[i, "component"] for i in range(net.m, net.n)]
    net.devices = list(dict.fromkeys(comps))
    net.dev_of_nl_bus = np.ascontiguousarray([net.devices.index(cn) for cn in comps], dtype=np.int32)
    Y, I = [], []
    for d in net.devices:
        Y_N = np.asarray(NE[d][1], dtype=np.complex128)
        Y.append(Y_N.reshape(H, H) if coupled else Y_N.reshape(H))
    for cn in comps:
        I.append(np.asarray(NE[cn][0], dtype=np.complex128).reshape(H))
    net.coupled = bool(coupled)
    net.Y_N = np.ascontiguousarray(Y, dtype=np.complex128).reshape((len(net.devices),) + ((H, H) if coupled else (H,)))
    net.I_N = np.ascontiguousarray(I, dtype=np.complex128).reshape(net.q, H)
    return net
