"""Seeded synthetic networks for the large BASELINE configurations:
a radial feeder (trunk + laterals) and a meshed network (ring + chords), written in the
reference's own CSV schema (``<net>_buses.csv`` / ``<net>_lines.csv``, HG:45-94) together
with down-scaled Norton-equivalent device files, so that they go through exactly the same
loaders (``netio.init_network`` / ``import_Norton_Equivalents``) as net1..net3.

Scaling.  The reference networks carry 100..250 W per bus on 0.5..4 Ohm branches of a 400 V
system; hundreds of such loads on one feeder have no power-flow solution (the voltage drop
grows with buses x depth).  The generators therefore scale the linear loads AND the
nonlinear devices by ``load_scale``: a device ``smps_x<k>`` is an SMPS k times smaller, i.e.
both its Norton current spectrum I_N and its Norton admittance Y_N are multiplied by k
(file ``smps_x<k>_NE.csv`` written next to the network).
"""
from __future__ import annotations

import os

import numpy as np
import pandas as pd

from . import netio


def _finish(name, outdir, parent_edges, n, n_nl, n_pv, rng, load_scale, ne_src_dir, r_rng, x_rng):
    """Relabel so that the bus order is slack, PV.., PQ.., nonlinear.. (HG:83), draw loads and
    line parameters, write the CSVs and the scaled NE files; -> (buses_csv, lines_csv)."""
    order = np.arange(n)
    others = rng.permutation(np.arange(1, n))
    nl = np.sort(others[:n_nl])
    pv = np.sort(others[n_nl:n_nl + n_pv])
    pq = np.sort(others[n_nl + n_pv:])
    order = np.concatenate([[0], pv, pq, nl])          # position -> old node
    new_id = np.empty(n, dtype=int)
    new_id[order] = np.arange(1, n + 1)                 # old node -> 1-based ID
    tag = ("%g" % load_scale).replace(".", "p")
    devs = ["smps_x" + tag, "ev_x" + tag]
    types, comps, P, Q = [], [], [], []
    for pos in range(n):
        if pos == 0:
            types.append("slack"); comps.append("generator"); P.append(0.0); Q.append(0.0)
        elif pos < 1 + n_pv:
            types.append("PV"); comps.append("gen_%d" % pos)
            P.append(-float(rng.uniform(50, 250)) * load_scale); Q.append(0.0)
        elif pos < n - n_nl:
            types.append("PQ"); comps.append("lin_load_%d" % pos)
            P.append(float(rng.uniform(50, 250)) * load_scale)
            Q.append(float(rng.uniform(50, 250)) * load_scale)
        else:
            types.append("nonlinear"); comps.append(devs[(pos - (n - n_nl)) % 2])
            P.append(250.0 * load_scale); Q.append(100.0 * load_scale)
    buses = pd.DataFrame({"ID": np.arange(1, n + 1), "type": types, "component": comps,
                          "S": np.zeros(n), "P": P, "Q": Q,
                          "X_sh": [0.005] + [0.0] * (n - 1)})
    L = len(parent_edges)
    lines = pd.DataFrame({"ID": np.arange(1, L + 1),
                          "fromID": [int(new_id[a]) for a, _ in parent_edges],
                          "toID": [int(new_id[b]) for _, b in parent_edges],
                          "R": rng.uniform(*r_rng, L), "X": rng.uniform(*x_rng, L),
                          "G": np.zeros(L), "B": np.zeros(L)})
    os.makedirs(outdir, exist_ok=True)
    pb = os.path.join(outdir, name + "_buses.csv")
    pl = os.path.join(outdir, name + "_lines.csv")
    buses.to_csv(pb, sep=";", index=False)
    lines.to_csv(pl, sep=";", index=False)
    for base, dev in (("smps", devs[0]), ("ev", devs[1])):
        src = netio.read_ne_csv(netio.find_ne_file(base, ne_src_dir))
        freqs = list(src.columns)
        netio.write_ne_csv(os.path.join(outdir, dev + "_NE.csv"), freqs,
                           src.loc[("Y_N_c", freqs), freqs].to_numpy() * load_scale,
                           src.loc["I_N_c"].to_numpy().ravel() * load_scale,
                           src.loc["Y_N_uc"].to_numpy().ravel() * load_scale,
                           src.loc["I_N_uc"].to_numpy().ravel() * load_scale)
    return pb, pl


def radial_feeder(outdir, ne_src_dir, n=200, frac_nonlinear=0.4, lateral_len=3, seed=20260004,
                  load_scale=0.04, n_pv=0, name=None, r_rng=(0.2, 1.0), x_rng=(0.5, 4.0)):
    """BASELINE config 4: radial feeder, a trunk whose every bus carries a lateral of
    ``lateral_len`` buses; ``frac_nonlinear`` of the buses are SMPS / EV devices (alternating)."""
    rng = np.random.default_rng(seed)
    edges, trunk_prev, node = [], 0, 1
    while node < n:
        t = node; node += 1
        edges.append((trunk_prev, t))
        prev = t
        for _ in range(lateral_len):
            if node >= n:
                break
            edges.append((prev, node)); prev = node; node += 1
        trunk_prev = t
    name = name or "syn_radial%d" % n
    return _finish(name, outdir, edges, n, int(round(frac_nonlinear * n)), n_pv, rng, load_scale,
                   ne_src_dir, r_rng, x_rng)


def meshed(outdir, ne_src_dir, n=1000, frac_nonlinear=0.4, avg_degree=3.0, seed=20260005,
           load_scale=0.01, n_pv=0, name=None, r_rng=(0.2, 1.0), x_rng=(0.5, 4.0)):
    """BASELINE config 5: ring plus seeded chords (average degree ~ ``avg_degree``)."""
    rng = np.random.default_rng(seed)
    edges = [(i, (i + 1) % n) for i in range(n)]
    have = set((min(a, b), max(a, b)) for a, b in edges)
    want = int(round(avg_degree * n / 2))
    while len(edges) < want:
        a, b = (int(v) for v in rng.integers(0, n, 2))
        key = (min(a, b), max(a, b))
        if a == b or key in have:
            continue
        have.add(key); edges.append((a, b))
    name = name or "syn_meshed%d" % n
    return _finish(name, outdir, edges, n, int(round(frac_nonlinear * n)), n_pv, rng, load_scale,
                   ne_src_dir, r_rng, x_rng)
