"""Norton-equivalent extraction from circuit-simulation results - the step BEFORE the solve path
(SURVEY 8(f) next-1), mirror of the reference's ``Circuit Simulation/NE_from_sim.py`` (``NE``).

The reference is a flat script for one device; here the ingest (``load_simulation``, NE:26-84) is
host-side pandas-free numpy, and the arithmetic - the two-point uncoupled equivalents
(NE:86-118) and the coupled equivalent, one (N+1) x (N+1) complex solve with N right-hand sides
per device (NE:141-173) - runs on the GPU for a whole batch of devices / operating points at once
(``hpf_ne_extract``, one CTA per device, Gauss-Jordan with partial pivoting in shared memory).
``export_NE`` writes the ``<device>_NE.csv`` the solve path reads (NE:195-209).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import netio


@dataclass
class Simulation:
    """Measurement set of one device (see oracle/ne_oracle.py for the shapes)."""
    freq: np.ndarray      # (N,) Hz, freq[0] = fundamental
    Vf: np.ndarray        # (2,)
    Vh: np.ndarray        # (2, K)
    I_f: np.ndarray       # (2, N)
    I_h: np.ndarray       # (2, K, N)


def load_simulation(mat_path, net_freq=50) -> Simulation:
    """Read a Simulink result file (struct ``all`` with ``results_f`` / ``results_h``, NE:30-33) and
    collect what the extraction uses: the first two magnitude sets of the harmonic sweeps and the
    two fundamental measurements, spectra restricted to the supply frequencies (NE:84,160-164)."""
    from scipy.io import loadmat
    data = loadmat(mat_path, squeeze_me=True, struct_as_record=False)
    df, dh = data["all"].results_f, data["all"].results_h
    if dh.shape[1] < 2:
        raise ValueError("At least 2 measurements needed for script to work.")       # NE:38-39
    if dh.shape[0] < 2:
        raise ValueError("At least 2 harmonics needed for script to work.")           # NE:40-41
    supply_f = [int(row[0].f_h) for row in dh]                                         # NE:45-46
    freq = np.array([net_freq] + supply_f)
    K, N = len(supply_f), len(supply_f) + 1
    H = np.asarray(dh[0, 0].H).astype(int)                                             # NE:59
    col = np.array([int(np.nonzero(H == f)[0][0]) for f in freq])
    spec = lambda r: (np.asarray(r.I_inj) * np.exp(1j * np.asarray(r.I_inj_phase)))[col]   # NE:72,80
    deg = lambda m, a: m * np.exp(1j * a * np.pi / 180)                                # NE:74,91
    Vf = np.array([deg(df[0].V_m_f, df[0].V_a_f), deg(df[1].V_m_f, df[1].V_a_f)])
    Vh = np.array([[deg(dh[k, a].V_m_h, dh[k, a].V_a_h) for k in range(K)] for a in range(2)])
    I_f = np.array([spec(df[0]), spec(df[1])])
    I_h = np.array([[spec(dh[k, a]) for k in range(K)] for a in range(2)])
    return Simulation(freq, Vf, Vh, I_f, I_h)


def get_NE_from_sim(sims, device=0):
    """Extract the Norton equivalents of a batch of devices (all with the same number of supply
    frequencies) on the GPU.  ``sims``: Simulation or list.  -> dict of arrays with a leading
    device axis: Y_N_c (D, N, N), I_N_c (D, N), Y_N_uc (D, N), I_N_uc (D, N), info (D,)."""
    import torch
    from .solver import ne_extract
    single = isinstance(sims, Simulation)
    sims = [sims] if single else list(sims)
    Vf = np.stack([s.Vf for s in sims]); Vh = np.stack([s.Vh for s in sims])
    I_f = np.stack([s.I_f for s in sims]); I_h = np.stack([s.I_h for s in sims])
    out = ne_extract(Vf, Vh, I_f, I_h, device=device)
    out = {k: v.cpu().numpy() for k, v in out.items()}
    if (out["info"] != 0).any():
        raise np.linalg.LinAlgError("singular measurement matrix for device(s) %s" % np.nonzero(out["info"])[0])
    return {k: v[0] for k, v in out.items()} if single else out


def export_NE(path, freq, ne):
    """``<device>_NE.csv`` in the reference's format (NE:195-209)."""
    netio.write_ne_csv(path, freq, ne["Y_N_c"], ne["I_N_c"], ne["Y_N_uc"], ne["I_N_uc"])
