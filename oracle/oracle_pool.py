"""Fan the CPU oracle (oracle/hpf_oracle.py) out over the host cores.

TEST / BENCH INFRASTRUCTURE ONLY: the checker of the at-scale parity tests and the ``cpu_baseline``
/ ``--impl reference`` legs of bench.py.  One worker process per core, the network and Y(h)
(HG:132-171) set up once per worker outside the timing, scenarios split in small chunks, BLAS
pinned to one thread per worker.  Workers are forked, so create the pool BEFORE a CUDA context
exists in the parent.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

_W = {}


def _init(on):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    import hpf_oracle as O
    _W.update(O=O, on=on, Y=O.build_admittance_matrices(on))


def _work(args):
    lo, P, Q, I_N, solver, keep_V = args
    O, on, Y = _W["O"], _W["on"], _W["Y"]
    S = P.shape[1]
    out = dict(lo=lo, n_iter_h=np.zeros(S, np.int32), n_iter_f=np.zeros(S, np.int32),
               status=np.zeros(S, np.int32), err_h=np.zeros(S))
    if keep_V:
        out["V_m"] = np.zeros((on.H, on.n, S))
        out["V_a"] = np.zeros((on.H, on.n, S))
        out["I_inj"] = np.zeros((on.n - on.m, on.H, S), dtype=complex)
    for s in range(S):
        o = O.hpf(on, P=P[:, s], Q=Q[:, s], I_N=I_N[:, :, s], Y=Y, solver=solver)
        out["n_iter_h"][s], out["n_iter_f"][s] = o["n_iter_h"], o["n_iter_f"]
        out["status"][s], out["err_h"][s] = o["status"], o["err_h"]
        if keep_V:
            out["V_m"][:, :, s], out["V_a"][:, :, s], out["I_inj"][:, :, s] = o["V_m"], o["V_a"], o["I_inj"]
    return out


class OraclePool:
    """``OraclePool(on, procs).solve(P, Q, I_N, solver)`` -> dict of batch-innermost result arrays
    (same layout as the C ABI) + ``wall_s``; ``solver`` = "superlu" (what the reference calls,
    HG:229,478) or "lapack" (the noise-floor variant)."""

    def __init__(self, on, procs=None):
        from multiprocessing import get_context
        self.procs = procs or os.cpu_count() or 1
        self.on = on
        self.pool = get_context("fork").Pool(self.procs, initializer=_init, initargs=(on,))
        self.pool.map(_work, [(0, on.P[:, None], on.Q[:, None], on.I_N[:, :, None], "superlu", False)] * self.procs)

    def solve(self, P, Q, I_N, solver="superlu", keep_V=True, chunk=None):
        S = P.shape[1]
        chunk = chunk or max(1, min(64, S // (self.procs * 4) or 1))
        jobs = [(lo, P[:, lo:lo + chunk], Q[:, lo:lo + chunk], I_N[:, :, lo:lo + chunk], solver, keep_V)
                for lo in range(0, S, chunk)]
        t0 = time.perf_counter()
        parts = self.pool.map(_work, jobs, chunksize=1)
        wall = time.perf_counter() - t0
        res = dict(wall_s=wall)
        for k in parts[0]:
            if k == "lo":
                continue
            res[k] = np.concatenate([p[k] for p in parts], axis=-1)
        return res

    def close(self):
        self.pool.close()
        self.pool.join()


def parity_rates(gpu, ref, floor=None):
    """Disagreement statistics of a GPU result dict against an oracle result dict (both batch-
    innermost): iteration-count mismatches and scenarios whose phasors differ by > 1e-9 relative
    (max over buses/harmonics of |V - V_ref| / |V_ref|).  ``floor``: a second oracle result (LAPACK
    step) - the same statistics of the reference against itself."""
    def ph(r):
        return r["V_m"] * np.exp(1j * r["V_a"])

    def one(a, b):
        S = len(b["n_iter_h"])
        Va, Vb = ph(a), ph(b)
        rel = (np.abs(Va - Vb) / np.abs(Vb)).reshape(-1, S).max(0)
        same = np.asarray(a["n_iter_h"]) == np.asarray(b["n_iter_h"])
        return dict(scenarios=int(S), iteration_mismatches=int((~same).sum()),
                    iteration_mismatch_rate=float((~same).mean()),
                    phasor_diff_gt_1e9=int((rel > 1e-9).sum()), phasor_diff_gt_1e9_rate=float((rel > 1e-9).mean()),
                    phasor_diff_median=float(np.median(rel)), phasor_diff_max=float(rel.max()),
                    phasor_diff_max_same_count=float(rel[same].max()) if same.any() else None,
                    n_iter_f_mismatches=int((np.asarray(a["n_iter_f"]) != np.asarray(b["n_iter_f"])).sum()))
    out = dict(vs_oracle=one(gpu, ref))
    if floor is not None:
        out["oracle_floor"] = one(floor, ref)
    return out
