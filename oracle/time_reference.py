"""Time the REFERENCE ITSELF (its own hpf(), HG:511-560) on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (the product never imports this).  Needs the reference checkout
(``$HPF_REFERENCE_ROOT`` or /root/reference): the reference's own code is exec-ed through the
in-memory compatibility patch of ``oracle/ref_exec.py`` - nothing is copied.  BASELINE.md section 4:
one reference interpreter per worker process, ``multiprocessing.Pool(P)`` with P = os.cpu_count(),
BLAS threads pinned to 1, stdout suppressed, the seeded scenarios of BASELINE config 3
(net3, coupled, odd harmonics <= 25; ``scenarios.draw_scenario`` seeds 0..S-1), wall-clock per
``hpf()`` call and for the whole fan-out.

    python oracle/time_reference.py [--scenarios 64] [--procs P] [--json out.json]

Used by ``bench.py --impl reference`` (kind "exec") when the checkout is present, and to produce
the builder-measured row of BASELINE.md section 2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

_S = {}


def _init(net, h_max, coupled):
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    import make_golden as MG
    t0 = time.perf_counter()
    _S["s"] = MG.session(net, h_max, coupled)          # exec of the module = its own nominal solve
    _S["t_setup"] = time.perf_counter() - t0
    _S["MG"] = MG


def _solve(args):
    seed, spread = args
    import numpy as np
    s, MG = _S["s"], _S["MG"]
    q, H = s.n - s.m, len(s.harmonics)
    P0 = s.base_buses["P"].to_numpy(dtype=float)
    Q0 = s.base_buses["Q"].to_numpy(dtype=float)
    P, Q, k = MG.draw_scenario(seed, P0, Q0, q, H, MG.SPREADS[spread])
    comps = [s.base_buses.loc[i, "component"] for i in range(s.m, s.n)]
    t0 = time.perf_counter()
    o = s.solve(P=P, Q=Q, I_N_scale={comp: k[j] for j, comp in enumerate(comps)})
    dt = time.perf_counter() - t0
    g = s.g
    t_loop = float(g["t_end_hpf_solve"] - g["t_start_hpf_solve"])          # the script's own timer (HG:617-618)
    return dict(seed=seed, s=dt, s_loop=t_loop, n_iter_h=int(o["n_iter_h"]), n_iter_f=int(o["n_iter_f"]),
                converged=bool(o["n_iter_h"] < 50 and o["n_iter_f"] < 30), err_h=float(o["err_h"]),
                V_m=np.asarray(o["V_m"]).tolist(), V_a=np.asarray(o["V_a"]).tolist())


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run(n_scen=64, procs=None, net="net3", h_max=25, coupled=True, spread="tight", seed0=0, keep_V=False):
    """-> dict with solves/s of the reference's own hpf() fanned out over `procs` processes."""
    from multiprocessing import get_context
    procs = procs or os.cpu_count() or 1
    with get_context("fork").Pool(procs, initializer=_init, initargs=(net, h_max, coupled)) as pool:
        pool.map(_solve, [(0, spread)] * procs)                  # first-call warm-up in every worker
        t0 = time.perf_counter()
        rows = pool.map(_solve, [(seed0 + s, spread) for s in range(n_scen)], chunksize=1)
        wall = time.perf_counter() - t0
    conv = sum(r["converged"] for r in rows)
    it = sum(r["n_iter_h"] for r in rows)
    per = sorted(r["s"] for r in rows)
    hist = {}
    for r in rows:
        hist[r["n_iter_h"]] = hist.get(r["n_iter_h"], 0) + 1
    out = dict(kind="exec", what="reference's own hpf() (HG:511-560) exec-ed through oracle/ref_exec.py",
               config="%s coupled=%s H<=%d, spread %s, seeds %d..%d" % (net, coupled, h_max, spread, seed0, seed0 + n_scen - 1),
               scenarios=n_scen, procs=procs, cpu_model=cpu_model(), wall_s=wall,
               solves_per_s=conv / wall, solves_per_s_per_core=conv / sum(per),
               converged=conv, s_per_solve_median=per[len(per) // 2], s_per_solve_mean=sum(per) / len(per),
               ms_per_nr_iteration_loop_only=1e3 * sum(r["s_loop"] for r in rows) / max(it, 1),
               n_iter_h_hist={str(k): v for k, v in sorted(hist.items())})
    if keep_V:
        out["rows"] = rows
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenarios", type=int, default=64)
    ap.add_argument("--procs", type=int, default=None)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    res = run(a.scenarios, a.procs)
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.json:
        open(a.json, "w").write(txt + "\n")
