"""Drop-in mirror of the reference's solve-path API (HG = "Harmonic Power Flow/
hcne_generalized.py"): same function names, argument order, defaults and return tuples
(SURVEY 8(b)), but every computation runs in the CUDA library:

    init_network(filename_buses, filename_lines)            HG:113   -> (buses, lines, m, n, c)
    build_admittance_matrices(buses, lines, harmonics)      HG:132   -> Y_all   (kernel 1)
    pf(Y, buses, thresh_f, max_iter_f)                      HG:244   -> (V, err_t, n_iter_f)
    import_Norton_Equivalents(buses, coupled)               HG:278   -> NE
    hpf(buses, lines, coupled, thresh_h, max_iter_h)        HG:511   -> (V, err_h, n_iter_h, J)
    get_THD(V)                                              HG:563   -> THD

The reference keeps its configuration in module globals assigned at the bottom of the file
(HG:578-597); here they live in ``settings`` (a ``netio.Settings``), changed with
``configure(H_MAX=25, ne_dir=...)``; the module constants (HARMONICS, ...) follow.
Differences on purpose: nothing is solved at import time, nothing is printed unless
``configure(verbose=True)``, and m/n/c are derived from ``buses`` instead of read from globals.
``hpf_batch`` is new API (the reference has no notion of a batch).  GPU handles are cached per
network (``_solver_for``): repeated ``hpf`` / ``pf`` / ``get_THD`` calls on the same network reuse
one handle and its structured-step set-up instead of creating and destroying one per call.  The
fourth element of ``hpf``'s return tuple, the Jacobian of the LAST iteration, is assembled on first
use (``LastJacobian``): the reference builds it anyway, here it would cost a replay of the solve.
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np
import pandas as pd
from scipy.sparse import csr_matrix

from . import netio
from .netio import Settings
from .solver import BatchSolver

settings = Settings()
verbose = False
idx = pd.IndexSlice


def _export_constants():
    g = globals()
    g["BASE_POWER"], g["BASE_VOLTAGE"] = settings.BASE_POWER, settings.BASE_VOLTAGE
    g["H_MAX"], g["NET_FREQ"] = settings.H_MAX, settings.NET_FREQ
    g["HARMONICS"], g["HARMONICS_FREQ"] = settings.HARMONICS, settings.HARMONICS_FREQ
    g["base_current"], g["base_admittance"] = settings.base_current, settings.base_admittance
    g["base_impedance"] = settings.base_impedance


def configure(**kw):
    """Change the module configuration (HG:578-585 and the hard-coded NE directory HG:289)."""
    global verbose
    if "verbose" in kw:
        verbose = bool(kw.pop("verbose"))
    for k, v in kw.items():
        if not hasattr(settings, k):
            raise AttributeError("unknown setting %r" % k)
        setattr(settings, k, v)
    _export_constants()


_export_constants()


_SOLVERS: "OrderedDict[str, BatchSolver]" = OrderedDict()
_MAX_SOLVERS = 8


def _solver_for(net, tag=""):
    """One BatchSolver (= one GPU handle) per distinct network, reused across calls (LRU of 8)."""
    h = hashlib.sha1(tag.encode())
    h.update(repr((net.n, net.m, net.c, bool(net.coupled), list(net.devices or []), settings.device)).encode())
    for a in (net.harmonics, net.from_id, net.to_id, net.R, net.X, net.G, net.B, net.X_sh, net.dev_of_nl_bus, net.Y_N):
        if a is not None:
            h.update(np.ascontiguousarray(a).tobytes())
    key = h.hexdigest()
    sol = _SOLVERS.get(key)
    if sol is not None and not getattr(sol, "_h", None):          # closed by a caller of hpf_batch
        sol = None
    if sol is None:
        sol = _SOLVERS[key] = BatchSolver(net, settings.device)
        while len(_SOLVERS) > _MAX_SOLVERS:
            _SOLVERS.popitem(last=False)[1].close()
    else:
        _SOLVERS.move_to_end(key)
    return sol


def release_solvers():
    """Destroy the cached GPU handles (they are also released at interpreter exit)."""
    while _SOLVERS:
        _SOLVERS.popitem()[1].close()


class LastJacobian:
    """The Jacobian of the last harmonic iteration (4th element of ``hpf``'s return tuple, HG:560),
    a scipy CSR matrix assembled ON FIRST USE: the solver keeps only the final state, so the last
    Jacobian needs a replay of the first n_iter_h - 1 steps (the kernels are deterministic) and one
    launch of the Jacobian kernel - paid only by callers that look at it."""

    def __init__(self, make):
        self._make, self._J = make, None

    def materialise(self):
        if self._J is None:
            self._J, self._make = self._make(), None
        return self._J

    def toarray(self):
        return self.materialise().toarray()

    def __getattr__(self, name):                       # shape, nnz, dot, tocsc, ... of the CSR matrix
        return getattr(self.materialise(), name)

    def __matmul__(self, other):
        return self.materialise() @ other


def _mnc(buses):
    nl = buses.index[buses["type"] == "nonlinear"]
    n = len(buses)
    m = int(min(nl)) if len(nl) > 0 else n
    c = int((buses.type == "PV").sum()) + 1
    return m, n, c


def init_network(filename_buses, filename_lines, from_csv=True):
    if not from_csv:
        raise NotImplementedError("init_*_manually (HG:64-74,97-110) is dead code in the reference "
                                  "(it reads a column that does not exist); use the CSV files")
    return netio.init_network(filename_buses, filename_lines, settings)


def _V_frame(V_m, V_a, harmonics, n):
    mi = pd.MultiIndex.from_product([list(harmonics), list(range(n))], names=["harmonic", "bus"])
    return pd.DataFrame({"V_m": np.asarray(V_m).reshape(-1), "V_a": np.asarray(V_a).reshape(-1)}, index=mi)


def build_admittance_matrices(buses, lines, harmonics):
    """Y(h) for all harmonics as the reference's (harmonic, bus) x bus complex DataFrame."""
    m, n, c = _mnc(buses)
    net = netio.pack_network(buses, lines, m, n, c, harmonics)
    net.m = n                      # Y does not depend on the bus classes; no devices needed
    Y = _solver_for(net, "Y").Y.cpu().numpy()
    mi = pd.MultiIndex.from_product([list(harmonics), buses.index.values], names=["harmonic", "bus"])
    return pd.DataFrame(Y.reshape(len(harmonics) * n, n), index=mi, columns=[buses.index.values])


def _empty_lines():
    return pd.DataFrame({k: [] for k in ("ID", "fromID", "toID", "R", "X", "G", "B")})


def pf(Y, buses, thresh_f=1e-6, max_iter_f=30, plt_convergence=False):
    """Fundamental power flow on the GIVEN admittance table (uses Y.loc[1] like HG:255)."""
    m, n, c = _mnc(buses)
    harmonics = settings.HARMONICS
    net = netio.pack_network(buses, _empty_lines(), m, n, c, harmonics)
    net.m = n
    sol = _solver_for(net, "pf")
    Yh = np.zeros((len(harmonics), n, n), dtype=np.complex128)
    Yh[0] = np.asarray(Y.loc[1], dtype=np.complex128)
    sol.set_Y(Yh)
    V_m, V_a, nf, err, hist = sol.fund_solve(net.P[:, None], net.Q[:, None], thresh_f, max_iter_f,
                                             history=True)
    n_iter_f = int(nf.item())
    hist = hist[:, 0].cpu().numpy()
    err_t = {k: float(hist[k + 1]) for k in range(n_iter_f)}             # HG:264
    V = _V_frame(V_m[:, :, 0].cpu().numpy(), V_a[:, :, 0].cpu().numpy(), harmonics, n)
    if verbose:
        print(V.loc[1])
        print(("Fundamental power flow converged after %d iterations." % n_iter_f)
              if n_iter_f < max_iter_f else ("Warning! Maximum of %d iterations reached." % n_iter_f))
    return V, err_t, n_iter_f


def import_Norton_Equivalents(buses, coupled):
    return netio.import_Norton_Equivalents(buses, coupled, settings)


def _pack(buses, lines, coupled, NE=None):
    m, n, c = _mnc(buses)
    netio.check_bus_order(buses, m, c)
    if NE is None:
        NE = import_Norton_Equivalents(buses, coupled) if m < n else {}
    return netio.pack_network(buses, lines, m, n, c, settings.HARMONICS, NE if m < n else None, coupled)


def make_solver(buses, lines, coupled, NE=None):
    """Pack a network + its Norton equivalents and bind it to the GPU (new API)."""
    return _solver_for(_pack(buses, lines, coupled, NE), "hpf")


def hpf(buses, lines, coupled, thresh_h=1e-4, max_iter_h=50, plt_convergence=False):
    """Harmonic power flow of ONE scenario, reference return tuple (V, err_h, n_iter_h, J)."""
    # (the handle is cached per network; the scenario - loads and Norton currents - is this call's)
    net = _pack(buses, lines, coupled)
    sol = _solver_for(net, "hpf")
    P, Q = net.P[:, None], net.Q[:, None]
    I_N = net.I_N[:, :, None] if net.q > 0 else None
    kw = dict(thresh_f=settings.thresh_f, max_iter_f=settings.max_iter_f, thresh_h=thresh_h)
    r = sol.solve(P, Q, I_N, max_iter_h=max_iter_h, **kw)
    n_iter_h = int(r.n_iter_h.item())
    err_h = float(r.err_h.item())
    V = _V_frame(r.V_m[:, :, 0].cpu().numpy(), r.V_a[:, :, 0].cpu().numpy(), settings.HARMONICS, net.n)
    J = None
    if n_iter_h > 0:
        def last_jacobian():
            # the reference returns the Jacobian of its LAST iteration (built before the last
            # update): replay n_iter_h - 1 steps (the kernel is deterministic) and assemble it
            raw = sol.solve(P, Q, I_N, max_iter_h=n_iter_h - 1, raw=True, want_I_inj=False, **kw)
            return csr_matrix(sol.jacobian_view(sol.jacobian(raw.V_m, raw.V_a))[0].cpu().numpy())
        J = LastJacobian(last_jacobian)
    if verbose:
        print(V)
        print(("Harmonic power flow converged after %d iterations." % n_iter_h)
              if n_iter_h < max_iter_h else ("Warning! Maximum of %d iterations reached." % n_iter_h))
    return V, err_h, n_iter_h, J


def hpf_batch(buses, lines, coupled, P, Q, I_N, thresh_h=1e-4, max_iter_h=50, NE=None, **kw):
    """Batched solve: P, Q [n, B] p.u., I_N [q, H, B] p.u. -> (BatchResult, BatchSolver)."""
    sol = make_solver(buses, lines, coupled, NE)
    return sol.solve(P, Q, I_N, thresh_f=settings.thresh_f, max_iter_f=settings.max_iter_f,
                     thresh_h=thresh_h, max_iter_h=max_iter_h, **kw), sol


def get_THD(V):
    """THD_F, THD_R per bus (HG:563-572), evaluated by the THD kernel."""
    harmonics = sorted(set(V.index.get_level_values(0)))
    n = len(set(V.index.get_level_values(1)))
    H = len(harmonics)
    z = np.zeros(n)
    net = netio.PackedNet(n=n, m=n, c=1, harmonics=np.ascontiguousarray(harmonics, dtype=np.int32),
                          from_id=np.zeros(0, np.int32), to_id=np.zeros(0, np.int32), R=np.zeros(0),
                          X=np.zeros(0), G=np.zeros(0), B=np.zeros(0), X_sh=z, P=z, Q=z)
    o = _solver_for(net, "thd").thd(V["V_m"].to_numpy().reshape(H, n, 1).copy()).cpu().numpy()
    return pd.DataFrame({"THD_F": o[0, :, 0], "THD_R": o[1, :, 0]})
