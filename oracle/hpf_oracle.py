"""CPU oracle: numpy restatement of the reference's harmonic power-flow solve path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module, and only as the checker / the CPU arm — never the product path
(``harmonic_power_flow_b200`` must fail loudly if its CUDA library is missing;
it never falls back to this code).

Parity status: PINNED against the reference itself.  The reference
(``Harmonic Power Flow/hcne_generalized.py`` = ``HG``) has no tests or golden
vectors of its own, so ``oracle/make_golden.py`` runs the
reference's own code in the build container (``oracle/ref_exec.py``) and commits
its inputs/outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against those fixtures (iteration counts identical, every
intermediate and result to <= 1e-12).

Every function cites the HG lines it restates.  Stacked vectors are
harmonic-major, bus-minor (index ``h*n + i``, HG:139-140,175-176), state arrays
are kept as ``(H, n)``.  Symbols: ``n`` buses, ``m`` = 0-based index of the first
nonlinear bus, ``c`` = 1 + number of PV buses (HG:122-127), ``H`` harmonic orders
including the fundamental.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.linalg import spsolve

BASE_POWER = 1000.0    # W   HG:578
BASE_VOLTAGE = 400.0   # V   HG:579
BASE_CURRENT = BASE_POWER / BASE_VOLTAGE            # HG:591
BASE_ADMITTANCE = BASE_CURRENT / BASE_VOLTAGE       # HG:592
BASE_IMPEDANCE = 1 / BASE_ADMITTANCE                # HG:593


@dataclass
class Net:
    """Per-unit network (what HG:45-128 produce), plain arrays."""
    n: int
    m: int
    c: int
    harmonics: np.ndarray          # (H,) int harmonic orders, harmonics[0] == 1
    line_from: np.ndarray          # (L,) 1-based bus IDs as in the CSV
    line_to: np.ndarray
    R: np.ndarray
    X: np.ndarray
    G: np.ndarray
    B: np.ndarray
    X_sh: np.ndarray               # (n,)
    P: np.ndarray                  # (n,) p.u., loads positive (HG:84-85)
    Q: np.ndarray
    # Norton equivalents per nonlinear bus (HG:301-308), p.u.
    I_N: np.ndarray = field(default=None)    # (q, H) complex
    Y_N: np.ndarray = field(default=None)    # (q, H, H) coupled / (q, H) uncoupled
    coupled: bool = True

    @property
    def H(self):
        return len(self.harmonics)

    @property
    def N(self):
        return 2 * self.n * self.H - 1 - self.c


def build_admittance_matrices(net: Net) -> np.ndarray:
    """Y(h) for every harmonic, (H, n, n) complex.  Restates HG:132-171 incl. quirks:

    * off-diagonals are ASSIGNED, so a later parallel line overwrites (HG:150-155);
    * the diagonal is minus the sequential row sum (HG:159,161) plus the bus shunt
      ``1/(j X_sh h)`` only when ``X_sh != 0 and h != 1`` (HG:158-159);
    * the pi-line shunt ``(G + j h B)/2`` is added to bus index ``k`` when a line's
      1-based ``fromID``/``toID`` EQUALS the 0-based ``k`` (HG:163-168) - i.e. it
      lands one bus too low and never on the last bus.
    """
    n, H = net.n, net.H
    Y_all = np.zeros((H, n, n), dtype=complex)
    for hi, h in enumerate(net.harmonics):
        h = int(h)
        Y = np.zeros((n, n), dtype=complex)
        for l in range(len(net.R)):
            f, t = int(net.line_from[l]) - 1, int(net.line_to[l]) - 1
            Y[f, t] = -1 / (net.R[l] + 1j * net.X[l] * h)
            Y[t, f] = Y[f, t]
        for k in range(n):
            s = 0
            for j in range(n):
                s = s + Y[k, j]          # builtin sum(): sequential, Y[k,k] is still 0
            if net.X_sh[k] != 0 and h != 1:
                Y[k, k] = -s + 1 / (1j * net.X_sh[k] * h)
            else:
                Y[k, k] = -s
            for l in range(len(net.R)):
                if net.line_from[l] == k or net.line_to[l] == k:
                    Y[k, k] = Y[k, k] + (net.G[l] + 1j * h * net.B[l]) / 2
        Y_all[hi] = Y
    return Y_all


def build_admittance_matrices_trafo(net: Net, tau, phase_shift_deg) -> np.ndarray:
    """Y(h) with transformer branches - restates build_admittance_matrices of the reference's
    "Fundamental Power Flow/pi_trafo_pf_test.py" (FPF:99-149): off-diagonals divided by
    tau e^{-+j phi} (FPF:117-124), the pi-shunt of a branch whose toID equals the 0-based row index
    divides the accumulated diagonal by tau^2, and only if its fromID did not match (FPF:137-145)."""
    n, H = net.n, net.H
    Y_all = np.zeros((H, n, n), dtype=complex)
    for hi, h in enumerate(net.harmonics):
        h = int(h)
        Y = np.zeros((n, n), dtype=complex)
        for l in range(len(net.R)):
            f, t = int(net.line_from[l]) - 1, int(net.line_to[l]) - 1
            Y[f, t] = -1 / (net.R[l] + 1j * net.X[l] * h) / \
                (tau[l] * np.exp(-1j * phase_shift_deg[l] / 180 * np.pi))
            Y[t, f] = -1 / (net.R[l] + 1j * net.X[l] * h) / \
                (tau[l] * np.exp(1j * phase_shift_deg[l] / 180 * np.pi))
        for k in range(n):
            s = 0
            for j in range(n):
                s = s + Y[k, j]
            if net.X_sh[k] != 0 and h != 1:
                Y[k, k] = -s + 1 / (1j * net.X_sh[k] * h)
            else:
                Y[k, k] = -s
            for l in range(len(net.R)):
                if net.line_from[l] == k:
                    Y[k, k] = Y[k, k] + (net.G[l] + 1j * h * net.B[l]) / 2
                elif net.line_to[l] == k:
                    Y[k, k] = (Y[k, k] + (net.G[l] + 1j * h * net.B[l]) / 2) / (tau[l] ** 2)
        Y_all[hi] = Y
    return Y_all


def init_voltages(net: Net):
    """Flat start: |V|=1 at h=1, 0.1 at h>1, angle 0 (HG:174-184)."""
    V_m = np.full((net.H, net.n), 0.1)
    V_m[0, :] = 1.0
    V_a = np.zeros((net.H, net.n))
    return V_m, V_a


def _solve(J, f, solver):
    if solver == "superlu":          # what HG:229,478 call (scipy spsolve -> SuperLU/COLAMD)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return spsolve(csr_matrix(J), f)
    if solver == "lapack":           # noise-floor variant
        return np.linalg.solve(J, f)
    raise ValueError(solver)


# ----------------------------------------------------------------------------
# fundamental stage
def fund_mismatch(net: Net, P, Q, V_m1, V_a1, Y1):
    """HG:195-202: V*conj(Y1 V) + S for ALL buses; f = [Re[1:], Im[c:]]."""
    V = V_m1 * np.exp(1j * V_a1)
    mis = V * np.conj(Y1.dot(V)) + (P + 1j * Q)
    f = np.r_[mis.real[1:], mis.imag[net.c:]]
    return f, np.abs(f).max()


def _power_derivatives(V, Vnorm, Y1):
    """dS/dtheta and dS/d|V| of S_i = V_i conj((Y1 V)_i) (HG:208-214 / HG:453-459).

    dSdA = j diag(V) conj(diag(I) - Y1 diag(V));
    dSdV = diag(Vnorm) conj(diag(I)) + diag(V) conj(Y1 diag(Vnorm)).
    """
    I = Y1 @ V
    dSdA = (1j * V)[:, None] * np.conj(np.diag(I) - Y1 * V[None, :])
    dSdV = np.diag(Vnorm * np.conj(I)) + V[:, None] * np.conj(Y1 * Vnorm[None, :])
    return dSdA, dSdV


def build_jacobian(net: Net, V_m1, V_a1, Y1):
    """Fundamental polar Jacobian (HG:205-223); normalises with abs(V) (HG:210)."""
    c = net.c
    V = V_m1 * np.exp(1j * V_a1)
    dSdA, dSdV = _power_derivatives(V, V / np.abs(V), Y1)
    return np.block([[dSdA[1:, 1:].real, dSdV[1:, c:].real],
                     [dSdA[c:, 1:].imag, dSdV[c:, c:].imag]])


def pf(net: Net, Y, P=None, Q=None, thresh_f=1e-6, max_iter_f=30, solver="superlu"):
    """Fundamental Newton-Raphson (HG:244-275).  Returns V_m, V_a (H,n), err list, n_iter_f."""
    P = net.P if P is None else P
    Q = net.Q if Q is None else Q
    n, c = net.n, net.c
    V_m, V_a = init_voltages(net)
    Y1 = Y[0]
    x = np.append(V_a[0, 1:], V_m[0, c:])                    # HG:187-192
    f, err = fund_mismatch(net, P, Q, V_m[0], V_a[0], Y1)
    errs = [err]
    n_iter_f = 0
    while err > thresh_f and n_iter_f < max_iter_f:           # HG:259
        J = build_jacobian(net, V_m[0], V_a[0], Y1)
        x = x - _solve(J, f, solver)                          # HG:226-230
        V_a[0, 1:] = x[:n - 1]                                # HG:233-235
        V_m[0, c:] = x[n - 1:]
        f, err = fund_mismatch(net, P, Q, V_m[0], V_a[0], Y1)
        errs.append(err)
        n_iter_f += 1
    return V_m, V_a, errs, n_iter_f


# ----------------------------------------------------------------------------
# harmonic stage
def current_injections(net: Net, V, I_N):
    """I_inj = I_N - Y_N V_bus for every nonlinear bus, (q, H) (HG:313-323)."""
    q = net.n - net.m
    out = np.empty((q, net.H), dtype=complex)
    for k in range(q):
        Vb = V[:, net.m + k]
        if net.coupled:
            out[k] = I_N[k] - net.Y_N[k].dot(Vb)              # HG:320
        else:
            out[k] = I_N[k] - np.diag(net.Y_N[k]).dot(Vb)     # HG:322
    return out


def harmonic_mismatch(net: Net, P, Q, V_m, V_a, Y, I_N):
    """HG:326-390.  f = [Re f_c, Im f_c[c-1:]], f_c = [dS (buses 1..m-1), dI_f (nonlinear
    buses, h=1), dI_h (all buses, h>=3)]; injections are ADDED (HG:351,354)."""
    n, m, c, H = net.n, net.m, net.c, net.H
    V = V_m * np.exp(1j * V_a)
    V1 = V[0]
    dS = (P[1:m] + 1j * Q[1:m]) + V1[1:m] * np.conj(Y[0][1:m, :] @ V1)   # HG:372-380
    dI_f = Y[0][m:, :] @ V1                                               # HG:335-339
    dI_h = np.einsum("hij,hj->hi", Y[1:], V[1:])                          # HG:342-345
    I_inj = current_injections(net, V, I_N)
    for k in range(n - m):
        dI_f[k] += I_inj[k, 0]
        dI_h[:, m + k] += I_inj[k, 1:]
    f_c = np.concatenate([dS, dI_f, dI_h.ravel()])
    f = np.concatenate([f_c.real, f_c[c - 1:].imag])                      # HG:388
    return f, np.linalg.norm(f, np.inf)


def build_harmonic_jacobian(net: Net, V_m, V_a, Y):
    """Dense N x N harmonic Jacobian (HG:401-473).

    Normalises with the SIGNED V_m (HG:405,422,455), so V/V_m = e^{j theta} even
    for negative magnitudes.  Row order [P(1..m-1); Re I; Q(c..m-1); Im I],
    column order [theta(1..nH-1), V_m(c..nH-1)] (HG:445-446,464-472).
    """
    n, m, c, H = net.n, net.m, net.c, net.H
    nH = n * H
    V = (V_m * np.exp(1j * V_a))
    Vn = V / V_m
    Vf, Vnf = V.ravel(), Vn.ravel()
    dIdV = np.zeros((nH, nH), dtype=complex)
    dIdA = np.zeros((nH, nH), dtype=complex)
    for h in range(H):                                                    # HG:407-411
        s = slice(h * n, (h + 1) * n)
        dIdV[s, s] = Y[h] * Vn[h][None, :]
        dIdA[s, s] = (1j * Y[h]) * V[h][None, :]
    for k in range(n - m):                                                # HG:425-443
        i = m + k
        rows = np.arange(H) * n + i
        if net.coupled:
            dIdV[np.ix_(rows, rows)] -= net.Y_N[k] * Vnf[rows][None, :]
            dIdA[np.ix_(rows, rows)] -= (1j * net.Y_N[k]) * Vf[rows][None, :]
        else:
            dIdV[rows, rows] -= net.Y_N[k] * Vnf[rows]
            dIdA[rows, rows] -= (1j * net.Y_N[k]) * Vf[rows]
    dIdA = dIdA[m:, 1:]                                                   # HG:445-446
    dIdV = dIdV[m:, c:]
    dS1dA1, dS1dV1 = _power_derivatives(V[0], Vn[0], Y[0])                # HG:451-459
    dSdA = np.zeros((n, nH), dtype=complex)
    dSdV = np.zeros((n, nH), dtype=complex)
    dSdA[:, :n] = dS1dA1
    dSdV[:, :n] = dS1dV1
    return np.block([[dSdA[1:m, 1:].real, dSdV[1:m, c:].real],
                     [dIdA.real, dIdV.real],
                     [dSdA[c:m, 1:].imag, dSdV[c:m, c:].imag],
                     [dIdA.imag, dIdV.imag]])


def postprocess(V_m, V_a):
    """HG:547-549: theta += pi where V_m < 0; theta %= 2 pi; V_m = |V_m|."""
    V_a = V_a.copy()
    V_m = V_m.copy()
    neg = V_m < 0
    V_a[neg] += np.pi
    V_a = V_a % (2 * np.pi)
    V_m[neg] = -V_m[neg]
    return V_m, V_a


def get_THD(net: Net, V_m):
    """HG:563-572: THD_F, THD_R per bus from magnitudes of harmonics >= 3."""
    hm = np.asarray(net.harmonics) >= 3
    out = np.zeros((net.n, 2))
    for b in range(net.n):
        sh = 0
        for v in V_m[hm, b]:
            sh = sh + v ** 2
        sa = 0
        for v in V_m[:, b]:
            sa = sa + v ** 2
        out[b, 0] = np.sqrt(sh) / V_m[0, b]
        out[b, 1] = np.sqrt(sh) / np.sqrt(sa)
    return out


def hpf(net: Net, P=None, Q=None, I_N=None, thresh_f=1e-6, max_iter_f=30,
        thresh_h=1e-4, max_iter_h=50, solver="superlu", Y=None, trace=False):
    """Whole solve (HG:511-560).  Returns a dict with the post-processed V, iteration
    counts (= number of Jacobian solves), error histories, I_inj and status."""
    P = net.P if P is None else np.asarray(P, dtype=float)
    Q = net.Q if Q is None else np.asarray(Q, dtype=float)
    I_N = net.I_N if I_N is None else np.asarray(I_N, dtype=complex)
    n, c, H = net.n, net.c, net.H
    nH = n * H
    if Y is None:
        Y = build_admittance_matrices(net)
    V_m, V_a, err_f_hist, n_iter_f = pf(net, Y, P, Q, thresh_f, max_iter_f, solver)
    V_fund = (V_m.copy(), V_a.copy())
    f, err_h = harmonic_mismatch(net, P, Q, V_m, V_a, Y, I_N)
    x = np.append(V_a.ravel()[1:], V_m.ravel()[c:])                       # HG:393-398
    err_h_hist = [err_h]
    tr = dict(f=[f.copy()], J=[], x=[])
    n_iter_h = 0
    J = None
    while err_h > thresh_h and n_iter_h < max_iter_h:                     # HG:536
        J = build_harmonic_jacobian(net, V_m, V_a, Y)
        x = x - _solve(J, f, solver)                                      # HG:476-479
        V_a.ravel()[1:] = x[:nH - 1]                                      # HG:484-485
        V_m.ravel()[c:] = x[nH - 1:]
        f, err_h = harmonic_mismatch(net, P, Q, V_m, V_a, Y, I_N)
        err_h_hist.append(err_h)
        if trace:
            tr["J"].append(J)
            tr["x"].append(x.copy())
            tr["f"].append(f.copy())
        n_iter_h += 1
    I_inj = current_injections(net, V_m * np.exp(1j * V_a), I_N)
    V_m_out, V_a_out = postprocess(V_m, V_a)
    if not np.all(np.isfinite(x)):
        status = 3
    elif n_iter_h >= max_iter_h or n_iter_f >= max_iter_f:               # HG:270-274,555-559
        status = 1
    else:
        status = 0
    out = dict(V_m=V_m_out, V_a=V_a_out, err_h=err_h, n_iter_h=n_iter_h,
               n_iter_f=n_iter_f, err_f_hist=np.array(err_f_hist),
               err_h_hist=np.array(err_h_hist), I_inj=I_inj, status=status,
               THD=get_THD(net, V_m_out), V_fund=V_fund, J_last=J)
    if trace:
        out["trace"] = tr
    return out


# ----------------------------------------------------------------------------
# fixture loader (tests/golden/networks.json + ne_devices.npz -> Net), p.u. as HG:57-60,89-92,301-308
def net_from_golden(golden_dir, name, h_max, coupled, net_freq=50):
    import json
    import os
    tab = json.load(open(os.path.join(golden_dir, "networks.json")))[name]
    dev = np.load(os.path.join(golden_dir, "ne_devices.npz"))
    b, l = tab["buses"], tab["lines"]
    types = b["type"]
    n = len(types)
    nl = [i for i, t in enumerate(types) if t == "nonlinear"]
    m = min(nl) if nl else n                                   # HG:122-125
    c = sum(t == "PV" for t in types) + 1                      # HG:127
    harmonics = np.arange(1, h_max + 1, 2)                     # HG:584
    freqs = net_freq * harmonics                               # HG:585
    I_N, Y_N = [], []
    for i in range(m, n):
        d = b["component"][i].lower()
        col = np.searchsorted(dev[d + "__freqs"], freqs)
        assert np.array_equal(dev[d + "__freqs"][col], freqs)
        if coupled:
            row = np.searchsorted(dev[d + "__Y_N_c_rowfreq"], freqs)
            I_N.append(dev[d + "__I_N_c"][col] / BASE_CURRENT)
            Y_N.append(dev[d + "__Y_N_c"][np.ix_(row, col)] / BASE_ADMITTANCE)
        else:
            I_N.append(dev[d + "__I_N_uc"][col] / BASE_CURRENT)
            Y_N.append(dev[d + "__Y_N_uc"][col] / BASE_ADMITTANCE)
    f = lambda a: np.asarray(a, dtype=float)
    return Net(n=n, m=m, c=c, harmonics=harmonics,
               line_from=np.asarray(l["fromID"]), line_to=np.asarray(l["toID"]),
               R=f(l["R"]) / BASE_IMPEDANCE, X=f(l["X"]) / BASE_IMPEDANCE,
               G=f(l["G"]) / BASE_ADMITTANCE, B=f(l["B"]) / BASE_ADMITTANCE,
               X_sh=f(b["X_sh"]) / BASE_IMPEDANCE, P=f(b["P"]) / BASE_POWER,
               Q=f(b["Q"]) / BASE_POWER, I_N=np.array(I_N), Y_N=np.array(Y_N),
               coupled=bool(coupled))
