"""Prototype: Newton step through the constant complex current-balance operator A."""
import sys, glob, os
sys.path.insert(0, "oracle")
import numpy as np, hpf_oracle as O
G = "tests/golden"

class Structured:
    def __init__(self, net, Y):
        n, m, c, H = net.n, net.m, net.c, net.H
        nH = n*H
        A = np.zeros((nH, nH), complex)
        for h in range(H):
            A[h*n:(h+1)*n, h*n:(h+1)*n] = Y[h]
        for k in range(n-m):
            rows = np.arange(H)*n + m + k
            if net.coupled:
                A[np.ix_(rows, rows)] -= net.Y_N[k]
            else:
                A[rows, rows] -= net.Y_N[k]
        self.A = A
        Z = np.arange(m, nH)          # stacked indices with current rows
        F = np.arange(0, m)           # fundamental linear buses
        self.Z, self.F = Z, F
        AZZ = A[np.ix_(Z, Z)]
        self.condA = np.linalg.cond(AZZ)
        self.Ainv = np.linalg.inv(AZZ)
        self.G = self.Ainv @ A[np.ix_(Z, F)]      # |Z| x m
        self.net, self.Y = net, Y

    def step(self, P, Q, V_m, V_a, f):
        """returns dx (N) such that x_new = x - dx, same ordering as reference."""
        net = self.net; n, m, c, H = net.n, net.m, net.c, net.H
        nH = n*H; q = n-m
        V = (V_m*np.exp(1j*V_a)).ravel(); E = np.exp(1j*V_a).ravel()
        Vm = V_m.ravel()
        # unpack f -> complex f_c (nH-1 entries e <-> s=e+1)
        fr = f[:nH-1]; fi = np.zeros(nH-1); fi[c-1:] = f[nH-1:]
        fc = fr + 1j*fi
        fI = fc[m-1:]                 # rows s=m..nH-1   (|Z|)
        fS = fc[:m-1]                 # power rows s=1..m-1
        u0 = -(self.Ainv @ fI)        # |Z|
        # fundamental unknowns of linear buses: theta_i (i=1..m-1), Vm_i (i=c..m-1)
        # u_F[i] = E_i*(dVm_i + j Vm_i dth_i); T_F maps dxF -> u_F (complex m-vector, slack=0)
        nth, nvm = m-1, m-c
        TF = np.zeros((m, nth+nvm), complex)
        for i in range(1, m): TF[i, i-1] = 1j*V[i]
        for i in range(c, m): TF[i, nth+i-c] = E[i]
        # power rows jac wrt all fundamental buses polar unknowns
        dSdA, dSdV = O._power_derivatives(V[:n], E[:n], self.Y[0])
        # dS = dSdA dth + dSdV dVm ; for NL fundamental buses: dth = Im(conj(E) u)/Vm, dVm = Re(conj(E) u)
        # u_Z1 = u0_Z1 - G_Z1 TF dxF
        GZ1 = self.G[:q, :]           # rows of Z for s=m..n-1
        u0Z1 = u0[:q]
        # complex rows i=1..m-1
        rowsS = np.arange(1, m)
        # contributions
        def polar_of(u, idx):  # u complex for fundamental buses idx -> (dth, dVm)
            w = np.conj(E[idx])*u
            return w.imag/Vm[idx], w.real
        nl = np.arange(m, n)
        # build small system M dxF = rhs  (complex rows -> real rows: Re for i>=1, Im for i>=c)
        M_c = dSdA[np.ix_(rowsS, np.arange(1, m))] @ np.eye(nth, nth+nvm) if nth else np.zeros((0, nth+nvm))
        M_c = np.zeros((m-1, nth+nvm), complex)
        M_c[:, :nth] = dSdA[np.ix_(rowsS, np.arange(1, m))]
        M_c[:, nth:] = dSdV[np.ix_(rowsS, np.arange(c, m))]
        rhs_c = -fS.copy()
        if q:
            # linear map from u_Z1 to dS: dS_i = sum_k dSdA[i,nl_k]*dth_k + dSdV[i,nl_k]*dVm_k
            # dth_k = Im(conj(E_k) u_k)/Vm_k ; dVm_k = Re(conj(E_k) u_k)
            # u_k = u0_k - (GZ1 TF dxF)_k
            W = GZ1 @ TF                # q x nx complex
            dth0, dvm0 = polar_of(u0Z1, nl)
            rhs_c -= dSdA[np.ix_(rowsS, nl)] @ dth0 + dSdV[np.ix_(rowsS, nl)] @ dvm0
            cE = np.conj(E[nl])[:, None]*W
            dthW, dvmW = cE.imag/Vm[nl][:, None], cE.real
            M_c -= dSdA[np.ix_(rowsS, nl)] @ dthW + dSdV[np.ix_(rowsS, nl)] @ dvmW
        Mr = np.vstack([M_c.real, M_c.imag[c-1:]])
        rr = np.concatenate([rhs_c.real, rhs_c.imag[c-1:]])
        dxF = np.linalg.solve(Mr, rr) if len(rr) else np.zeros(0)
        uZ = u0 - self.G @ (TF @ dxF)
        # assemble full increments (Newton: x_new = x + delta; reference: x_new = x - dx -> dx = -delta)
        dth = np.zeros(nH); dvm = np.zeros(nH)
        dth[1:m] = dxF[:nth]; dvm[c:m] = dxF[nth:]
        w = np.conj(E[self.Z])*uZ
        dth[self.Z] = w.imag/Vm[self.Z]; dvm[self.Z] = w.real
        delta = np.concatenate([dth[1:], dvm[c:]])
        return -delta

def hpf_structured(net, P, Q, I_N, Y, S, thresh_h=1e-4, max_iter_h=50):
    n, c, H = net.n, net.c, net.H; nH = n*H
    V_m, V_a, _, nf = O.pf(net, Y, P, Q, solver="lapack")
    f, err = O.harmonic_mismatch(net, P, Q, V_m, V_a, Y, I_N)
    x = np.append(V_a.ravel()[1:], V_m.ravel()[c:])
    it = 0
    while err > thresh_h and it < max_iter_h:
        x = x - S.step(P, Q, V_m, V_a, f)
        V_a.ravel()[1:] = x[:nH-1]; V_m.ravel()[c:] = x[nH-1:]
        f, err = O.harmonic_mismatch(net, P, Q, V_m, V_a, Y, I_N)
        it += 1
    Vm2, Va2 = O.postprocess(V_m, V_a)
    return Vm2, Va2, it, err

if __name__ == "__main__":
    # 1) single-step check vs dense solve
    for p in sorted(glob.glob(G+"/case_*.npz")):
        d = np.load(p)
        net = O.net_from_golden(G, str(d["net"]), int(d["h_max"]), bool(d["coupled"]))
        Y = O.build_admittance_matrices(net)
        S = Structured(net, Y)
        Vm, Va = d["V_fund_m"].copy(), d["V_fund_a"].copy()
        f, err = O.harmonic_mismatch(net, net.P, net.Q, Vm, Va, Y, net.I_N)
        dx = S.step(net.P, net.Q, Vm, Va, f)
        x0 = np.append(Va.ravel()[1:], Vm.ravel()[net.c:])
        x1 = x0 - dx
        Vm2, Va2, it, e = hpf_structured(net, net.P, net.Q, net.I_N, Y, S)
        V = Vm2*np.exp(1j*Va2); Vg = d["V_m"]*np.exp(1j*d["V_a"])
        print(os.path.basename(p), "cond(A_ZZ)=%.2e" % S.condA, "x1 diff", abs(x1-d["x1"]).max(), "iters", it, int(d["n_iter_h"]), "V rel", (abs(V-Vg)/abs(Vg)).max())
    for p in sorted(glob.glob(G+"/scen_*.npz")):
        d = np.load(p)
        net = O.net_from_golden(G, str(d["net"]), int(d["h_max"]), bool(d["coupled"]))
        Y = O.build_admittance_matrices(net); S = Structured(net, Y)
        n = len(d["seed"]); mism = 0; rel = []
        for s in range(n):
            Vm2, Va2, it, e = hpf_structured(net, d["P"][s], d["Q"][s], d["I_N"][s], Y, S)
            mism += it != d["n_iter_h"][s]
            V = Vm2*np.exp(1j*Va2); Vg = d["V_m"][s]*np.exp(1j*d["V_a"][s])
            rel.append((abs(V-Vg)/abs(Vg)).max())
        rel = np.array(rel)
        Vl = d["V_m_lapack"]*np.exp(1j*d["V_a_lapack"]); Vg = d["V_m"]*np.exp(1j*d["V_a"])
        floor = (abs(Vl-Vg)/abs(Vg)).reshape(n,-1).max(1)
        print(os.path.basename(p), "iter mismatches", mism, "/", n, "(ref floor %d)" % (d["n_iter_h"]!=d["n_iter_h_lapack"]).sum(),
              "rel>1e-9:", (rel>1e-9).sum(), "(floor %d)" % (floor>1e-9).sum(), "max %.2e median %.2e" % (rel.max(), np.median(rel)))
