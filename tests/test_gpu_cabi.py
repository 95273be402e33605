"""The C ABI used the way INTEGRATION.md shows: raw ctypes + numpy host buffers, no torch."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_raw_ctypes_host_buffer_flow(tmp_path):
    lib = C.CDLL(os.path.join(ROOT, "harmonic_power_flow_b200", "libhpf_b200.so"))
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.hpf_create.argtypes = [C.POINTER(vp), i]
    lib.hpf_set_network.argtypes = [vp, i, i, i, i, vp, i, vp, vp, vp, vp, vp, vp, vp]
    lib.hpf_set_devices.argtypes = [vp, i, i, vp, vp]
    lib.hpf_build_Y.argtypes = [vp, vp, vp]
    lib.hpf_solve_host.argtypes = [vp, i, vp, vp, vp, d, i, d, i] + [vp] * 7
    lib.hpf_destroy.argtypes = [vp]
    lib.hpf_last_error.restype = C.c_char_p
    lib.hpf_last_error.argtypes = [vp]
    p = lambda a: a.ctypes.data_as(vp)

    g = helpers.load_case("net3_c_h25")
    net, _, _ = helpers.packed_from_files("net3", 25, True, tmp_path)
    h = vp()
    assert lib.hpf_create(C.byref(h), 0) == 0
    # call order is enforced
    assert lib.hpf_build_Y(h, None, None) == -1 and b"hpf_set_network" in lib.hpf_last_error(h)
    assert lib.hpf_set_network(h, net.n, net.m, net.c, net.H, p(net.harmonics), len(net.R), p(net.from_id),
                               p(net.to_id), p(net.R), p(net.X), p(net.G), p(net.B), p(net.X_sh)) == 0
    YN = np.ascontiguousarray(net.Y_N)
    assert lib.hpf_set_devices(h, len(net.devices), 1, p(YN), p(net.dev_of_nl_bus)) == 0
    assert lib.hpf_build_Y(h, None, None) == 0
    B = 3
    P = np.ascontiguousarray(np.repeat(net.P[:, None], B, 1))
    Q = np.ascontiguousarray(np.repeat(net.Q[:, None], B, 1))
    I_N = np.ascontiguousarray(np.repeat(net.I_N[:, :, None], B, 2))
    V_m = np.empty((net.H, net.n, B)); V_a = np.empty((net.H, net.n, B))
    I_inj = np.empty((net.q, net.H, B), complex)
    nf = np.empty(B, np.int32); nh = np.empty(B, np.int32); st = np.empty(B, np.int32); err = np.empty(B)
    rc = lib.hpf_solve_host(h, B, p(P), p(Q), p(I_N), 1e-6, 30, 1e-4, 50, p(V_m), p(V_a), p(I_inj),
                            p(nf), p(nh), p(err), p(st))
    assert rc == 0, lib.hpf_last_error(h)
    assert (nf == int(g["n_iter_f"])).all() and (nh == int(g["n_iter_h"])).all() and (st == 0).all()
    V = helpers.phasor(V_m[:, :, 1], V_a[:, :, 1])
    Vg = helpers.phasor(g["V_m"], g["V_a"])
    assert (np.abs(V - Vg) / np.abs(Vg)).max() < 1e-9
    assert lib.hpf_destroy(h) == 0
