"""CPU oracle for the Norton-equivalent extraction: numpy restatement of the
reference's ``Circuit Simulation/NE_from_sim.py`` (``NE`` below) on plain arrays.

TEST INFRASTRUCTURE ONLY (same rules as ``hpf_oracle.py``).  Parity status: PINNED -
``oracle/make_golden_ne.py`` runs the reference's own script (``oracle/ref_exec_ne.py``) on the two
simulation result files that ship with the reference and commits inputs + outputs under
``tests/golden/ne_extract_*.npz``; ``tests/test_oracle.py`` checks this restatement against them.

Measurement set of one device (what NE:33-84 collect from the .mat file), N = 1 + K frequencies
``freq = [50, f_1 .. f_K]``:
  Vf   (2,)       fundamental supply phasor of the two fundamental-only measurements (NE:104-105)
  Vh   (2, K)     harmonic supply phasor, magnitude set 1 / 2, per supply frequency (NE:91-92)
  I_f  (2, N)     injected-current spectrum of the two fundamental-only measurements
  I_h  (2, K, N)  spectrum with the harmonic source at f_k, magnitude set 1 / 2
"""
from __future__ import annotations

import numpy as np


def uncoupled(Vf, Vh, I_f, I_h):
    """Thunberg two-point Norton equivalents per frequency (NE:86-118).  -> Y_N_uc (N,), I_N_uc (N,)."""
    K = Vh.shape[1]
    idx = np.arange(K)
    dI_h = I_h[1][idx, idx + 1] - I_h[0][idx, idx + 1]            # NE:89-90 + np.diag (NE:96)
    Y_h = dI_h / (Vh[0] - Vh[1])                                   # NE:96
    I_h_src = Y_h * Vh[0] + I_h[0][idx, idx + 1]                   # NE:99
    dI_f = I_f[1][0] - I_f[0][0]                                   # NE:102-103
    Y_f = dI_f / (Vf[0] - Vf[1])                                   # NE:109
    I_f_src = Y_f * Vf[0] + I_f[0][0]                              # NE:111
    return np.r_[Y_f, Y_h], np.r_[I_f_src, I_h_src]               # NE:114-115


def measurement_matrix(Vf, Vh):
    """-V_mes of NE:145-154: rows = measurements (fund. 1, harmonic k = 1..K, fund. 2), columns =
    supply voltage at freq[0..N-1] and the constant 1 that multiplies I_N."""
    K = Vh.shape[1]
    N = K + 1
    V = np.zeros((N + 1, N + 1), dtype=complex)
    V[:, N] = -1.0                                                 # NE:146
    V[:, 0] = Vf[0]                                                # NE:147
    V[N, 0] = Vf[1]                                                # NE:149
    for k in range(K):
        V[k + 1, k + 1] = Vh[0][k]                                 # NE:151-152
    return -V                                                      # NE:153


def coupled(Vf, Vh, I_f, I_h):
    """Almeida coupled Norton equivalent (NE:141-173).  -> Y_N_c (N, N) [row = harmonic of the
    current, column = harmonic of the voltage], I_N_c (N,)."""
    K = Vh.shape[1]
    N = K + 1
    M = measurement_matrix(Vf, Vh)
    I_c = np.vstack([I_f[0][None, :], I_h[0], I_f[1][None, :]])    # NE:160-167
    YI = np.linalg.inv(M).dot(I_c)                                 # NE:154,169
    return YI[:N].T.copy(), YI[N].copy()                           # NE:172-173


def forward(Y_N_c, I_N_c, V):
    """I_inj = I_N - Y_N V (HG:313-323): the model the extraction inverts."""
    return I_N_c - Y_N_c @ V


def synth_measurements(Y_N_c, I_N_c, Vf, Vh):
    """Measurement set a device with the given coupled NE would produce (round-trip tests)."""
    N = len(I_N_c)
    K = N - 1
    I_f = np.empty((2, N), dtype=complex)
    I_h = np.empty((2, K, N), dtype=complex)
    for a in range(2):
        v = np.zeros(N, dtype=complex); v[0] = Vf[a]
        I_f[a] = forward(Y_N_c, I_N_c, v)
        for k in range(K):
            v = np.zeros(N, dtype=complex); v[0] = Vf[0]; v[k + 1] = Vh[a][k]
            I_h[a, k] = forward(Y_N_c, I_N_c, v)
    return I_f, I_h
