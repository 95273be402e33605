"""Seeded synthetic scenario generator: every scenario s scales the nominal
loads and the Norton current spectrum with ``numpy.random.default_rng(s)``, drawing in the
order P factors [n], Q factors [n], |k| [q, H], angle(k) [q, H].  Host-side numpy; the
arrays are returned batch-innermost as the C ABI wants them."""
from __future__ import annotations

import numpy as np

SPREADS = {"tight": (0.1, 0.05, 0.02), "wide": (0.5, 0.2, 0.1)}


def draw_scenario(seed, P0, Q0, I_N0, spread="tight"):
    dp, dm, da = SPREADS[spread] if isinstance(spread, str) else spread
    q, H = I_N0.shape
    rng = np.random.default_rng(seed)
    P = P0 * rng.uniform(1 - dp, 1 + dp, len(P0))
    Q = Q0 * rng.uniform(1 - dp, 1 + dp, len(Q0))
    mag = rng.uniform(1 - dm, 1 + dm, (q, H))
    ang = rng.uniform(-da, da, (q, H))
    return P, Q, I_N0 * (mag * np.exp(1j * ang))


def make_batch(net, B, spread="tight", seed0=0, exact_prefix=None):
    """-> P [n, B], Q [n, B], I_N [q, H, B].  The first ``exact_prefix`` scenarios (default: all
    up to 4096) are drawn with one Generator per seed exactly like the golden fixtures; the
    rest use one vectorised stream (seed = seed0 + 10**9) - identical distribution, and the
    whole batch is a pure function of (B, spread, seed0)."""
    n, q, H = net.n, net.q, net.H
    dp, dm, da = SPREADS[spread] if isinstance(spread, str) else spread
    P = np.empty((n, B)); Q = np.empty((n, B)); I_N = np.empty((q, H, B), dtype=np.complex128)
    k = min(B, 4096 if exact_prefix is None else exact_prefix)
    for s in range(k):
        P[:, s], Q[:, s], I_N[:, :, s] = draw_scenario(seed0 + s, net.P, net.Q, net.I_N, spread)
    if B > k:
        rng = np.random.default_rng(seed0 + 10 ** 9)
        r = B - k
        P[:, k:] = net.P[:, None] * rng.uniform(1 - dp, 1 + dp, (n, r))
        Q[:, k:] = net.Q[:, None] * rng.uniform(1 - dp, 1 + dp, (n, r))
        mag = rng.uniform(1 - dm, 1 + dm, (q, H, r))
        ang = rng.uniform(-da, da, (q, H, r))
        I_N[:, :, k:] = net.I_N[:, :, None] * (mag * np.exp(1j * ang))
    return P, Q, I_N
