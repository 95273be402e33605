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


def make_batch(net, B, spread="tight", seed0=0):
    """-> P [n, B], Q [n, B], I_N [q, H, B]: scenario b is ``draw_scenario(seed0 + b, ...)`` - ONE
    generator for every consumer (golden fixtures, tests, bench GPU arm, bench CPU arm), so a
    scenario index means the same inputs everywhere (about 25 us per scenario on the host)."""
    n, (q, H) = len(net.P), np.shape(net.I_N)
    P = np.empty((n, B)); Q = np.empty((n, B)); I_N = np.empty((q, H, B), dtype=np.complex128)
    for s in range(B):
        P[:, s], Q[:, s], I_N[:, :, s] = draw_scenario(seed0 + s, net.P, net.Q, net.I_N, spread)
    return P, Q, I_N
