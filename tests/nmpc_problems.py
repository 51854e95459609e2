"""Seeded synthetic NMPC instances in the reference's parameter layout
(src/mpc/mpc_generator.py:71-79,93-104; assembled at src/path_generator.py:378-379)."""
import math

import numpy as np

DEFAULT_WEIGHTS = [0.0, 10.0, 0.0, 0.0, 0.0, 0.0, 0.0, 200.0, 10.0, 5.0]   # configs/default.yaml:21-31 in z0[10:20] order
SMOOTH_WEIGHTS = [1.0, 10.0, 0.0, 0.0, 0.0, 5.0, 0.2, 20.0, 8.0, 20.0]    # configs/smooth_velocity.yaml:21-29 (+qv=10)
MIXED_WEIGHTS = [1.0, 10.0, 0.1, 0.1, 0.2, 5.0, 0.2, 20.0, 8.0, 20.0]     # every term of the cost active


def param_len(N, Nobs, Nd):
    return 20 + N + 3 * Nobs + 5 * Nd * N + 3 * N


def synth(N, Nobs, Nd, B, seed=0, active=True, weights=None):
    """Random pose, a piecewise-linear reference sampled every 0.33 m (rough_ref spacing,
    src/mpc/mpc_generator.py:20), circles near the path, one moving ellipse near the path
    (if `active`) and phantom unit discs at the origin in the other dynamic slots
    (src/path_generator.py:274-280)."""
    rng = np.random.default_rng(seed)
    P = np.zeros((B, param_len(N, Nobs, Nd)))
    for b in range(B):
        p = P[b]
        x0, y0 = rng.uniform(2, 50, 2)
        th0 = rng.uniform(-math.pi, math.pi)
        pts = [(x0 + rng.normal(0, .2), y0 + rng.normal(0, .2))]
        hd = th0 + rng.normal(0, .5)
        for _ in range(N):
            if rng.random() < 0.15:
                hd += rng.uniform(-1.5, 1.5)
            pts.append((pts[-1][0] + 0.33 * math.cos(hd), pts[-1][1] + 0.33 * math.sin(hd)))
        p[0:3] = [x0, y0, th0]
        p[3:5] = [rng.uniform(0, 1.5), rng.uniform(-.5, .5)]
        p[5:8] = [pts[N][0], pts[N][1], hd]
        p[8:10] = p[3:5]
        if weights is None:
            p[10:20] = DEFAULT_WEIGHTS if rng.random() < 0.5 else MIXED_WEIGHTS
        else:
            p[10:20] = weights
        p[20:20 + N] = 1.5
        bc = 20 + N
        nreal = rng.integers(0, min(Nobs, 4) + 1) if Nobs else 0
        for k in range(nreal):
            j = rng.integers(0, N)
            off = rng.normal(0, 0.5, 2) if active else rng.normal(3, 0.1, 2)
            p[bc + 3 * k:bc + 3 * k + 3] = [pts[j][0] + off[0], pts[j][1] + off[1], 0.5]
        be = bc + 3 * Nobs
        for k in range(Nd):
            for t in range(N):
                e = p[be + k * 5 * N + 5 * t: be + k * 5 * N + 5 * t + 5]
                if active and k == 0:
                    j = min(N - 1, t + 2)
                    e[:] = [pts[j][0] + 0.3, pts[j][1] - 0.2, 0.7, 1.1, 0.4]
                else:
                    e[:] = [0, 0, 1, 1, 0]
        br = be + 5 * Nd * N
        for i in range(N):
            p[br + 3 * i:br + 3 * i + 3] = [pts[i][0], pts[i][1], hd]
    return P


def random_controls(N, B, seed=0, vmin=-0.5, vmax=1.5, wmax=0.5):
    rng = np.random.default_rng(seed + 1000)
    U = np.zeros((B, 2 * N))
    U[:, 0::2] = rng.uniform(vmin, vmax, (B, N))
    U[:, 1::2] = rng.uniform(-wmax, wmax, (B, N))
    return U
