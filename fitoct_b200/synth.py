"""synthData.R-shaped synthetic depth profiles (the workload of every BASELINE.json config).

Follows /root/reference/synthData.R:3-11 (grid, truth, noise law) and :21,35,49,63 (the four sinc
modulations); SURVEY.md §8(d) fixes the seeding (the reference sets no seed in that file).
"""
from __future__ import annotations

import numpy as np

A_TRUE, B_TRUE, L0_TRUE, S_NOISE = 1000.0, 2000.0, 150.0, 0.5  # synthData.R:4-7
SEED_KEY = 20181120  # date of the reference UI release (ShinyInterface/ui.R:515-516)


def depth_grid() -> np.ndarray:
    return np.arange(20.0, 501.0)  # synthData.R:3  x = 20:500, N = 481


def modulation(kind: int, x: np.ndarray) -> np.ndarray:
    """kind 0: none (monoExp); 1..4: synthData.R:21,35,49,63."""
    if kind == 0:
        return np.zeros_like(x)
    if kind == 1:
        return 10.0 * np.sin(x / 50.0) / x
    if kind == 2:
        return 10.0 * np.sin(x / 25.0) / x
    if kind == 3:
        return 10.0 * np.sin(x / 75.0) / x
    if kind == 4:
        return np.sin((x - 250.0) / 20.0) / (x - 250.0 + 0.1)
    raise ValueError(kind)


def make_profiles(n: int, *, first_id: int = 0, modulated_only: bool = False, fitted_uy: bool = False,
                  ru_theta: float = 0.05):
    """Returns dict(x[N], Y[n,N], UY[n,N], theta0[n,3], Sigma0[n,3,3], mod_kind[n], ids[n]).

    Profile j (global id) draws its noise from numpy Philox(key=(SEED_KEY, j)) so that a profile
    does not depend on the batch it is generated in.  Noise sd uses y0 (the unmodulated curve), as
    synthData.R:23,37,51,65 does.
    """
    x = depth_grid()
    N = x.size
    y0 = A_TRUE + B_TRUE * np.exp(-x / L0_TRUE)
    sd = S_NOISE * np.sqrt(y0 - A_TRUE + 1.0)  # synthData.R:11
    ids = np.arange(first_id, first_id + n, dtype=np.int64)
    kinds = (ids % 4 + 1) if modulated_only else (ids % 5)
    Y = np.empty((n, N))
    th0 = np.empty((n, 3))
    truth = np.array([A_TRUE, B_TRUE, 2.0 * L0_TRUE])  # dataType = 2  =>  theta3 = 2*l0
    for r, j in enumerate(ids):
        g = np.random.Generator(np.random.Philox(key=np.array([SEED_KEY, int(j)], dtype=np.uint64)))
        m = modulation(int(kinds[r]), x)
        Y[r] = A_TRUE + B_TRUE * np.exp(-x / (L0_TRUE * (1.0 + m))) + sd * g.standard_normal(N)
        th0[r] = truth * (1.0 + 0.01 * g.standard_normal(3))  # stand-in for the MAP estimate fed as theta0
    if fitted_uy:
        # the estimateNoise form uy = a_1 exp(-x/a_2) (ShinyInterface/ui.R:81), fitted to the true sd law
        a2 = 2.0 * L0_TRUE
        a1 = S_NOISE * np.sqrt(B_TRUE)
        UY = np.broadcast_to(np.maximum(a1 * np.exp(-x / a2), S_NOISE), (n, N)).copy()
    else:
        UY = np.broadcast_to(sd, (n, N)).copy()
    S0 = np.zeros((n, 3, 3))
    for k in range(3):
        S0[:, k, k] = (ru_theta * th0[:, k]) ** 2  # FitOCT.R:46 ru_theta
    return dict(x=x, Y=Y, UY=UY, theta0=th0, Sigma0=S0, mod_kind=kinds, ids=ids)
