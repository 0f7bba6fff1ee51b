"""50-digit mpmath restatement of MODEL_SPEC.md §1-6.  TEST INFRASTRUCTURE ONLY (parity unpinned, see
oracle/foct_oracle.h): an implementation independent of both the C oracle and the CUDA kernels, used by
tests/golden/make_golden.py to produce the committed golden vectors.  The gradient is NOT the analytic
formula of MODEL_SPEC §5 but a high-order numerical derivative of the log density at 50 digits, so the
golden gradient checks the analytic derivation as well as its implementation.
"""
from __future__ import annotations

import mpmath as mp

mp.mp.dps = 50


def grid(Nn, gridType):
    dx = mp.mpf(1) / (Nn + 1)
    lo, hi = (dx / 2, 1 - dx / 2) if gridType == 0 else (mp.mpf(0), mp.mpf(1))
    return [lo if Nn == 1 else lo + (hi - lo) * k / (Nn - 1) for k in range(Nn)]


def kern(a, b, rho, kernel):
    d = a - b
    return mp.exp(-(d * d) / (2 * rho * rho if kernel == 0 else rho * rho))


def basis(x, Nn, gridType, rho, kernel, jitter):
    """B[k][i], exact to working precision (LU solve of the jittered Gram matrix)."""
    x = [mp.mpf(v) for v in x]
    rho, jitter = mp.mpf(rho), mp.mpf(jitter)
    xg = grid(Nn, gridType)
    xmin, xmax = min(x), max(x)
    K = mp.matrix(Nn, Nn)
    for i in range(Nn):
        for j in range(Nn):
            K[i, j] = kern(xg[i], xg[j], rho, kernel) + (jitter if i == j else 0)
    Kinv = K ** -1
    B = [[None] * len(x) for _ in range(Nn)]
    for i, xi in enumerate(x):
        xp = (xi - xmin) / (xmax - xmin)
        kv = [kern(xp, xg[k], rho, kernel) for k in range(Nn)]
        for k in range(Nn):
            B[k][i] = mp.fsum(kv[j] * Kinv[j, k] for j in range(Nn))
    return B


def logp(case, q, B):
    """lp(q) per MODEL_SPEC §4 and chi2 (numerator of br).  case: dict of plain Python numbers/lists."""
    kind, Nn = case["kind"], case["Nn"] if case["kind"] == 0 else 0
    sp = case["spec"]
    x, y, uy = case["x"], case["y"], case["uy"]
    N = len(x)
    c = mp.mpf(case["dataType"])
    th1, th2, th3 = q[0], q[1], q[2]
    yg = q[3:3 + Nn]
    gp = kind == 0
    qlam = q[3 + Nn] if gp else mp.mpf(0)
    qsig = q[4 + Nn] if gp else mp.mpf(0)
    lam, sig = mp.exp(qlam), mp.exp(qsig)
    lp = mp.mpf(0)
    chi2 = mp.mpf(0)
    if not case["prior_PD"]:
        szz = mp.mpf(0)
        for i in range(N):
            dl = mp.fsum(B[k][i] * yg[k] for k in range(Nn)) if Nn else mp.mpf(0)
            if sp["modulation"] == 0:
                m = th1 + th2 * mp.exp(-c * mp.mpf(x[i]) / (th3 * (1 + dl)))
            else:
                m = th1 + th2 * mp.exp(-c * mp.mpf(x[i]) / th3) * (1 + dl)
            r = mp.mpf(y[i]) - m
            z = r / (sig * mp.mpf(uy[i]))
            szz += z * z
            chi2 += (r / mp.mpf(uy[i])) ** 2
        lp += -szz / 2 - N * qsig - mp.fsum(mp.log(mp.mpf(u)) for u in uy)
    if sp["theta_prior"] == 0:
        S = mp.matrix(3, 3)
        for a in range(3):
            for b in range(3):
                S[a, b] = mp.mpf(case["Sigma0"][a * 3 + b])
        d = mp.matrix([th1 - mp.mpf(case["theta0"][0]), th2 - mp.mpf(case["theta0"][1]), th3 - mp.mpf(case["theta0"][2])])
        lp += -(d.T * (S ** -1) * d)[0, 0] / 2
    if gp:
        if sp["ygp_prior"] == 0:
            lp += -Nn * qlam - mp.fsum(v * v for v in yg) / (2 * lam * lam)
        else:
            lp += -Nn * qlam - mp.fsum(abs(v) for v in yg) / lam
        rate = mp.mpf(case["lambda_rate"])
        lp += (qlam - rate * lam) if sp["lambda_prior"] == 0 else (-rate * lam)
        if sp["sigma_sd"] > 0:
            lp += -((sig - mp.mpf(sp["sigma_mean"])) / mp.mpf(sp["sigma_sd"])) ** 2 / 2
        lp += qlam + qsig
    return lp, chi2


def logp_grad(case, q, B):
    q = [mp.mpf(v) for v in q]
    lp, chi2 = logp(case, q, B)
    g = []
    for d in range(len(q)):
        def f(t, d=d):
            qq = list(q)
            qq[d] = t
            return logp(case, qq, B)[0]
        g.append(mp.diff(f, q[d], h=mp.mpf(10) ** -12 * max(1, abs(q[d]))))
    return lp, g, chi2
