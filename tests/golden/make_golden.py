"""Generates tests/golden/model_cases.json from the 50-digit mpmath restatement (oracle/mp_model.py).

PARITY UNPINNED: the reference holds no golden vectors for this path and FitOCTLib / rstan
cannot run here, so these vectors pin the oracle and the CUDA kernels to MODEL_SPEC.md, not to rstan.
Run:  python tests/golden/make_golden.py      (about a minute; output committed)
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fitoct_b200 import synth  # noqa: E402
from oracle import mp_model as M  # noqa: E402

DEFAULT_SPEC = dict(modulation=0, kernel=0, jitter=1e-9, ygp_prior=0, lambda_prior=0, sigma_mean=1.0, sigma_sd=0.1,
                    theta_prior=0, br_ndf=0)


def synth_case(name, kind, N_keep, Nn, gridType, dataType, spec_over, prior_PD, rho, seed, n_q=2, mod_kind=1):
    rng = np.random.default_rng(seed)
    S = synth.make_profiles(5)
    sel = np.sort(rng.choice(481, size=N_keep, replace=False)) if N_keep < 481 else np.arange(481)
    x = S["x"][sel]
    y = S["Y"][mod_kind][sel]
    uy = S["UY"][mod_kind][sel] * (1.0 + 0.1 * rng.random(N_keep))
    th0 = S["theta0"][mod_kind].copy()
    if dataType == 1:
        th0[2] /= 2.0
    A = rng.standard_normal((3, 3))
    Sig = np.diag((0.05 * th0) ** 2) + 0.1 * (A @ A.T)  # non-diagonal on purpose
    spec = dict(DEFAULT_SPEC, **spec_over)
    if kind == 1:
        spec["theta_prior"] = spec_over.get("theta_prior", 1)
    D = Nn + 5 if kind == 0 else 3
    qs = []
    for _ in range(n_q):
        q = np.zeros(D)
        q[:3] = th0 * (1 + 0.02 * rng.standard_normal(3))
        if kind == 0:
            q[3:3 + Nn] = 0.05 * rng.standard_normal(Nn)
            q[3 + Nn] = np.log(0.1) + 0.3 * rng.standard_normal()
            q[4 + Nn] = 0.2 * rng.standard_normal()
        qs.append(q.tolist())
    return dict(name=name, kind=kind, x=x.tolist(), y=y.tolist(), uy=uy.tolist(), dataType=dataType, Nn=Nn,
                gridType=gridType, rho=rho, lambda_rate=0.1, theta0=th0.tolist(), Sigma0=Sig.reshape(9).tolist(),
                prior_PD=prior_PD, spec=spec, q=qs)


def main():
    cases = [
        synth_case("expgp_default_N481_Nn10", 0, 481, 10, 0, 2, {}, 0, 0.1, 1, n_q=2),
        synth_case("expgp_switches_N50_Nn5", 0, 50, 5, 1, 1,
                   dict(modulation=1, kernel=1, ygp_prior=1, lambda_prior=1, sigma_sd=0.0, br_ndf=1), 0, 0.2, 2),
        synth_case("expgp_N200_Nn20_extremal", 0, 200, 20, 1, 2, {}, 0, 0.05, 3, n_q=1, mod_kind=3),
        synth_case("expgp_priorPD_Nn7", 0, 33, 7, 0, 2, {}, 1, 1.0 / 7, 4),
        synth_case("expgp_Nn1_N97", 0, 97, 1, 0, 2, {}, 0, 1.0, 5, n_q=1),
        synth_case("monoexp_flat_N481", 1, 481, 0, 0, 2, {}, 0, 1.0, 6, mod_kind=0),
        synth_case("monoexp_mvn_amp_N64", 1, 64, 0, 0, 1, dict(theta_prior=0, modulation=1), 0, 1.0, 7, mod_kind=0),
    ]
    out = []
    for c in cases:
        Nn = c["Nn"] if c["kind"] == 0 else 0
        B = M.basis(c["x"], Nn, c["gridType"], c["rho"], c["spec"]["kernel"], c["spec"]["jitter"]) if Nn else []
        # the double-rounded basis is what the fp64 implementations are fed for the lp/grad check
        B64 = [[float(v) for v in row] for row in B]
        Bmp = [[mp.mpf(v) for v in row] for row in B64]
        res = []
        for q in c["q"]:
            lp, g, chi2 = M.logp_grad(c, q, Bmp)
            res.append(dict(lp=mp.nstr(lp, 25), grad=[mp.nstr(v, 25) for v in g], chi2=mp.nstr(chi2, 25)))
        c["basis"] = B64
        c["expected"] = res
        out.append(c)
        print(c["name"], "done", flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_cases.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
