"""GPU parity of method = 'vb' (Stan's mean-field ADVI, MODEL_SPEC §14) against the CPU oracle through the C ABI.
Both sides draw from the same Philox sites, so the iterates coincide up to fp64 rounding (exp/log of libm vs CUDA, FMA
contraction, reduction order); the stochastic-gradient map is contractive near the optimum, so the difference stays small."""
import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth

pytestmark = pytest.mark.gpu


def _batch(n, Nn=10, **kw):
    S = synth.make_profiles(n, modulated_only=True)
    return S, abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, **kw)


@pytest.mark.parametrize("Nn", [5, 10])
def test_vb_matches_oracle(L, O, Nn):
    n = 4
    S, batch = _batch(n, Nn)
    spec = abi.default_spec(abi.FOCT_EXPGP)
    cfg = abi.default_vb_cfg(omega0=-3.0, output_samples=50, seed=77)
    g = L.vb(abi.FOCT_EXPGP, batch, n, spec, cfg)
    o = O.vb(abi.FOCT_EXPGP, batch, n, spec, cfg)
    assert g["status"].tolist() == o["status"].tolist() == [0] * n
    assert g["iters"].tolist() == o["iters"].tolist()
    assert np.array_equal(g["eta"], o["eta"])
    sd = np.exp(o["omega"])
    assert np.max(np.abs(g["mu"] - o["mu"]) / sd) < 1e-6          # in units of the approximation's own sd
    assert np.max(np.abs(g["omega"] - o["omega"])) < 1e-6
    assert np.allclose(g["elbo"], o["elbo"], rtol=1e-9)
    assert np.allclose(g["mean"], o["mean"], rtol=1e-7, atol=1e-9)
    assert np.allclose(g["draws"], o["draws"], rtol=1e-6, atol=1e-8)
    assert np.all(g["draws"][..., -1] == 0.0)                       # lp__ = 0, as Stan writes ADVI output


def test_vb_failure_mode_and_fixed_eta(L, O):
    """Stan's start (omega = 0) leaves the model's domain: status 2 on both sides; adapt_engaged = 0 uses eta as given."""
    S, batch = _batch(2)
    spec = abi.default_spec(abi.FOCT_EXPGP)
    g = L.vb(abi.FOCT_EXPGP, batch, 2, spec, abi.default_vb_cfg(), draws=False)
    o = O.vb(abi.FOCT_EXPGP, batch, 2, spec, abi.default_vb_cfg(), draws=False)
    assert g["status"].tolist() == o["status"].tolist() == [2, 2]
    assert g["iters"].tolist() == o["iters"].tolist() and np.array_equal(g["eta"], o["eta"], equal_nan=True)
    cfg = abi.default_vb_cfg(omega0=-4.0, adapt_engaged=0, eta=0.1, iter=400, output_samples=0)
    g = L.vb(abi.FOCT_EXPGP, batch, 2, spec, cfg, draws=False)
    o = O.vb(abi.FOCT_EXPGP, batch, 2, spec, cfg, draws=False)
    assert g["iters"].tolist() == o["iters"].tolist() and np.all(g["eta"] == 0.1)
    assert np.max(np.abs(g["mu"] - o["mu"]) / np.exp(o["omega"])) < 1e-6
    from fitoct_b200._lib import FitOCTError
    with pytest.raises(FitOCTError):
        L.vb(abi.FOCT_EXPGP, batch, 2, spec, abi.default_vb_cfg(elbo_samples=0))


def test_vb_monoexp_and_api(L, O):
    n = 3
    S = synth.make_profiles(n)
    batch = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=0)
    spec = abi.default_spec(abi.FOCT_MONOEXP)
    cfg = abi.default_vb_cfg(omega0=-3.0, output_samples=20, seed=5)
    g = L.vb(abi.FOCT_MONOEXP, batch, n, spec, cfg)
    o = O.vb(abi.FOCT_MONOEXP, batch, n, spec, cfg)
    assert g["iters"].tolist() == o["iters"].tolist() and g["status"].tolist() == o["status"].tolist()
    assert np.max(np.abs(g["mu"] - o["mu"]) / np.exp(o["omega"])) < 1e-6
    # the reference-shaped call: fitExpGP(..., method = 'vb') (FitOCT.R:42)
    from fitoct_b200 import api
    Sm = synth.make_profiles(1, modulated_only=True)
    fit = api.fitExpGP(Sm["x"], Sm["Y"][0], Sm["UY"][0], dataType=2, Nn=10, gridType="internal", method="vb",
                       theta0=Sm["theta0"][0], Sigma0=Sm["Sigma0"][0], lambda_rate=0.1, rho_scale=0)
    assert fit["method"] == "vb" and fit["fit"].vb["converged"]
    ex = fit["fit"].extract(["theta", "sigma"])
    assert ex["theta"].shape == (1000, 3) and abs(ex["theta"][:, 2].mean() - 300) < 15
    # sanity against the sampler's posterior (ADVI stopped at tol_rel_obj = 0.01 is a rough answer: that is the method)
    ref = api.fitExpGP(Sm["x"], Sm["Y"][0], Sm["UY"][0], dataType=2, Nn=10, gridType="internal", method="sample",
                       theta0=Sm["theta0"][0], Sigma0=Sm["Sigma0"][0], lambda_rate=0.1, rho_scale=0, nb_warmup=300, nb_iter=800)
    tab = ref["fit"].summary_table
    z = (fit["fit"].vb["mean"][:13] - tab[:13, 0]) / tab[:13, 2]
    assert np.max(np.abs(z)) < 6.0 and np.median(np.abs(z)) < 1.5
    with pytest.raises(RuntimeError):
        api.fitExpGP(Sm["x"], Sm["Y"][0], Sm["UY"][0], Nn=10, method="vb", theta0=Sm["theta0"][0], Sigma0=Sm["Sigma0"][0],
                     control=dict(omega0=0.0))


def test_vb_edge_configs(L, O):
    """Fewer iterations than one ELBO evaluation: the loop ends at the limit with no estimate; grad_samples > 1; no draws."""
    S, batch = _batch(2)
    spec = abi.default_spec(abi.FOCT_EXPGP)
    cfg = abi.default_vb_cfg(omega0=-3.0, iter=60, adapt_engaged=0, eta=0.1, grad_samples=3, output_samples=0)
    g = L.vb(abi.FOCT_EXPGP, batch, 2, spec, cfg, draws=False)
    o = O.vb(abi.FOCT_EXPGP, batch, 2, spec, cfg, draws=False)
    assert g["status"].tolist() == o["status"].tolist() == [1, 1] and g["iters"].tolist() == [60, 60]
    assert np.all(g["elbo"] == 0.0) and np.all(o["elbo"] == 0.0)
    assert np.max(np.abs(g["mu"] - o["mu"]) / np.exp(o["omega"])) < 1e-7
    assert np.allclose(g["mean"], o["mean"], rtol=1e-9, atol=1e-12)
