"""CPU: the oracle's NUTS / adaptation / summary / MAP against analytic answers.

The only known-answer sampler target the reference itself holds is Tests/testGamma.R:19-47
(lambda ~ exponential(1/10) on a lower=0 parameter, mean 10, sd 10, median 10 ln 2)."""
import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth


def mcse_ok(x_chains, truth, n_eff, z=4.0):
    sd = x_chains.std(ddof=1)
    return abs(x_chains.mean() - truth) <= z * sd / np.sqrt(n_eff)


def test_independent_normal_moments(O):
    sd = np.array([1.0, 10.0, 0.1, 3.0, 0.5])
    cfg = abi.default_cfg(n_warmup=300, n_iter=1800, seed=3)
    dr, sp = O.sample_analytic(0, sd, cfg)
    S = O.summary(dr)
    for d in range(5):
        assert abs(S[d, 0]) <= 4 * sd[d] / np.sqrt(S[d, 8])          # mean within 4 MCSE
        assert abs(S[d, 2] / sd[d] - 1) < 0.06                       # sd
        assert 0.99 < S[d, 9] < 1.01                                  # split-Rhat
        assert S[d, 8] > 1500 and S[d, 10] > 1500
    assert 0.7 < sp[..., 0].mean() < 0.95                             # adapted towards delta = 0.8
    assert sp[..., 4].sum() == 0                                      # no divergences
    # the metric adapts to the scales: a well-adapted chain needs short trees
    assert sp[..., 2].mean() < 3.5


def test_testGamma_exponential_known_answer(O):
    # Tests/testGamma.R:27,35,42-47  (control adapt_delta 0.99, max_treedepth 12)
    cfg = abi.default_cfg(n_warmup=500, n_iter=5500, seed=1234, adapt_delta=0.99, max_treedepth=12)
    dr, sp = O.sample_analytic(1, np.array([0.1]), cfg)
    lam = np.exp(dr[..., 0])
    n_eff = O.summary(lam[..., None])[0, 8]
    assert n_eff > 2000
    assert abs(lam.mean() - 10.0) <= 4 * lam.std() / np.sqrt(n_eff)
    assert abs(lam.std() / 10.0 - 1) < 0.08
    assert abs(np.median(lam) - 10 * np.log(2)) < 0.4
    assert sp[..., 0].mean() > 0.95


def test_same_seed_same_chain_different_seed_differs(O):
    cfg = abi.default_cfg(n_warmup=50, n_iter=100, seed=9)
    a, _ = O.sample_analytic(0, np.ones(3), cfg)
    b, _ = O.sample_analytic(0, np.ones(3), cfg)
    cfg.seed = 10
    c, _ = O.sample_analytic(0, np.ones(3), cfg)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert not np.array_equal(a[:, 0], a[:, 1])  # chains use distinct Philox keys


def test_short_warmup_rules(O):
    # warmup < 20: no metric adaptation (step size only); warmup 100 (ctrlParams.yaml:1) -> 15/75/10 split
    S = synth.make_profiles(1, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=5)
    spec = abi.default_spec()
    out = O.sample(0, b, 1, spec, abi.default_cfg(n_warmup=10, n_iter=20, seed=1), n_threads=4)
    assert np.all(out["inv_metric"] == 1.0)
    out = O.sample(0, b, 1, spec, abi.default_cfg(n_warmup=100, n_iter=120, seed=1), n_threads=4)
    assert np.all(out["inv_metric"] != 1.0) and np.all(out["inv_metric"] > 0)
    assert np.all(np.isfinite(out["draws"]))


def numpy_ess(x):
    """Independent numpy restatement of MODEL_SPEC §9 n_eff for chains x[c, t] (FFT autocovariance)."""
    C, n = x.shape
    acov = []
    for c in range(C):
        xc = x[c] - x[c].mean()
        f = np.fft.rfft(xc, 2 * n)
        acov.append(np.fft.irfft(f * np.conj(f))[:n] / n)
    acov = np.array(acov)
    mean_var = (acov[:, 0] * n / (n - 1)).mean()
    var_plus = mean_var * (n - 1) / n + (x.mean(axis=1).var(ddof=1) if C > 1 else 0.0)
    rho = np.zeros(n + 2)
    rho_even, rho_odd = 1.0, 1 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[0], rho[1] = rho_even, rho_odd
    s = 1
    while s < n - 4 and rho_even + rho_odd > 0:
        rho_even = 1 - (mean_var - acov[:, s + 1].mean()) / var_plus
        rho_odd = 1 - (mean_var - acov[:, s + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[s + 1], rho[s + 2] = rho_even, rho_odd
        s += 2
    max_s = s
    if rho_even > 0:
        rho[max_s + 1] = rho_even
    for s in range(1, max_s - 2, 2):
        if rho[s + 1] + rho[s + 2] > rho[s - 1] + rho[s]:
            rho[s + 1] = (rho[s - 1] + rho[s]) / 2
            rho[s + 2] = rho[s + 1]
    tau = -1 + 2 * rho[: max_s + 1].sum() + rho[max_s + 1]
    return C * n / max(tau, 1.0 / np.log10(C * n))      # rstan's ess_rfun: tau_hat = max(tau_hat, 1 / log10(S))


@pytest.mark.parametrize("phi,n", [(0.0, 400), (0.7, 1000), (0.95, 777), (-0.5, 500), (0.9, 40), (-0.9, 300), (-0.97, 60)])
def test_summary_against_numpy(O, phi, n):
    rng = np.random.default_rng(int(1000 * abs(phi)) + n)
    C = 4
    x = np.zeros((n, C, 2))
    e = rng.standard_normal((n, C, 2))
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    x[..., 1] = 5.0 + 2.0 * x[..., 1] + np.arange(C)[None, :] * 0.3   # offset chains -> Rhat > 1
    S = O.summary(x)
    for p in range(2):
        col = x[..., p]
        flat = col.T.reshape(-1)
        assert np.isclose(S[p, 0], flat.mean(), rtol=1e-12)
        assert np.isclose(S[p, 2], flat.std(ddof=1), rtol=1e-12)
        np.testing.assert_allclose(S[p, 3:8], np.quantile(flat, [0.025, 0.25, 0.5, 0.75, 0.975]), rtol=1e-12)
        assert np.isclose(S[p, 8], numpy_ess(col.T), rtol=1e-9)
        h = n // 2
        halves = np.concatenate([col[:h].T, col[n - h:].T])
        W = halves.var(axis=1, ddof=1).mean()
        Bv = halves.mean(axis=1).var(ddof=1)
        assert np.isclose(S[p, 9], np.sqrt((W * (h - 1) / h + Bv) / W), rtol=1e-10)
        # rank-normalised split chains through the same estimator
        from scipy.stats import norm, rankdata
        order = np.concatenate([np.stack([col[:h, c], col[n - h:, c]]) for c in range(C)])  # [2C, h] oracle layout
        z = norm.ppf((rankdata(order.reshape(-1), method="ordinal") - 0.375) / (order.size + 0.25)).reshape(order.shape)
        assert np.isclose(S[p, 10], numpy_ess(z), rtol=1e-8)
        assert np.isclose(S[p, 1], S[p, 2] / np.sqrt(S[p, 8]), rtol=1e-12)
        # a strongly antithetic column (tau <= 0 before the clamp) still has a finite, positive n_eff and se_mean
        assert np.isfinite(S[p, 8]) and 0 < S[p, 8] <= C * n * np.log10(C * n) * (1 + 1e-12) and np.isfinite(S[p, 1])
    if abs(phi) < 0.9:
        assert S[1, 9] > S[0, 9]


def test_summary_edge_cases(O):
    x = np.ones((50, 2, 1))
    S = O.summary(x)
    assert S[0, 0] == 1.0 and S[0, 2] == 0.0 and np.isnan(S[0, 8]) and np.isnan(S[0, 9])
    x = np.random.default_rng(0).standard_normal((50, 2, 1)); x[3, 1, 0] = np.nan
    assert np.all(np.isnan(O.summary(x)))


def test_monoexp_map_recovers_truth(O):
    S = synth.make_profiles(10)
    idx = np.where(S["mod_kind"] == 0)[0]
    b = abi.make_problems_dense(S["x"], S["Y"][idx], S["UY"][idx], S["theta0"][idx], S["Sigma0"][idx], Nn=0)
    theta, H, br, st = O.monoexp_map(b, len(idx), abi.default_spec(abi.FOCT_MONOEXP))
    assert np.all(st == 0)
    for j in range(len(idx)):
        cov = np.linalg.inv(-H[j])
        z = (theta[j] - np.array([1000.0, 2000.0, 300.0])) / np.sqrt(np.diag(cov))
        assert np.all(np.abs(z) < 4.5), z
        assert 0.75 < br[j] < 1.25
        # gradient of lp vanishes at the optimum
        _, g, _ = O.logp_grad(abi.FOCT_MONOEXP, b, j, abi.default_spec(abi.FOCT_MONOEXP), theta[j][None, :])
        assert np.all(np.abs(g) < 1e-6)


def test_expgp_parameter_recovery_and_rhat(O):
    # SURVEY §8c ladder (4): truth of synthData.R is recovered, sigma factor ~ 1, split-Rhat < 1.05 at 300 draws
    S = synth.make_profiles(2, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10)
    out = O.sample(0, b, 2, abi.default_spec(), abi.default_cfg(n_warmup=300, n_iter=600, seed=5))
    for j in range(2):
        s = out["summary"][j]
        assert np.all(s[:15, 9] < 1.06)
        assert abs(s[14, 0] - 1.0) < 0.15            # sigma
        assert 0.7 < s[15, 0] < 1.3                   # br
        assert abs(s[2, 0] - 300.0) < 5 * s[2, 2]     # theta3 = 2*l0 (dataType 2)
        m, _, dl = O.predict(0, b, j, abi.default_spec(), out["draws"][j, ::20, 0])
        truth = synth.modulation(int(S["mod_kind"][j]), S["x"])
        # posterior-mean modulation tracks the sinc curve of synthData.R:21 within a loose band
        assert np.abs(dl.mean(axis=0) - truth)[40:].max() < 0.08
    assert out["n_divergent"].sum() <= 2


def test_expgp_map_is_a_stationary_point(O):
    # MODEL_SPEC §10: BFGS optimum of lp - Jacobian; gradient vanishes, Hessian negative definite and symmetric,
    # and the mode sits inside the bulk of the NUTS posterior of the same profile
    S = synth.make_profiles(3, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10)
    spec = abi.default_spec()
    par, H, st = O.expgp_map(b, 3, spec)
    assert np.all(st == 0)
    for j in range(3):
        q = np.concatenate([par[j, :13], np.log(par[j, 13:15])])[None]
        lp, g, c2 = O.logp_grad(0, b, j, spec, q)
        gj = g[0].copy(); gj[13] -= 1; gj[14] -= 1
        scale = np.sqrt(np.abs(np.diag(H[j])))            # natural gradient scale of each coordinate
        assert np.all(np.abs(gj) / scale < 1e-3)
        assert np.isclose(par[j, 16], lp[0] - q[0, 13] - q[0, 14], rtol=1e-14)
        assert np.isclose(par[j, 15], c2[0] / (481 - 13), rtol=1e-12)
        Hs = 0.5 * (H[j] + H[j].T)
        assert np.abs(H[j] - H[j].T).max() < 1e-6 * np.abs(H[j]).max()
        assert np.linalg.eigvalsh(Hs).max() < 0
    out = O.sample(0, b, 1, spec, abi.default_cfg(n_warmup=300, n_iter=600, seed=8), n_threads=4)
    s = out["summary"][0]
    assert np.all(np.abs(par[0, :13] - s[:13, 0]) < 4 * s[:13, 2])


def test_correlated_normal_moments(O):
    # dense precision matrix, correlations up to 0.95: the diagonal metric cannot decorrelate, so this exercises
    # long trees, the three U-turn checks per merge and the multinomial sampling; known covariance = inv(A)
    rng = np.random.default_rng(5)
    D = 6
    sd = np.array([1.0, 5.0, 0.2, 2.0, 1.0, 0.5])
    R = np.full((D, D), 0.3) + 0.7 * np.eye(D)
    R[0, 1] = R[1, 0] = 0.95
    R[2, 3] = R[3, 2] = -0.9
    w, V = np.linalg.eigh(R)
    R = (V * np.maximum(w, 0.02)) @ V.T
    d = np.sqrt(np.diag(R)); R = R / np.outer(d, d)
    cov = R * np.outer(sd, sd)
    A = np.linalg.inv(cov)
    cfg = abi.default_cfg(n_warmup=500, n_iter=3000, seed=21)
    dr, sp = O.sample_analytic(2, A.reshape(-1), cfg)
    S = O.summary(dr)
    flat = dr.reshape(-1, D)
    assert np.all(S[:, 9] < 1.01)
    assert np.all(np.abs(S[:, 0]) < 4 * sd / np.sqrt(S[:, 8]))               # means within 4 MCSE
    assert np.all(np.abs(S[:, 2] / sd - 1) < 0.08)                           # marginal sds
    C = np.corrcoef(flat.T)
    assert abs(C[0, 1] - R[0, 1]) < 0.02 and abs(C[2, 3] - R[2, 3]) < 0.03   # strong correlations recovered
    assert np.abs(C - R).max() < 0.06
    assert sp[..., 4].sum() == 0 and sp[..., 2].max() <= 10
    assert sp[..., 2].mean() > 2.5                                            # correlation => deeper trees than iid


# ---- method = 'vb' (MODEL_SPEC §14): known answers of mean-field ADVI ----
def test_advi_recovers_independent_normal(O):
    """On independent normals the mean-field family contains the target: mu -> 0, exp(omega) -> sd."""
    from fitoct_b200 import _abi as abi
    sd = np.array([0.5, 2.0, 10.0, 0.1])
    cfg = abi.default_vb_cfg(tol_rel_obj=1e-4, iter=20000, seed=3)
    r = O.vb_analytic(0, sd, cfg, np.array([1.0, -3.0, 5.0, 0.3]))
    assert r["status"] in (0, 1) and r["eta"] in (100.0, 10.0, 1.0, 0.1, 0.01)
    assert np.all(np.abs(r["mu"]) < np.maximum(0.15 * sd, 0.05))   # the normalised steps jitter by ~eta/sqrt(k) whatever the scale
    assert np.allclose(np.exp(r["omega"]), sd, rtol=0.15)


def test_advi_on_correlated_normal_gives_the_mean_field_optimum(O):
    """For N(0, Lambda^-1) the KL(q||p)-optimal diagonal normal has variances 1/Lambda_dd (NOT the marginal variances)."""
    from fitoct_b200 import _abi as abi
    A = np.array([[2.0, 0.9, 0.0], [0.9, 1.0, 0.3], [0.0, 0.3, 4.0]])
    cfg = abi.default_vb_cfg(tol_rel_obj=1e-4, iter=20000, seed=4)
    r = O.vb_analytic(2, A.ravel(), cfg, np.array([1.0, -1.0, 0.5]))
    assert np.all(np.abs(r["mu"]) < 0.1)
    assert np.allclose(np.exp(r["omega"]), 1 / np.sqrt(np.diag(A)), rtol=0.15)
    marg = np.sqrt(np.diag(np.linalg.inv(A)))
    assert np.all(np.exp(r["omega"])[:2] < marg[:2])   # mean-field under-disperses the correlated pair


def test_advi_default_stopping_rule_and_failure_mode(O):
    from fitoct_b200 import _abi as abi, synth
    cfg = abi.default_vb_cfg(seed=5)
    r = O.vb_analytic(0, np.array([1.0, 1.0]), cfg, np.array([0.5, -0.5]))
    # an ELBO of order 1 estimated from 100 draws never settles to 1 %: the run ends at the iteration limit, as in Stan
    assert r["status"] == 1 and r["iters"] == 10000
    # Stan's start (omega = 0, unit sd on every unconstrained component) puts 1 + dL <= 0 in the FitOCT model: ADVI stops
    # with "dropped evaluations" (status 2); a narrower start converges
    S = synth.make_profiles(1, modulated_only=True)
    batch = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=10)
    spec = abi.default_spec(abi.FOCT_EXPGP)
    assert O.vb(abi.FOCT_EXPGP, batch, 1, spec, abi.default_vb_cfg(), draws=False)["status"][0] == 2
    out = O.vb(abi.FOCT_EXPGP, batch, 1, spec, abi.default_vb_cfg(omega0=-3.0, output_samples=200))
    # eval every 100 steps, first relative change is 1, upper median of 3 entries: never before 300 iterations
    assert out["status"][0] == 0 and out["iters"][0] % 100 == 0 and out["iters"][0] >= 300
    m = out["mean"][0]
    assert abs(m[0] - 1000) < 15 and abs(m[1] - 2000) < 60 and abs(m[2] - 300) < 15 and 0.8 < m[14] < 1.3
    assert np.all(out["draws"][0][:, -1] == 0.0) and np.all(np.isfinite(out["draws"][0][:, -2]))
