"""The oracle for the steps either side of the path (MODEL_SPEC §11-13), pinned against scipy where a public
implementation of the same mathematics exists.  CPU only."""
import numpy as np
import pytest
from scipy.interpolate import BSpline, make_smoothing_spline
from scipy.stats import chi2

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth


def _profile(seed=1, n=481):
    rng = np.random.default_rng(seed)
    x = np.arange(20.0, 20.0 + n)
    y0 = 1000 + 2000 * np.exp(-x / 150)
    sd = 0.5 * np.sqrt(y0 - 1000 + 1)
    return x, y0 + sd * rng.standard_normal(n), sd


def test_nknots_matches_r_break_points(O):
    # documented values of R's .nknots.smspl
    assert [O.nknots(n) for n in (10, 49, 50, 200, 800, 3200)] == [10, 49, 50, 100, 140, 200]
    assert O.nknots(481) == 117
    assert O.nknots(3200 + 32) == 202


@pytest.mark.parametrize("spar", [0.2, 0.6, 1.0])
def test_all_knots_spline_is_the_natural_smoothing_spline(O, spar):
    """With every x a knot the penalised B-spline fit is the natural cubic smoothing spline: scipy's
    make_smoothing_spline at the same penalty (lambda rescaled from t = (x-x1)/(xN-x1) to x units)."""
    x, y, _ = _profile(3, 200)
    ys, info = O.smooth_spline(x, y, spar=spar, all_knots=True)
    ref = make_smoothing_spline(x, y, lam=info["lam"] * (x[-1] - x[0]) ** 3)(x)
    assert np.max(np.abs(ys - ref)) <= 1e-9 * np.max(np.abs(y))


def test_df_is_the_trace_of_the_hat_matrix(O):
    """df(lambda) from the band-of-inverse recurrence == trace of the dense hat matrix built with scipy's B-splines."""
    x, y, _ = _profile(5)
    ys, info = O.smooth_spline(x, y, df=15.0)
    assert abs(info["df"] - 15.0) <= 1e-9
    assert -1.5 <= info["spar"] <= 1.5
    t = (x - x[0]) / (x[-1] - x[0])
    nkn = O.nknots(x.size)
    idx = np.floor(np.linspace(1, x.size, nkn)).astype(int) - 1
    idx[-1] = x.size - 1
    T = np.concatenate([[t[0]] * 3, t[idx], [t[-1]] * 3])
    X = BSpline.design_matrix(t, T, 3).toarray()
    nk = X.shape[1]
    # penalty by fine Gauss quadrature of B''B''
    Om = np.zeros((nk, nk))
    gx, gw = np.polynomial.legendre.leggauss(3)
    d2 = BSpline(T, np.eye(nk), 3).derivative(2)
    for a, b in zip(T[3:-4], T[4:-3]):
        if b > a:
            pts = 0.5 * (b - a) * gx + 0.5 * (a + b)
            D2 = d2(pts).T  # [nk, 3]
            Om += 0.5 * (b - a) * (D2 * gw) @ D2.T
    A = X.T @ X + info["lam"] * Om
    H = X @ np.linalg.solve(A, X.T)
    assert abs(np.trace(H) - info["df"]) <= 1e-7
    assert np.max(np.abs(H @ y - ys)) <= 1e-8 * np.max(np.abs(y))


def test_spline_reproduces_a_line_and_limits(O):
    x = np.linspace(0.0, 10.0, 120)
    y = 1 + 2 * x
    ys, info = O.smooth_spline(x, y, spar=0.8)        # a straight line is in the penalty's null space: reproduced at any lambda
    assert np.max(np.abs(ys - y)) < 1e-10 * np.max(np.abs(y))
    yl, il = O.smooth_spline(x, np.sin(x), spar=1.5)  # heavy smoothing -> close to the regression line (df -> 2)
    assert il["df"] < 2.6
    with pytest.raises(RuntimeError):
        O.smooth_spline(np.array([0.0, 1.0, 1.0, 2.0, 3.0]), np.zeros(5), df=3.0)  # x must increase strictly


def test_noise_fit_recovers_the_generating_law(O):
    rng = np.random.default_rng(11)
    x = np.arange(20.0, 501.0)
    a1, a2 = 20.0, 250.0
    est = np.array([O.noise_fit(x, a1 * np.exp(-x / a2) * rng.standard_normal(x.size)) for _ in range(40)])
    m, se = est.mean(0), est.std(0) / np.sqrt(40)
    assert abs(m[0] - a1) < 4 * se[0] + 0.02 * a1
    assert abs(m[1] - a2) < 4 * se[1] + 0.03 * a2
    # score is zero at the optimum: d/dv of the profile likelihood
    r = a1 * np.exp(-x / a2) * rng.standard_normal(x.size)
    th = O.noise_fit(x, r)
    v = 1 / th[1]
    w = r * r * np.exp(2 * v * x)
    assert abs(x.mean() - (w * x).sum() / w.sum()) < 1e-8 * x.mean()
    assert np.isclose(th[0] ** 2, np.mean(w), rtol=1e-12)
    # noise growing with depth is reported flat: a2 clamps at maxRate
    th = O.noise_fit(x, np.exp(x / 200) * rng.standard_normal(x.size))
    assert th[1] == pytest.approx(1e4)


def test_qchisq_and_gate_match_scipy(O):
    for ndf in (3, 17, 100, 478, 5000):
        for p in (0.025, 0.5, 0.975):
            assert O.qchisq(p, ndf) == pytest.approx(chi2.ppf(p, ndf), rel=1e-12)
    br = np.array([0.8, 0.95, 1.0, 1.1, 1.2, np.nan])
    ci, alert = O.print_br(br, 478)
    assert ci == pytest.approx(chi2.ppf([0.025, 0.975], 478) / 478, rel=1e-12)
    assert alert.tolist() == [1, 0, 0, 0, 1, 1]


def test_exp_prior_mono_and_abc(O):
    spec = abi.default_spec(abi.FOCT_MONOEXP)
    S = synth.make_profiles(3)
    prof = [dict(x=S["x"], y=S["Y"][j], uy=S["UY"][j]) for j in range(3)]
    batch = abi.make_problems_dense(S["x"], S["Y"], S["UY"], np.tile([0.0, 0.0, 1.0], (3, 1)), np.tile(np.eye(3), (3, 1, 1)),
                                    dataType=2, Nn=0)
    th, H, br, st = O.monoexp_map(batch, 3, spec)
    t0, S0, ru = O.exp_prior(batch, 3, "mono", th, H, ru_theta=0.05)
    cov = np.linalg.inv(-H)
    for j in range(3):
        sd = np.sqrt(np.diag(cov[j]))
        cor = cov[j] / np.outer(sd, sd)
        u = 0.05 * th[j]
        assert np.allclose(S0[j], np.outer(u, u) * cor, rtol=1e-10)
        assert np.allclose(t0[j], th[j]) and ru[j] == 0.05
    t0, S0, ru = O.exp_prior(batch, 3, "abc", th, H)
    for j in range(3):
        x, y = prof[j]["x"], prof[j]["y"]
        sd = np.sqrt(np.diag(cov[j]))
        cor = cov[j] / np.outer(sd, sd)
        e = np.exp(-2 * x / th[j, 2])
        J = np.stack([np.ones_like(x), e, th[j, 1] * e * 2 * x / th[j, 2] ** 2], axis=1)
        Cm = np.outer(th[j], th[j]) * cor
        s = np.sqrt(np.einsum("ia,ab,ib->i", J, Cm, J))
        q95 = np.quantile(np.abs(y - th[j, 0] - th[j, 1] * e), 0.95)
        assert ru[j] == pytest.approx(q95 / (1.96 * s.mean()), rel=1e-10)
        assert np.allclose(np.sqrt(np.diag(S0[j])), ru[j] * np.abs(th[j]), rtol=1e-10)
        assert np.all(np.linalg.eigvalsh(S0[j]) > 0)
