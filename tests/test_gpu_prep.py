"""GPU parity of the steps either side of the sampling path (MODEL_SPEC §11-13) against the CPU
oracle, through the C ABI: estimateNoise, the printBr gate, estimateExpPrior and the batch pipeline."""
import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth

pytestmark = pytest.mark.gpu

RTOL_SPLINE = 1e-9   # fp64 banded solves, summation order differs between warp and serial code
RTOL_NOISE = 1e-8


def _xy_batch(xs, ys, dataType=2):
    return abi.make_problems([dict(x=x, y=y, uy=np.ones_like(x), dataType=dataType, Nn=0, gridType=0, rho=1.0,
                                   lambda_rate=0.0, theta0=(0, 0, 1), Sigma0=np.eye(3), prior_PD=0, id=j)
                              for j, (x, y) in enumerate(zip(xs, ys))])


def _ragged_profiles():
    rng = np.random.default_rng(7)
    xs, ys = [], []
    for N in (481, 481, 30, 60, 199, 200, 1000, 3300):
        x = 20.0 + np.arange(N) * (480.0 / (N - 1)) + (rng.uniform(-0.2, 0.2, N) * (480.0 / (N - 1)) if N != 481 else 0.0)
        y0 = 1000 + 2000 * np.exp(-x / 150)
        sd = 0.5 * np.sqrt(y0 - 1000 + 1)
        xs.append(x)
        ys.append(y0 * (1 + 0.002 * np.sin(x / 25)) + sd * rng.standard_normal(N))
    return xs, ys


def test_estimate_noise_matches_oracle_on_ragged_batch(L, O):
    xs, ys = _ragged_profiles()
    batch = _xy_batch(xs, ys)
    n = len(xs)
    g = L.estimate_noise(batch, n, df=15.0)
    uy_o, ys_o, th_o, info_o = O.estimate_noise(batch, n, df=15.0)
    assert g["status"].tolist() == [0] * n
    for j in range(n):
        scale = np.max(np.abs(ys[j]))
        assert np.max(np.abs(g["ySmooth"][j] - ys_o[j])) <= RTOL_SPLINE * scale, j
        assert np.allclose(g["theta"][j], th_o[j], rtol=RTOL_NOISE), (j, g["theta"][j], th_o[j])
        assert np.allclose(g["uy"][j], uy_o[j], rtol=RTOL_NOISE)
        assert abs(g["info"][j, 2] - 15.0) <= 1e-9 and abs(g["info"][j, 0] - info_o[j, 0]) <= 1e-8
    # the fitted noise law is close to the generating one (a1 = 0.5*sqrt(2000), a2 = 300) on the full-length profiles
    assert abs(g["theta"][0, 0] - 22.4) < 4 and abs(g["theta"][0, 1] - 300) < 80


@pytest.mark.parametrize("df", [4.0, 15.0, 30.0])
def test_estimate_noise_df_sweep(L, O, df):
    """Tests/statsSplineSmooth.R sweeps df 2..20 and looks at sd(residuals): the same statistic from both sides."""
    S = synth.make_profiles(6, modulated_only=True)
    batch = abi.make_problems_dense(S["x"], S["Y"], np.ones_like(S["Y"]), np.tile([0.0, 0.0, 1.0], (6, 1)),
                                    np.tile(np.eye(3), (6, 1, 1)), Nn=0)
    g = L.estimate_noise(batch, 6, df=df)
    _, ys_o, th_o, _ = O.estimate_noise(batch, 6, df=df)
    for j in range(6):
        sd_g = np.std(S["Y"][j] - g["ySmooth"][j], ddof=1)
        sd_o = np.std(S["Y"][j] - ys_o[j], ddof=1)
        assert sd_g == pytest.approx(sd_o, rel=1e-9)
        assert np.allclose(g["theta"][j], th_o[j], rtol=RTOL_NOISE)


def test_estimate_noise_rejects_bad_input(L):
    x = np.array([1.0, 2.0, 2.0, 3.0, 4.0, 5.0])
    g = L.estimate_noise(_xy_batch([x], [np.ones(6)]), 1, df=3.0)
    assert g["status"][0] == 3 and np.all(np.isnan(g["uy"][0]))
    from fitoct_b200._lib import FitOCTError
    with pytest.raises(FitOCTError):
        L.estimate_noise(_xy_batch([np.arange(6.0)], [np.ones(6)]), 1, df=40.0)   # more df than coefficients
    with pytest.raises(FitOCTError):
        L.estimate_noise(_xy_batch([np.arange(3.0)], [np.ones(3)]), 1, df=2.0)    # N < 4


def test_print_br_and_exp_prior_match_oracle(L, O):
    n = 16
    S = synth.make_profiles(n)
    spec = abi.default_spec(abi.FOCT_MONOEXP)
    batch = abi.make_problems_dense(S["x"], S["Y"], S["UY"], np.tile([0.0, 0.0, 1.0], (n, 1)), np.tile(np.eye(3), (n, 1, 1)), Nn=0)
    th, H, br, st = L.monoexp_map(batch, n, spec)
    ci_g, al_g = L.print_br(abi.FOCT_MONOEXP, batch, n, spec, br)
    ci_o, al_o = O.print_br(br, S["x"].size - 3)
    assert np.allclose(ci_g, ci_o[None, :], rtol=1e-12) and al_g.tolist() == al_o.tolist()
    assert 0 < al_g.sum() < n   # the synthetic family mixes unmodulated (OK) and modulated (alert) profiles
    for pt in ("mono", "abc"):
        t0g, S0g, rug = L.estimate_exp_prior(batch, n, pt, th, H, ru_theta=0.07)
        t0o, S0o, ruo = O.exp_prior(batch, n, pt, th, H, ru_theta=0.07)
        assert np.array_equal(t0g, th) and np.allclose(t0g, t0o, rtol=0, atol=0)
        assert np.allclose(rug, ruo, rtol=1e-11), pt
        assert np.allclose(S0g, S0o, rtol=1e-10, atol=0), pt


def test_pipeline_equals_the_step_by_step_calls(L):
    n = 12
    S = synth.make_profiles(n)
    xy = abi.make_problems_dense(S["x"], S["Y"], np.ones_like(S["Y"]), np.tile([0.0, 0.0, 1.0], (n, 1)),
                                 np.tile(np.eye(3), (n, 1, 1)), Nn=0)
    pc = L.pipeline_cfg(Nn=8, prior_type=abi.FOCT_PRIOR_ABC, smooth_df=15.0)
    cfg = abi.default_cfg(chains=2, n_warmup=60, n_iter=120, seed=99)
    out = L.pipeline(xy, n, pc, cfg, draws=True, summary=True)
    # the same thing by hand, as FitOCT.R:89-124 spells it
    nz = L.estimate_noise(xy, n, df=15.0)
    UY = np.stack(nz["uy"])
    assert np.array_equal(np.stack(out["uy"]), UY) and np.array_equal(out["noise_theta"], nz["theta"])
    spec_m = abi.default_spec(abi.FOCT_MONOEXP)
    mono = abi.make_problems_dense(S["x"], S["Y"], UY, np.tile([0.0, 0.0, 1.0], (n, 1)), np.tile(np.eye(3), (n, 1, 1)), Nn=0)
    th, H, br, st = L.monoexp_map(mono, n, spec_m)
    assert np.array_equal(out["mono_theta"], th) and np.array_equal(out["mono_br"], br)
    ci, alert = L.print_br(abi.FOCT_MONOEXP, mono, n, spec_m, br)
    assert out["alert"].tolist() == alert.tolist()
    t0, S0, ru = L.estimate_exp_prior(mono, n, "abc", th, H)
    assert np.array_equal(out["theta0"], t0) and np.array_equal(out["Sigma0"], S0)
    idx = np.flatnonzero(alert)
    assert out["n_expgp"] == idx.size and out["expgp_index"].tolist() == idx.tolist() and 0 < idx.size < n
    gp = abi.make_problems_dense(S["x"], S["Y"][idx], UY[idx], t0[idx], S0[idx], Nn=8, gridType=0, rho=1.0 / 8,
                                 lambda_rate=0.1, ids=idx)
    ref = L.sample(abi.FOCT_EXPGP, gp, idx.size, abi.default_spec(abi.FOCT_EXPGP), cfg, draws=True, summary=True)
    assert np.array_equal(out["expgp"]["draws"], ref["draws"])
    assert np.array_equal(out["expgp"]["summary"], ref["summary"], equal_nan=True)
    # gate = 0 sends every profile on
    pc.gate = 0
    out_all = L.pipeline(xy, n, pc, cfg, draws=False, summary=True)
    assert out_all["n_expgp"] == n


def test_reference_script_flow_through_the_api():
    """FitOCT.R:84-124 written with the mirrored operator names."""
    from fitoct_b200 import api
    S = synth.make_profiles(2, modulated_only=True)
    from fitoct_b200.io import selX
    x, y = selX(S["x"], S["Y"][0], depthSel=None, subSample=1)
    fits = api.estimateNoise(x, y, df=15)
    uy = fits["uy"]
    assert uy.shape == x.shape and np.all(uy > 0) and fits["fit"]["df"] == pytest.approx(15.0, abs=1e-8)
    fitm = api.fitMonoExp(x, y, uy, dataType=2)
    br = api.printBr(fitm["fit"], silent=True)
    assert br["alert"] is not None            # a modulated profile fails the mono-exponential fit
    for pt in ("mono", "abc"):
        pri = api.estimateExpPrior(x, uy, 2, pt, out=fitm, ru_theta=0.05, eps=1e-3)
        assert np.allclose(pri["theta0"], fitm["best_theta"]) and np.all(np.linalg.eigvalsh(pri["Sigma0"]) > 0)
    fitGP = api.fitExpGP(x, y, uy, dataType=2, Nn=10, gridType="internal", method="sample", theta0=pri["theta0"],
                         Sigma0=pri["Sigma0"], lambda_rate=0.1, rho_scale=0, nb_warmup=150, nb_iter=300, prior_PD=0)
    brGP = api.printBr(fitGP["fit"], N=x.size, silent=True)
    assert brGP["br"] < br["br"] and brGP["alert"] is None   # the GP-modulated model repairs the fit


def test_fitoct_batch_driver_and_scale(L):
    from fitoct_b200 import api
    import time
    n = 1000
    S = synth.make_profiles(n)
    xy = abi.make_problems_dense(S["x"], S["Y"], np.ones_like(S["Y"]), np.tile([0.0, 0.0, 1.0], (n, 1)),
                                 np.tile(np.eye(3), (n, 1, 1)), Nn=0)
    t0 = time.perf_counter()
    g = L.estimate_noise(xy, n, df=15.0)
    dt = time.perf_counter() - t0
    assert np.all(g["status"] == 0) and np.max(np.abs(g["info"][:, 2] - 15.0)) < 1e-9
    a = g["theta"]
    assert abs(np.median(a[:, 0]) - 22.4) < 2.5 and abs(np.median(a[:, 1]) - 300) < 40
    print(f"estimate_noise: {n} profiles in {dt * 1e3:.1f} ms end to end")
    out = api.FitOCT_batch(S["x"], S["Y"][:40], dict(nb_warmup=100, nb_sample=100, Nn=10, priorType="abc"), chains=4)
    kinds = S["mod_kind"][:40]
    # the gate passes (nearly) every unmodulated profile and stops the strongly modulated ones
    assert np.mean(out["alert"][kinds == 0]) <= 0.25 and np.mean(out["alert"][kinds == 1]) >= 0.75
    k = out["n_expgp"]
    assert k == int(out["alert"].sum()) and out["expgp"]["summary"].shape[0] == k
    assert np.all(np.isfinite(out["expgp"]["summary"][:, :, 0]))


def test_fitoct_batch_optim_and_vb_methods(L):
    """ctrlParams' `method` (FitOCT.R:42): the Shiny default 'optim' and 'vb' through the same batch driver."""
    from fitoct_b200 import api
    S = synth.make_profiles(10)
    kinds = S["mod_kind"]
    o = api.FitOCT_batch(S["x"], S["Y"], dict(method="optim", Nn=8))
    k = o["n_expgp"]
    assert o["method"] == "optim" and k == int(o["alert"].sum()) == int((kinds > 0).sum())
    assert o["expgp"]["par"].shape == (k, 8 + 7) and o["expgp"]["hessian"].shape == (k, 13, 13)
    assert np.all(o["expgp"]["status"] == 0)
    br_gp = o["expgp"]["par"][:, 8 + 5]
    assert np.all(br_gp < o["mono_br"][o["expgp_index"]])          # the modulated model repairs every alerted fit
    v = api.FitOCT_batch(S["x"], S["Y"], dict(method="vb", Nn=8))
    assert v["n_expgp"] == k and np.all(v["expgp"]["status"] == 0)
    # ADVI means (stopped at Stan's loose tol_rel_obj) and the MAP agree on the decay parameters to several per cent
    assert np.allclose(v["expgp"]["mean"][:, :3], o["expgp"]["par"][:, :3], rtol=0.1)
    with pytest.raises(ValueError):
        api.FitOCT_batch(S["x"], S["Y"], dict(method="laplace"))


def test_batch_driver_on_a_data_directory(tmp_path, capsys):
    """`python -m fitoct_b200 <dataDir>`: FitOCT.R:72-131 over a directory of Courbe.csv files."""
    from fitoct_b200 import __main__ as drv
    from fitoct_b200 import io as fio
    S = synth.make_profiles(5)                       # modulation kinds 0..4
    data = tmp_path / "DataSynth"
    for j in range(5):
        d = data / f"set{j}"
        d.mkdir(parents=True)
        fio.write_courbe_csv(str(d / "Courbe.csv"), S["x"], S["Y"][j])
    ctrl = tmp_path / "ctrlParams.yaml"
    ctrl.write_text("nb_warmup: 100\nnb_sample: 100\nNn: 6\nrho_scale: 0\n")
    out = tmp_path / "Results"
    assert drv.main([str(data), "--ctrl", str(ctrl), "--out", str(out), "--chains", "2"]) == 0
    printed = capsys.readouterr().out
    assert "DataSynth_set0" in printed and "MonoExp fit OK" in printed and "ExpGP (sample)" in printed
    txt0 = (out / "DataSynth_set0_ctrl.txt").read_text()
    assert "MonoExp decay parameters" in txt0 and "ExpGP parameters" not in txt0 and "WARNING" not in txt0
    txt1 = (out / "DataSynth_set1_ctrl.txt").read_text()
    assert "WARNING" in txt1 and "ExpGP parameters" in txt1 and "yGP[6]" in txt1
    assert not (out / "DataSynth_set0_ExpGP_1.csv").exists()
    back = fio.read_stan_csv([str(out / f"DataSynth_set1_ExpGP_{c}.csv") for c in (1, 2)])
    assert back["draws"].shape == (100, 2, 6 + 7) and np.all(np.isfinite(back["draws"]))
    # the Shiny default method through the same driver
    ctrl.write_text("method: optim\nNn: 6\n")
    assert drv.main([str(data), "--ctrl", str(ctrl), "--out", str(tmp_path / "R2")]) == 0
    assert "theta  :" in (tmp_path / "R2" / "DataSynth_set2_ctrl.txt").read_text()


def test_prep_edge_cases(L, O):
    rng = np.random.default_rng(3)
    # the smallest spline R accepts (4 points: every x a knot) and a df just above the straight line
    x = np.array([0.0, 1.0, 2.5, 4.0])
    y = np.array([1.0, 2.2, 2.9, 4.5])
    b = _xy_batch([x], [y])
    g = L.estimate_noise(b, 1, df=2.5)
    _, ys_o, th_o, info_o = O.estimate_noise(b, 1, df=2.5)
    assert g["status"][0] == 0 and np.allclose(g["ySmooth"][0], ys_o[0], rtol=1e-9) and abs(g["info"][0, 2] - 2.5) < 1e-9
    # a pipeline whose gate lets nothing through returns without sampling
    S = synth.make_profiles(10)
    keep = np.flatnonzero(S["mod_kind"] == 0)
    xy = abi.make_problems_dense(S["x"], S["Y"][keep], np.ones_like(S["Y"][keep]), np.tile([0.0, 0.0, 1.0], (keep.size, 1)),
                                 np.tile(np.eye(3), (keep.size, 1, 1)), Nn=0)
    cfg = abi.default_cfg(chains=2, n_warmup=30, n_iter=60)
    out = L.pipeline(xy, keep.size, L.pipeline_cfg(), cfg)
    assert out["n_expgp"] == 0 and out["alert"].sum() == 0 and out["expgp"]["summary"].shape[0] == 0
    # a single profile, gate off
    out1 = L.pipeline(xy, 1, L.pipeline_cfg(gate=0, Nn=5), cfg)
    assert out1["n_expgp"] == 1 and np.all(np.isfinite(out1["expgp"]["summary"][0, :10, 0]))
    # print_br needs degrees of freedom
    from fitoct_b200._lib import FitOCTError
    tiny = abi.make_problems([dict(x=x, y=y, uy=np.ones(4), dataType=2, Nn=10)])
    ci, al = L.print_br(abi.FOCT_MONOEXP, tiny, 1, abi.default_spec(abi.FOCT_MONOEXP), np.array([1.0]))   # N - 3 = 1 is fine ...
    assert ci.shape == (1, 2) and al.tolist() == [0]
    with pytest.raises(FitOCTError):
        L.print_br(abi.FOCT_EXPGP, tiny, 1, abi.default_spec(abi.FOCT_EXPGP), np.array([1.0]))           # ... N - 3 - Nn < 1 is not


def test_pipeline_survives_degenerate_profiles(L):
    """One bad Courbe.csv must not abort the directory (FitOCT.R:74-84 loops over files independently): a constant signal
    (no decay to fit: the MonoExp Hessian is singular) and a profile of pure noise ride in a batch of good ones.  The call
    succeeds, every profile gets a status, the good ones are fitted exactly as they are without the bad ones."""
    S = synth.make_profiles(6, modulated_only=True)
    Y = S["Y"].copy()
    Y[1] = 1234.5                                                    # constant
    Y[4] = 1000.0 + np.random.default_rng(8).standard_normal(Y.shape[1])   # noise around a constant
    dummy = np.ones_like(Y)
    th = np.tile([0.0, 0.0, 1.0], (6, 1)); S0 = np.tile(np.eye(3), (6, 1, 1))
    b = abi.make_problems_dense(S["x"], Y, dummy, th, S0, Nn=0, ids=S["ids"])
    cfg = abi.default_cfg(chains=2, n_warmup=60, n_iter=120, seed=5)
    out = L.pipeline(b, 6, L.pipeline_cfg(gate=1, Nn=6), cfg)
    st = out["status"]
    assert set(st.tolist()) <= {0, 1, 2, 3} and np.all(st[[0, 2, 3, 5]] == 0)      # the modulated profiles are sampled
    assert st[1] in (1, 2, 3) and st[4] in (1, 2, 3)
    k = out["n_expgp"]
    assert k == int(np.sum((st == 0) | (st == 2))) and sorted(out["expgp_index"].tolist()) == sorted(np.flatnonzero((st == 0) | (st == 2)).tolist())
    good = [i for i, j in enumerate(out["expgp_index"]) if st[j] == 0]
    assert np.all(np.isfinite(out["expgp"]["summary"][good][:, :11, 0]))
    # the good profiles alone: bit-identical fits (a profile's draws depend on its id, not on its neighbours)
    keep = [0, 2, 3, 5]
    b2 = abi.make_problems_dense(S["x"], Y[keep], dummy[keep], th[keep], S0[keep], Nn=0, ids=S["ids"][keep])
    out2 = L.pipeline(b2, 4, L.pipeline_cfg(gate=1, Nn=6), cfg)
    assert np.all(out2["status"] == 0)
    sel = [list(out["expgp_index"]).index(j) for j in keep]
    np.testing.assert_array_equal(out["expgp"]["summary"][sel], out2["expgp"]["summary"])


def test_pripost_flow_prior_predictive_known_answers(L):
    """priPost.R:2-16: the same call with prior_PD = 1 samples the prior.  (a) the three-parameter model with the MVN prior
    switched on is a pure Gaussian target with known moments; (b) for fitExpGP the prior of (yGP, lambda) is a funnel that
    NUTS mixes slowly in (as it does in Stan), so only robust statements are made: br is dropped, the prior is centred on
    theta0 and wider than the posterior."""
    from fitoct_b200 import api
    S = synth.make_profiles(2, modulated_only=True)
    x, y, uy = S["x"], S["Y"][0], S["UY"][0]
    theta0 = S["theta0"][0]
    sd0 = 0.05 * theta0
    cor = np.array([[1.0, -0.5, 0.3], [-0.5, 1.0, -0.4], [0.3, -0.4, 1.0]])
    Sigma0 = np.outer(sd0, sd0) * cor
    # (a)
    spec = abi.default_spec(abi.FOCT_MONOEXP)
    spec.theta_prior = 0
    b = abi.make_problems_dense(x, y[None, :], uy[None, :], theta0[None, :], Sigma0[None], Nn=0, prior_PD=1)
    out = L.sample(abi.FOCT_MONOEXP, b, 1, spec, abi.default_cfg(chains=4, n_warmup=500, n_iter=2500, seed=11))
    t = out["summary"][0]
    assert np.all(np.abs(t[:3, 0] - theta0) < 4 * t[:3, 1])
    assert np.allclose(t[:3, 2], sd0, rtol=0.06)
    c = np.corrcoef(out["draws"][0].reshape(-1, 5)[:, :3].T)
    assert np.max(np.abs(c - cor)) < 0.06
    assert np.all(np.isnan(t[3]))                                                          # br is undefined without data
    # (b)
    kw = dict(dataType=2, Nn=10, gridType="internal", method="sample", theta0=theta0, Sigma0=Sigma0, lambda_rate=0.1,
              rho_scale=0, nb_warmup=500, nb_iter=1500)
    pri = api.fitExpGP(x, y, uy, prior_PD=1, **kw)
    pos = api.fitExpGP(x, y, uy, prior_PD=0, **kw)
    assert pri["prior_PD"] == 1 and np.all(np.isnan(pri["fit"].extract("br")["br"]))      # plotExpGP.R:42-43 drops br
    tp, tq = pri["fit"].summary_table, pos["fit"].summary_table
    assert np.all(np.abs(tp[:3, 0] - theta0) < 5 * tp[:3, 1] + 0.005 * sd0)
    assert np.allclose(tp[:3, 2], sd0, rtol=0.3) and abs(tp[14, 0] - 1.0) < 0.05
    assert np.all(tq[:3, 2] < tp[:3, 2]) and tq[14, 2] < tp[14, 2]                         # the data inform theta and sigma
