"""CPU: the on-disk formats either side of the path: Courbe.csv trees, Stan-CSV, _ctrl.txt."""
import os

import numpy as np

from fitoct_b200 import _abi as abi
from fitoct_b200 import api, io, synth


def test_courbe_csv_roundtrip_and_dir_scan(tmp_path):
    S = synth.make_profiles(3)
    d = tmp_path / "DataSynth"
    for j, name in enumerate(["monoExp", "sincExp", "sincExp1"]):        # synthData.R:16,29,43
        (d / name).mkdir(parents=True)
        io.write_courbe_csv(str(d / name / "Courbe.csv"), S["x"], S["Y"][j])
    (d / "notes").mkdir()                                                 # a directory without Courbe.csv is skipped
    sets = io.read_data_dir(str(d))
    assert [t for t, _, _ in sets] == ["DataSynth_monoExp", "DataSynth_sincExp", "DataSynth_sincExp1"]   # FitOCT.R:80 tag
    for j, (_, x, y) in enumerate(sets):
        np.testing.assert_array_equal(x, S["x"])
        np.testing.assert_allclose(y, S["Y"][j], rtol=0, atol=0)
    first = open(d / "monoExp" / "Courbe.csv").readline().strip()
    assert first == '"x","y"'                                             # write.csv header


def test_selX():
    x = np.arange(20.0, 501.0); y = x * 2
    a, b = io.selX(x, y, None, 1)
    assert a.size == 481
    a, b = io.selX(x, y, (100, 200), 2)                                   # depthSel window then sub-sampling (FitOCT.R:85)
    assert a[0] == 100 and a[-1] == 200 and np.all(np.diff(a) == 2) and np.all(b == 2 * a)


def _fake_fit(save_warmup):
    names = abi.param_names(abi.FOCT_EXPGP, 3)
    rng = np.random.default_rng(1)
    n_w, n_it, C = 6, 14, 2
    n_saved = n_it if save_warmup else n_it - n_w
    draws = rng.standard_normal((n_saved, C, len(names)))
    draws[0, 0, names.index("br")] = np.nan
    sp = rng.random((n_saved, C, 6))
    return api.StanFit(names, draws, sp, n_warmup=n_w, n_iter=n_it, save_warmup=save_warmup,
                       summary_table=rng.random((len(names), 11)), stepsize=np.array([0.02, 0.03]),
                       inv_metric=rng.random((C, 8)), n_divergent=np.zeros(C))


def test_stan_csv_roundtrip(tmp_path):
    for save_warmup in (True, False):
        fit = _fake_fit(save_warmup)
        paths = io.write_stan_csv(fit, str(tmp_path / f"fit{int(save_warmup)}"))
        assert len(paths) == 2 and all(os.path.exists(p) for p in paths)
        head = [l for l in open(paths[0]) if not l.startswith("#")][0].strip().split(",")
        assert head[:7] == ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
        assert head[7:] == ["theta.1", "theta.2", "theta.3", "yGP.1", "yGP.2", "yGP.3", "lambda", "sigma", "br"]
        txt = open(paths[1]).read()
        assert "# Adaptation terminated" in txt and "# Step size = 0.03" in txt
        assert f"num_warmup = {fit.n_warmup}" in txt and f"save_warmup = {int(save_warmup)}" in txt
        back = io.read_stan_csv(paths)
        np.testing.assert_array_equal(back["draws"], fit.draws)           # repr() round-trips doubles exactly, NaN kept
        np.testing.assert_array_equal(back["sampler_params"], fit.sampler_params)
        np.testing.assert_array_equal(back["stepsize"], fit.stepsize)
        np.testing.assert_array_equal(back["inv_metric"], fit.inv_metric)
        assert back["par_names"] == ["theta.1", "theta.2", "theta.3", "yGP.1", "yGP.2", "yGP.3", "lambda", "sigma", "br", "lp__"]


def test_ctrl_txt(tmp_path):
    fit = _fake_fit(True)
    p = str(tmp_path / "DataSynth_sincExp_ctrl.txt")
    io.write_ctrl_txt(p, fit, append=False)
    io.write_ctrl_txt(p, fit, pars=("theta",), title="again")
    txt = open(p).read()
    assert " ExpGP parameters:" in txt and " again:" in txt                # plotExpGP.R:6
    assert "theta[1]" in txt and "yGP[3]" in txt and "lp__" not in txt.split("again")[0].split("Rhat")[1]
    assert "n_eff" in txt and "Rhat" in txt and "2.5%" in txt
