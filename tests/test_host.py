"""CPU: host-side logic — control parameters, synthetic workload generator, batch marshalling, sharding."""
import ctypes as C
import os

import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import api, shard, synth


def test_ctrl_params_defaults_and_yaml_override(tmp_path):
    d = api.load_ctrl_params(None)
    # FitOCT.R:37-53
    assert (d["dataType"], d["nb_warmup"], d["nb_sample"], d["Nn"], d["gridType"], d["rho_scale"], d["lambda_rate"],
            d["ru_theta"], d["method"]) == (2, 500, 1000, 10, "internal", 0.1, 0.1, 0.05, "sample")
    # the shipped ctrlParams.yaml:1-5
    y = tmp_path / "ctrlParams.yaml"
    y.write_text("nb_warmup: 100\nnb_sample: 100\ngridType: extremal\nNn: 15\nrho_scale: 0\n")
    d = api.load_ctrl_params(str(y))
    assert (d["nb_warmup"], d["nb_sample"], d["gridType"], d["Nn"], d["rho_scale"]) == (100, 100, "extremal", 15, 0)
    assert d["dataType"] == 2  # untouched default
    # FitOCT.R:119  ifelse(rho_scale==0, 1./Nn, rho_scale)
    assert api.resolve_rho(d["rho_scale"], d["Nn"]) == pytest.approx(1 / 15)
    assert api.resolve_rho(0.1, 10) == 0.1


def test_synth_matches_synthData_R():
    S = synth.make_profiles(10)
    x = S["x"]
    assert x[0] == 20 and x[-1] == 500 and x.size == 481                     # synthData.R:3
    y0 = 1000 + 2000 * np.exp(-x / 150)
    np.testing.assert_allclose(S["UY"][0], 0.5 * np.sqrt(y0 - 1000 + 1))       # synthData.R:11
    assert list(S["mod_kind"][:5]) == [0, 1, 2, 3, 4]
    for j in range(10):
        m = synth.modulation(int(S["mod_kind"][j]), x)
        z = (S["Y"][j] - (1000 + 2000 * np.exp(-x / (150 * (1 + m))))) / S["UY"][j]
        assert abs(z.mean()) < 0.2 and 0.85 < z.std() < 1.15
    np.testing.assert_allclose(synth.modulation(1, x), 10 * np.sin(x / 50) / x)
    np.testing.assert_allclose(synth.modulation(4, x), np.sin((x - 250) / 20) / (x - 250 + 0.1))
    # a profile does not depend on the batch it was generated in
    S2 = synth.make_profiles(3, first_id=7)
    np.testing.assert_array_equal(S2["Y"][0], S["Y"][7])
    # config 5 uses modulated profiles only (the br gate skips good mono-exp fits, FitOCT.R:100)
    assert set(synth.make_profiles(8, modulated_only=True)["mod_kind"]) == {1, 2, 3, 4}


def test_problem_marshalling_dense_equals_listwise():
    S = synth.make_profiles(3)
    dense = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=7, gridType=1, rho=0.2,
                                    lambda_rate=0.3, prior_PD=1, ids=[5, 6, 9], dataType=1)
    lst = abi.make_problems([dict(x=S["x"], y=S["Y"][j], uy=S["UY"][j], theta0=S["theta0"][j], Sigma0=S["Sigma0"][j], Nn=7,
                                  gridType="extremal", rho=0.2, lambda_rate=0.3, prior_PD=1, id=[5, 6, 9][j], dataType=1)
                             for j in range(3)])
    for j in range(3):
        a, b = dense.array[j], lst.array[j]
        for f, _ in abi.Problem._fields_:
            if f in ("x", "y", "uy"):
                np.testing.assert_array_equal(np.ctypeslib.as_array(getattr(a, f), (a.N,)),
                                              np.ctypeslib.as_array(getattr(b, f), (b.N,)))
            elif f in ("theta0", "Sigma0"):
                assert list(getattr(a, f)) == list(getattr(b, f))
            else:
                assert getattr(a, f) == getattr(b, f), f
    with pytest.raises(ValueError):
        abi.make_problems([dict(x=[1, 2, 3], y=[1, 2], uy=[1, 2, 3])])


def test_param_names_follow_plotExpGP():
    # plotExpGP.R:9,41: theta, yGP, lambda, sigma, br, lp__
    n = abi.param_names(abi.FOCT_EXPGP, 3)
    assert n == ["theta[1]", "theta[2]", "theta[3]", "yGP[1]", "yGP[2]", "yGP[3]", "lambda", "sigma", "br", "lp__"]
    assert abi.param_names(abi.FOCT_MONOEXP, 0) == ["theta[1]", "theta[2]", "theta[3]", "br", "lp__"]
    assert abi.dims(abi.FOCT_EXPGP, 10) == (15, 17)


def test_stanfit_accessors():
    names = abi.param_names(abi.FOCT_EXPGP, 2)
    rng = np.random.default_rng(0)
    draws = rng.standard_normal((30, 4, len(names)))
    fit = api.StanFit(names, draws, np.zeros((30, 4, 6)), n_warmup=10, n_iter=30, save_warmup=True,
                      summary_table=np.zeros((len(names), 11)), stepsize=np.ones(4), inv_metric=np.ones((4, 7)),
                      n_divergent=np.zeros(4))
    ex = fit.extract("br")
    assert ex["br"].shape == (80,)                        # 20 post-warm-up x 4 chains (plotExpGP.R:11)
    assert fit.extract(["theta", "yGP"])["theta"].shape == (80, 3)
    assert fit.as_matrix(["theta", "lp__"]).shape == (80, 4)   # plotExpGP.R:41-44
    assert fit.extract("sigma", inc_warmup=True)["sigma"].shape == (120,)
    assert fit.summary(["lambda", "sigma"])["rownames"] == ["lambda", "sigma"]
    with pytest.raises(KeyError):
        fit.extract("nope")
    assert "theta[1]" in str(fit)


def test_fitExpGP_argument_errors():
    x = synth.depth_grid()
    with pytest.raises(ValueError):
        api.fitExpGP(x, x, x, theta0=None, Sigma0=None)
    with pytest.raises(ValueError):
        api.fitExpGP(x, x, x, theta0=[1, 1, 1], Sigma0=np.eye(3), method="bogus")
    with pytest.raises(TypeError):
        api.fitExpGP(x, x + 1, x, theta0=[1, 1, 1], Sigma0=np.eye(3), method="vb", control=dict(tol_rel_obj="tight"))


def test_shard_ranges_partition_the_batch():
    for n in (1, 7, 1000, 100000):
        for w in (1, 2, 4, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)


def test_r_shim_compiles_against_the_public_header(tmp_path):
    """R is absent from the image: the `.Call` shim is compiled against include/fitoct_b200.h and the stand-in
    R runtime of tests/r_stub (declarations of exactly the R API the shim uses), so a signature drift in the ABI or a
    misuse of the R API is caught here; tests/test_gpu_rshim.py then RUNS it on the GPU box."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    flags = ["-std=gnu99", "-Wall", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
             "-Werror=int-conversion", "-I", os.path.join(root, "tests", "r_stub"), "-I", os.path.join(root, "include")]
    for src in (os.path.join(root, "r-pkg", "src", "shim.c"), os.path.join(root, "tests", "r_stub", "rstub.c"),
                os.path.join(root, "tests", "r_stub", "drive_shim.c")):
        r = subprocess.run([cc, *flags, "-c", src, "-o", str(tmp_path / (os.path.basename(src) + ".o"))], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    # every .Call entry the R wrappers use is registered
    shim = open(os.path.join(root, "r-pkg", "src", "shim.c")).read()
    rsrc = open(os.path.join(root, "r-pkg", "R", "fit.R")).read()
    import re
    for name in set(re.findall(r'\.Call\("(foct_R_\w+)"', rsrc)):
        assert '{"%s", (DL_FUNC)&%s,' % (name, name) in shim, name
