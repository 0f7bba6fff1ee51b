"""baseline/compare_rstan.py is the tool that turns an off-box rstan run (baseline/run_rstan.R) into the 3-MCSE / R-hat parity
table.  R cannot run here, so the tool is exercised on STAND-IN files: the CPU oracle's draws written in exactly the layout
rstan::summary(...)$summary / as.matrix(stanfit) have through write.csv.  CPU test: oracle (seed A) against oracle (seed B);
GPU test (-m gpu): the CUDA sampler against the oracle on 32 profiles — the posterior-parity test of the north_star
(|z| < 3 at the chance rate, split R-hat < 1.01 after the run-until-converged rounds)."""
import csv
import importlib.util
import os

import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec_ = importlib.util.spec_from_file_location("compare_rstan", os.path.join(ROOT, "baseline", "compare_rstan.py"))
CR = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(CR)


def write_inputs(indir, S, n):
    os.makedirs(indir, exist_ok=True)
    for j in range(n):
        np.savetxt(os.path.join(indir, f"Courbe_{j}.csv"), np.c_[S["x"], S["Y"][j], S["UY"][j]], delimiter=",", header="x,y,uy", comments="")
    np.savetxt(os.path.join(indir, "theta0.csv"), S["theta0"][:n], delimiter=",")


def write_rstan_standin(outdir, res, Nn):
    """What baseline/run_rstan.R writes, from a sampler result dict."""
    os.makedirs(outdir, exist_ok=True)
    names = CR.par_names(Nn)
    for j in range(res["draws"].shape[0]):
        with open(os.path.join(outdir, f"summary_{j}.csv"), "w", newline="") as fh:
            w = csv.writer(fh, quoting=csv.QUOTE_NONNUMERIC)
            w.writerow(["", "mean", "se_mean", "sd", "2.5%", "25%", "50%", "75%", "97.5%", "n_eff", "Rhat"])
            for k, nm in enumerate(names):
                w.writerow([nm] + [float(v) for v in res["summary"][j, k, :10]])
        d = res["draws"][j]                                   # [iter, chain, par] -> as.matrix stacks the chains
        flat = np.concatenate([d[:, c, :] for c in range(d.shape[1])], axis=0)
        with open(os.path.join(outdir, f"draws_{j}.csv"), "w", newline="") as fh:
            w = csv.writer(fh, quoting=csv.QUOTE_NONNUMERIC)
            w.writerow(names)
            w.writerows(flat.tolist())


def batch(S, n, Nn):
    # the same problems compare_rstan.fit_gpu builds from the CSV files
    profs = [dict(x=S["x"], y=S["Y"][j], uy=S["UY"][j], dataType=2, Nn=Nn, gridType=0, rho=1.0 / Nn, lambda_rate=0.1,
                  theta0=S["theta0"][j], Sigma0=np.diag((0.05 * S["theta0"][j]) ** 2), prior_PD=0, id=j) for j in range(n)]
    return abi.make_problems(profs)


def test_compare_tool_on_oracle_standins(tmp_path, O):
    n, Nn = 2, 5
    S = synth.make_profiles(n, modulated_only=True)
    write_inputs(tmp_path / "in", S, n)
    b = batch(S, n, Nn)
    a = O.sample(0, b, n, abi.default_spec(), abi.default_cfg(n_warmup=150, n_iter=450, seed=1))
    c = O.sample(0, b, n, abi.default_spec(), abi.default_cfg(n_warmup=150, n_iter=450, seed=2))
    write_rstan_standin(tmp_path / "out", a, Nn)
    rows, verdict = CR.compare(str(tmp_path / "in"), str(tmp_path / "out"), gpu=c, Nn=Nn, rhat_max=1.1)
    assert len(rows) == n * (Nn + 5) and verdict["tests"] == 4 * len(rows)
    assert verdict["parity"], verdict
    assert verdict["worst_abs_z"] < 5
    # the reader maps rstan's names and columns back exactly
    rs = CR.read_rstan_summary(str(tmp_path / "out" / "summary_1.csv"))
    assert abs(rs["yGP[3]"]["50%"] - a["summary"][1, 5, 5]) < 1e-12 and abs(rs["lp__"]["Rhat"] - a["summary"][1, -1, 9]) < 1e-12
    # and the tool does fail when the two fits differ: shift one side's theta[1] by 10 of its standard deviations
    bad = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in c.items()}
    shift = 10 * bad["summary"][0, 0, 2]
    bad["draws"][0, :, :, 0] += shift
    bad["summary"][0, 0, [0, 3, 4, 5, 6, 7]] += shift
    rows, verdict = CR.compare(str(tmp_path / "in"), str(tmp_path / "out"), gpu=bad, Nn=Nn, rhat_max=1.1)
    assert not verdict["parity"] and verdict["worst_abs_z"] > 5


@pytest.mark.gpu
def test_posterior_parity_32_profiles(tmp_path, L, O):
    """north_star: posterior means / quantiles within 3 MCSE, R-hat < 1.01 — against the CPU restatement (rstan cannot run
    here), on 32 profiles x 15 parameters x (mean + 3 quantiles) = 1920 comparisons."""
    n, Nn = 32, 10
    S = synth.make_profiles(n, modulated_only=True, first_id=900)
    write_inputs(tmp_path / "in", S, n)
    ref = O.sample(0, batch(S, n, Nn), n, abi.default_spec(), abi.default_cfg(n_warmup=500, n_iter=1500, seed=4321))
    write_rstan_standin(tmp_path / "out", ref, Nn)
    rows, verdict = CR.compare(str(tmp_path / "in"), str(tmp_path / "out"), Nn=Nn, n_warmup=500, n_sample=1000, seed=77,
                               rhat_target=1.01, max_extend=12)
    print(verdict)
    assert verdict["tests"] == 1920
    assert verdict["rhat_max_gpu"] < 1.01, verdict
    assert verdict["beyond_3_mcse"] <= verdict["allowed_by_chance"] and verdict["worst_abs_z"] < 5.0, verdict
    assert verdict["parity"]
