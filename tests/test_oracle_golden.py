"""CPU: the C oracle against the committed 50-digit golden vectors (tests/golden/make_golden.py) and the
published Philox4x32-10 known-answer vectors.  Parity unpinned w.r.t. rstan: these pin the
oracle to MODEL_SPEC.md."""
import numpy as np
import pytest

from conftest import case_to_batch, grad_tol_ok
from fitoct_b200 import _abi as abi


def test_philox_known_answers(O):
    # Random123 kat_vectors, philox4x32 with 10 rounds
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert O.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert O.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_grid_matches_server_R(O):
    # ShinyInterface/server.R:626-635: internal seq(dx/2, 1-dx/2, length.out=n), extremal seq(0, 1, length.out=n)
    for n in (1, 2, 5, 10, 15, 20, 25):
        dx = 1.0 / (n + 1)
        np.testing.assert_allclose(O.grid(n, 0), np.linspace(dx / 2, 1 - dx / 2, n) if n > 1 else [dx / 2], rtol=0, atol=1e-15)
        np.testing.assert_allclose(O.grid(n, 1), np.linspace(0, 1, n) if n > 1 else [0.0], rtol=0, atol=1e-15)


def test_golden_cases_cover_switches(golden):
    specs = [c["spec"] for c in golden]
    for key in ("modulation", "kernel", "ygp_prior", "lambda_prior", "theta_prior", "br_ndf"):
        assert {s[key] for s in specs} == {0, 1}, key
    assert {c["prior_PD"] for c in golden} == {0, 1}
    assert {c["kind"] for c in golden} == {0, 1}
    assert {c["dataType"] for c in golden} == {1, 2}


@pytest.mark.parametrize("idx", range(7))
def test_oracle_logp_grad_vs_golden(O, golden, idx):
    case = golden[idx]
    batch, spec = case_to_batch(case)
    B = np.array(case["basis"]) if case["kind"] == 0 else None
    q = np.array(case["q"])
    lp, g, chi2, at = O.logp_grad(case["kind"], batch, 0, spec, q, B=B, want_abs=True)
    for k, e in enumerate(case["expected"]):
        lp_ref, g_ref = float(e["lp"]), np.array([float(v) for v in e["grad"]])
        assert abs(lp[k] - lp_ref) <= 1e-13 * abs(lp_ref) + 1e-300
        assert grad_tol_ok(g[k], g_ref, at[k], 1e-13)
        if not case["prior_PD"]:
            assert abs(chi2[k] - float(e["chi2"])) <= 1e-13 * float(e["chi2"])
        else:
            assert np.isnan(chi2[k])


@pytest.mark.parametrize("idx", range(5))
def test_oracle_basis_vs_golden(O, golden, idx):
    case = golden[idx]
    batch, spec = case_to_batch(case)
    B = O.basis(batch, 0, spec)
    Bref = np.array(case["basis"])
    # Cholesky of a Gaussian Gram matrix: error ~ cond(Kgg) * eps; cond reaches 1e8 at Nn=20, rho=0.05
    assert np.abs(B - Bref).max() <= 2e-7 * np.abs(Bref).max()
    # rows of the GP conditional mean of a constant reproduce the constant up to the jitter
    assert np.allclose(B.sum(axis=0).mean(), 1.0, atol=0.2)


def test_basis_interpolates_control_points(O):
    # B evaluated AT a control point is the unit vector (conditional mean of a noise-free GP), up to jitter
    Nn = 6
    xg = O.grid(Nn, 1)
    x = np.concatenate([xg * 100.0 + 7.0, [7.0, 107.0]])
    x = np.unique(x)
    prof = dict(x=x, y=np.ones_like(x), uy=np.ones_like(x), Nn=Nn, gridType=1, rho=0.3, theta0=(1, 1, 1), Sigma0=np.eye(3))
    batch = abi.make_problems([prof])
    B = O.basis(batch, 0, abi.default_spec())
    for k in range(Nn):
        i = int(np.argmin(np.abs(x - (xg[k] * 100.0 + 7.0))))
        e = np.zeros(Nn); e[k] = 1.0
        np.testing.assert_allclose(B[:, i], e, atol=1e-5)


def test_oracle_gradient_matches_finite_differences(O, golden):
    case = golden[0]
    batch, spec = case_to_batch(case)
    q = np.array(case["q"][:1])
    _, g, _ = O.logp_grad(0, batch, 0, spec, q)
    for d in range(q.shape[1]):
        h = 1e-6 * max(1.0, abs(q[0, d]))
        qp, qm = q.copy(), q.copy()
        qp[0, d] += h; qm[0, d] -= h
        fd = (O.logp_grad(0, batch, 0, spec, qp)[0][0] - O.logp_grad(0, batch, 0, spec, qm)[0][0]) / (2 * h)
        assert abs(fd - g[0, d]) <= 1e-6 * max(abs(fd), 1.0)


def test_predict_consistent_with_chi2(O, golden):
    case = golden[0]
    batch, spec = case_to_batch(case)
    q = np.array(case["q"])
    _, _, chi2 = O.logp_grad(0, batch, 0, spec, q)
    Nn = case["Nn"]
    rows = np.zeros((2, Nn + 7))
    rows[:, :3 + Nn] = q[:, :3 + Nn]
    rows[:, 3 + Nn] = np.exp(q[:, 3 + Nn]); rows[:, 4 + Nn] = np.exp(q[:, 4 + Nn])
    m, resid, dl = O.predict(0, batch, 0, spec, rows)
    uy = np.array(case["uy"])
    np.testing.assert_allclose(((resid / uy) ** 2).sum(axis=1), chi2, rtol=1e-12)
    np.testing.assert_allclose(m + resid, np.broadcast_to(np.array(case["y"]), m.shape), rtol=1e-14)
