"""GPU (-m gpu): continuation of a run (ABI 2: inv_metric_init / stepsize_init / iter_offset / last_q), the
run-until-converged rounds (rhat_target), and a deterministic GPU-vs-oracle comparison across the warm-up windows.

The reference has no continuation (rstan cannot extend a stanfit; FitOCT.R:43-44 raises nb_warmup / nb_sample and
refits), so these are property tests plus oracle identities:
  * a run split in two (k iterations, then a continuation of the rest) is BIT-IDENTICAL to the unsplit run — the Philox
    sites are indexed by the global iteration number and the whole sampler state is (q, inv_metric, stepsize);
  * the continuation builds the same trees as the oracle's continuation;
  * the thinned draws of a continued profile are exactly the draws of an explicit continuation.
"""
import ctypes as C

import numpy as np
import pytest

from fitoct_b200 import _abi as abi
from fitoct_b200 import synth

pytestmark = pytest.mark.gpu


def batch_of(n, Nn=10, first_id=0):
    S = synth.make_profiles(n, modulated_only=True, first_id=first_id)
    return S, abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=Nn, ids=S["ids"])


@pytest.mark.parametrize("Nn", [10, 15])   # two chains per warp / one chain per warp
def test_split_run_is_bit_identical_to_the_unsplit_run(L, Nn, sampling_kernel):
    _, b = batch_of(3, Nn)
    spec = abi.default_spec()
    full = L.sample(0, b, 3, spec, abi.default_cfg(n_warmup=40, n_iter=90, seed=7))
    first = L.sample(0, b, 3, spec, abi.default_cfg(n_warmup=40, n_iter=60, seed=7))
    np.testing.assert_array_equal(first["draws"], full["draws"][:, :20])
    keep = []
    cfg2 = L.continuation_cfg(abi.default_cfg(n_warmup=40, n_iter=60, seed=7), first, n_more=30, iters_done=60, keep=keep)
    rest = L.sample(0, b, 3, spec, cfg2)
    np.testing.assert_array_equal(rest["draws"], full["draws"][:, 20:])
    np.testing.assert_array_equal(rest["sampler_params"], full["sampler_params"][:, 20:])
    np.testing.assert_array_equal(rest["last_q"], full["last_q"])
    np.testing.assert_array_equal(rest["inv_metric"], full["inv_metric"])   # no adaptation in a continuation
    np.testing.assert_array_equal(rest["stepsize"], full["stepsize"])
    # last_q is the unconstrained image of the last draw
    D = Nn + 5
    last = full["draws"][:, -1, :, :D].copy()
    last[..., Nn + 3:] = np.log(last[..., Nn + 3:])
    np.testing.assert_allclose(full["last_q"], last, rtol=1e-13, atol=1e-15)


def test_continuation_builds_the_same_trees_as_the_oracle(L, O, sampling_kernel):
    _, b = batch_of(2, 10, first_id=11)
    spec = abi.default_spec()
    cfg = abi.default_cfg(n_warmup=30, n_iter=40, seed=99)
    first = O.sample(0, b, 2, spec, cfg)     # both continuations start from the ORACLE's adapted state
    keep = []
    cfg2 = L.continuation_cfg(cfg, first, n_more=8, iters_done=40, keep=keep)
    out = L.sample(0, b, 2, spec, cfg2)
    ref = O.sample(0, b, 2, spec, cfg2)
    K = 5
    np.testing.assert_array_equal(out["sampler_params"][:, :K, :, 2:5], ref["sampler_params"][:, :K, :, 2:5])
    np.testing.assert_allclose(out["draws"][:, :K], ref["draws"][:, :K], rtol=1e-7, atol=1e-9)
    np.testing.assert_array_equal(out["stepsize"], first["stepsize"])


@pytest.mark.parametrize("Nn,n_warmup", [(10, 150), (10, 40), (15, 150)])
def test_adaptation_windows_match_the_oracle(L, O, Nn, n_warmup, sampling_kernel):
    """Deterministic comparison ACROSS the metric windows (tree identity stopped at transition 6, before
    any window closed).  Shallow trees (max_treedepth 3) and a cautious step size (adapt_delta 0.95) keep the floating-
    point chaos of the trajectories small — on the CPU restatement a 1e-15 perturbation of the data grows to 1e-11 over
    these 160 transitions — so the GPU and the oracle stay on ONE path through init_buffer, the metric window
    (regularised Welford variance), the step-size re-initialisation, the dual-averaging restart and the final
    exp(x_bar), and every output can be compared directly."""
    _, b = batch_of(2, Nn, first_id=5)
    spec = abi.default_spec()
    n_iter = n_warmup + 10
    cfg = abi.default_cfg(n_warmup=n_warmup, n_iter=n_iter, seed=31, save_warmup=1, max_treedepth=3, adapt_delta=0.95)
    out = L.sample(0, b, 2, spec, cfg)
    ref = O.sample(0, b, 2, spec, cfg)
    sp, spr = out["sampler_params"], ref["sampler_params"]
    # window schedule (Stan: 75 / 25 / 50; 15 % / 75 % / 10 % below 150 warm-up iterations): one window, closing here
    close = 99 if n_warmup >= 150 else int(0.15 * n_warmup) + (n_warmup - int(0.15 * n_warmup) - int(0.1 * n_warmup)) - 1
    np.testing.assert_array_equal(sp[..., 2:5], spr[..., 2:5])                         # depth, n_leapfrog, divergent
    np.testing.assert_allclose(sp[..., 1], spr[..., 1], rtol=1e-7)                     # step size of every transition
    np.testing.assert_allclose(sp[..., [0, 5]], spr[..., [0, 5]], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(out["draws"], ref["draws"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(out["inv_metric"], ref["inv_metric"], rtol=1e-7)
    np.testing.assert_allclose(out["stepsize"], ref["stepsize"], rtol=1e-7)
    # the comparison did cross what it claims to cross: the metric left the unit matrix at the window close, where the
    # step size was re-initialised (a jump by a power of two from init_stepsize, not a dual-averaging step), and the
    # sampling phase runs at exp(x_bar)
    assert np.all(out["inv_metric"] != 1.0)
    jump = sp[:, close + 1, :, 1] / sp[:, close, :, 1]
    assert np.all(np.abs(np.log(jump)) > 0.05)
    np.testing.assert_array_equal(sp[:, n_warmup:, :, 1], np.broadcast_to(out["stepsize"][:, None, :], sp[:, n_warmup:, :, 1].shape))


def test_run_until_converged(L, O, sampling_kernel):
    n = 12
    _, b = batch_of(n, 10, first_id=40)
    spec = abi.default_spec()
    base = dict(n_warmup=150, n_iter=250, seed=5)
    plain = L.sample(0, b, n, spec, abi.default_cfg(**base))
    rh0 = np.nanmax(plain["summary"][:, :15, 9], axis=1)
    target = float(np.sort(rh0)[n // 2])           # half of the profiles are above the target after the plain run
    cfg = abi.default_cfg(**base)
    cfg.rhat_target, cfg.max_extend, cfg.extend_iter = target, 3, 40
    out = L.sample(0, b, n, spec, cfg)
    ne = out["n_extend"]
    rh = np.nanmax(out["summary"][:, :15, 9], axis=1)
    # exactly the profiles above the target were continued; every profile ends below it or at the round limit
    np.testing.assert_array_equal(ne > 0, rh0 >= target)
    assert np.all((rh < target) | (ne == 3))
    assert np.all(rh[ne > 0] <= rh0[ne > 0] + 0.02)
    # untouched profiles are bit-identical to the plain run
    same = ne == 0
    np.testing.assert_array_equal(out["draws"][same], plain["draws"][same])
    np.testing.assert_array_equal(out["summary"][same], plain["summary"][same])
    # a continued profile: its returned draws are thinned evenly out of (plain run ++ explicit continuation), and its
    # summary is the summary of all of them
    j = int(np.argmax(ne))
    e = int(ne[j])
    keep = []
    cfg2 = L.continuation_cfg(abi.default_cfg(**base), plain, n_more=40 * e, iters_done=250, keep=keep)
    cont = L.sample(0, b, n, spec, cfg2)
    allj = np.concatenate([plain["draws"][j], cont["draws"][j]], axis=0)
    total = 100 + 40 * e
    rows = ((np.arange(100) + 1) * total) // 100 - 1
    np.testing.assert_array_equal(out["draws"][j], allj[rows])
    np.testing.assert_allclose(out["summary"][j], O.summary(allj), rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(out["last_q"][j], cont["last_q"][j])
    # leapfrog counts accumulate over the rounds
    assert np.all(out["n_leapfrog"][ne > 0, :, 1] > plain["n_leapfrog"][ne > 0, :, 1])
    np.testing.assert_array_equal(out["n_leapfrog"][j], plain["n_leapfrog"][j] + cont["n_leapfrog"][j])


def test_rhat_target_needs_the_summary(L):
    _, b = batch_of(1)
    cfg = abi.default_cfg(n_warmup=20, n_iter=40)
    cfg.rhat_target, cfg.max_extend = 1.01, 2
    with pytest.raises(L.FitOCTError, match="summary"):
        L.sample(0, b, 1, abi.default_spec(), cfg, draws=True, summary=False)
    cfg.rhat_target = 0.9
    with pytest.raises(L.FitOCTError, match="rhat_target"):
        L.sample(0, b, 1, abi.default_spec(), cfg)


def test_progress_callback_and_cancel(L):
    """foct_sample_cb: the callback runs on the calling thread while the kernels run; a truthy return cancels."""
    _, b = batch_of(40, 10, first_id=300)
    spec = abi.default_spec()
    cfg = abi.default_cfg(n_warmup=200, n_iter=500, seed=3)
    seen = []
    out = L.sample(0, b, 40, spec, cfg, draws=False, progress=lambda f, ph: seen.append((f, ph)) and False, poll_ms=2)
    ref = L.sample(0, b, 40, spec, cfg, draws=False)
    np.testing.assert_array_equal(out["summary"], ref["summary"])          # polling does not change the result
    fr = [f for f, _ in seen]
    assert len(fr) >= 3 and fr == sorted(fr) and fr[-1] == 1.0 and 0.0 <= fr[0] < 1.0
    assert seen[0][1] == "Warmup" and seen[-1][1] == "Sampling"
    calls = []
    with pytest.raises(L.FitOCTError) as ei:
        L.sample(0, b, 40, spec, abi.default_cfg(n_warmup=2000, n_iter=6000, seed=3), draws=False,
                 progress=lambda f, ph: calls.append(f) or len(calls) >= 3, poll_ms=2)
    assert ei.value.code == -5 and len(calls) == 3 and calls[-1] < 0.5
    again = L.sample(0, b, 40, spec, cfg, draws=False)                     # the library is fine after a cancelled run
    np.testing.assert_array_equal(again["summary"], ref["summary"])
    # the plan-level form
    plan = L.Plan(0, b, 40, spec, abi.default_cfg(n_warmup=2000, n_iter=6000, seed=3), want_draws=False, want_summary=True)
    plan.run()
    done, f0 = plan.query()
    assert not done and 0.0 <= f0 < 1.0
    plan.cancel()
    with pytest.raises(L.FitOCTError) as ei:
        plan.sync()
    assert ei.value.code == -5
    plan.close()


@pytest.mark.parametrize("ticks", [37, 500])
def test_time_sliced_work_items_are_bit_identical(L, monkeypatch, ticks):
    """nuts2_kernel's time slicing (more work items than resident CTAs): a fit that is suspended after `ticks` gradient
    evaluations — in the middle of a tree, a step-size trial or an adaptation window — stored, queued and resumed by
    whichever CTA is free gives exactly the draws, sampler parameters, adapted state and counts of the uninterrupted fit."""
    n = 24
    _, b = batch_of(n, 10, first_id=300)   # one depth grid for the batch: the shared-basis kernel
    spec = abi.default_spec()
    cfg = abi.default_cfg(n_warmup=60, n_iter=100, seed=21, save_warmup=1)
    monkeypatch.setenv("FOCT_FORCE_PAIR", "1")   # (a batch this small would otherwise run one chain per warp, unsliced)
    monkeypatch.setenv("FOCT_MAX_GRID", "5")     # 24 items on 5 CTAs: every slice ends with other items waiting
    monkeypatch.setenv("FOCT_SLICE_TICKS", "0")
    plain = L.sample(0, b, n, spec, cfg)
    monkeypatch.setenv("FOCT_SLICE_TICKS", str(ticks))
    sliced = L.sample(0, b, n, spec, cfg)
    for k in ("draws", "sampler_params", "summary", "stepsize", "inv_metric", "last_q", "n_leapfrog", "n_divergent"):
        np.testing.assert_array_equal(sliced[k], plain[k], err_msg=k)
    assert plain["n_leapfrog"].sum() > 20 * ticks * n   # the fits were long enough to be suspended many times
    # plain round robin instead of the progress-aware rule (who yields does not change what a fit computes)
    monkeypatch.setenv("FOCT_SLICE_PRIO", "0")
    rr = L.sample(0, b, n, spec, cfg)
    monkeypatch.delenv("FOCT_SLICE_PRIO")
    for k in ("draws", "sampler_params", "stepsize", "n_leapfrog"):
        np.testing.assert_array_equal(rr[k], plain[k], err_msg=k)
    # CTA-level work items (the scheduling of the staged-blob variant) on the shared-basis kernel
    monkeypatch.setenv("FOCT_CTA_ITEMS", "1")
    ci = L.sample(0, b, n, spec, cfg)
    monkeypatch.delenv("FOCT_CTA_ITEMS")
    for k in ("draws", "sampler_params", "stepsize", "n_leapfrog"):
        np.testing.assert_array_equal(ci[k], plain[k], err_msg=k)
    # the other variant of the kernel: the whole blob staged per item (ragged depth grids: no shared basis); it sums in
    # the same order
    for env in ({"FOCT_NO_SHARED_BASIS": "1"},):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        monkeypatch.setenv("FOCT_SLICE_TICKS", "0")
        plain2 = L.sample(0, b, n, spec, cfg)
        monkeypatch.setenv("FOCT_SLICE_TICKS", str(ticks))
        sliced2 = L.sample(0, b, n, spec, cfg)
        for k in ("draws", "sampler_params", "stepsize", "inv_metric", "n_leapfrog"):
            np.testing.assert_array_equal(sliced2[k], plain2[k], err_msg=k)
            np.testing.assert_array_equal(plain2[k], plain[k], err_msg=k)
        for k in env:
            monkeypatch.delenv(k)
