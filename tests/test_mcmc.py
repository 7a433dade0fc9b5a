"""Ensemble MCMC (BASELINE config 4): the stretch move of csrc/mcmc_api.cuh against its numpy restatement (oracle/mcmc_ref.py)."""
import os

import numpy as np
import pytest

from conftest import pkg


def test_generator_is_splitmix64():
    """The stateless generator of the sampler: known splitmix64 outputs, uniformity, independence of the draws."""
    from oracle import mcmc_ref as mr

    # splitmix64 stream from state 0 (public test vector of the algorithm): first output
    assert int(mr.mix64(np.uint64(0))) == 0xE220A8397B1DCDAF
    u = mr.u01(12345, 3, 1, np.arange(200_000), 2)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    v = mr.u01(12345, 3, 1, np.arange(200_000), 1)
    assert abs(np.corrcoef(u, v)[0, 1]) < 0.01
    assert not np.array_equal(u, mr.u01(12346, 3, 1, np.arange(200_000), 2))


def test_oracle_stretch_move_samples_a_gaussian():
    """The oracle's move is a valid MCMC: an isotropic Gaussian target is recovered (mean, variance) by the ensemble."""
    from oracle import mcmc_ref as mr

    rng = np.random.default_rng(0)
    n, d = 400, 3
    x = rng.standard_normal((n, d)) * 0.1
    lp = lambda y: -0.5 * (y * y).sum(axis=1)  # noqa: E731
    logp = lp(x)
    keep = []
    for step in range(600):
        for half in (0, 1):
            mr.half_step(x, logp, half, 99, step, lp)
        if step >= 200 and step % 10 == 0:
            keep.append(x.copy())
    s = np.concatenate(keep)
    assert np.all(np.abs(s.mean(axis=0)) < 0.05) and np.all(np.abs(s.var(axis=0) - 1.0) < 0.1)


def _setup(trained_fixture, rm, walkers):
    emu = pkg("emulator")
    pp = pkg("preprocess")
    f = trained_fixture
    e = emu.DirectEmulator(stats=pp.NormStats(f["pmin"], f["pmax"], f["mu"], f["sd"]))
    e.load_model(f["path"])
    truth = f["par_test"][3]
    obs = (rm.predict(truth, f["kernels"], f["biases"], f["relu"], f["pmin"], f["pmax"], f["mu"], f["sd"])
           + np.random.default_rng(5).normal(size=451) * 20.0).astype(np.float32)
    return e, f, truth, obs, None


def _box(f, pp):
    """The walkers' box: the training range in the coordinates par_transform maps to [-1, 1] (preprocess.py:74-78, :105-108) --
    NormStats.par_min / par_max are already in those (log10 on fstar, Vc, fx) coordinates."""
    logm = np.array([1, 1, 1, 0, 0, 0, 0], dtype=bool)
    return np.asarray(f["pmin"], np.float64), np.asarray(f["pmax"], np.float64), logm


@pytest.mark.gpu
def test_one_step_matches_the_oracle(rm, trained_fixture):
    """One full stretch-move step of 2,048 walkers on the GPU (FP32 path) against the numpy restatement with the float64 oracle
    as likelihood: walkers whose accept decision is not marginal (and whose partner's was not) end at bit-identical positions."""
    import torch
    from oracle import mcmc_ref as mr

    mc = pkg("mcmc")
    pp = pkg("preprocess")
    e, f, truth, obs, _ = _setup(trained_fixture, rm, 2048)
    h = e._handle()
    lo, hi, logm = _box(f, pp)
    t_truth = np.where(logm, np.log10(truth), truth)
    s = mc.StretchMoveSampler(e, obs, 20.0, lo, hi, walkers=2048, seed=11, precision="fp32")
    s.ball(t_truth, 0.02 * (hi - lo))
    x0 = s.x.cpu().numpy().copy()

    def log_prob(y):
        inside = np.all((y >= lo) & (y <= hi), axis=1)
        phys = np.where(logm, 10.0 ** y, y)
        pred = rm.predict(phys, f["kernels"], f["biases"], f["relu"], f["pmin"], f["pmax"], f["mu"], f["sd"], squeeze=False)
        chi = (((pred - obs.astype(np.float64)) / 20.0) ** 2).sum(axis=1)
        return np.where(inside, -0.5 * chi, -np.inf)

    frac = s.run(1)
    x1 = s.x.cpu().numpy()
    lp1 = s.logp.cpu().numpy()
    xo = x0.copy()
    lpo = log_prob(xo)
    m = len(xo) // 2
    acc0, mar0 = mr.half_step(xo, lpo, 0, s.seed, 0, log_prob)
    _, _, j1 = mr.propose(xo, 1, s.seed, 0)
    acc1, mar1 = mr.half_step(xo, lpo, 1, s.seed, 0, log_prob)
    safe0 = np.abs(mar0) > 0.5
    safe1 = (np.abs(mar1) > 0.5) & safe0[j1]  # half 1 pairs with the UPDATED half 0
    assert safe0.mean() > 0.8 and safe1.mean() > 0.6
    assert np.array_equal(x1[:m][safe0], xo[:m][safe0])
    assert np.array_equal(x1[m:][safe1], xo[m:][safe1])
    ok = np.concatenate([safe0, safe1])
    assert np.allclose(lp1[ok], lpo[ok], rtol=1e-4, atol=2e-2)
    assert abs(frac - (acc0.sum() + acc1.sum()) / len(xo)) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "fp16e4m3"])
def test_runs_are_reproducible_and_resumable(rm, trained_fixture, prec):
    """Same seed -> bit-identical chains; 3 + 2 steps = 5 steps; the ensemble concentrates around the truth."""
    mc = pkg("mcmc")
    pp = pkg("preprocess")
    e, f, truth, obs, _ = _setup(trained_fixture, rm, 4096)
    if prec != "fp32" and not e._handle().info()["tc_supported"]:
        pytest.skip("tensor-core kernel not available")
    lo, hi, logm = _box(f, pp)
    t_truth = np.where(logm, np.log10(truth), truth)

    def chain(parts):
        s = mc.StretchMoveSampler(e, obs, 20.0, lo, hi, walkers=4096, seed=3, precision=prec)
        s.ball(t_truth, 0.05 * (hi - lo))
        for k in parts:
            s.run(k)
        return s.x.cpu().numpy(), s.logp.cpu().numpy(), s

    xa, la, _ = chain([5])
    xb, lb, _ = chain([3, 2])
    xc, lc, s = chain([5])
    assert np.array_equal(xa, xc) and np.array_equal(la, lc)
    assert np.array_equal(xa, xb) and np.array_equal(la, lb)
    assert np.all(xa >= lo) and np.all(xa <= hi) and np.all(np.isfinite(la))
    frac = s.run(40)
    assert 0.05 < frac < 0.9
    mean, cov = s.moments()
    assert mean.shape == (7,) and cov.shape == (7, 7) and np.all(np.diag(cov) > 0)
    # chi^2 of the ensemble is of the order of the number of bins (451 bins, 7 parameters)
    assert 300 < float(-2 * s.logp.mean()) < 700
