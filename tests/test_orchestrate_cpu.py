"""CPU tests of the callers around the hot path: training-set generation (linna/util.py:749-897, :1166-1254),
the chain statistics and the chain store (linna/sampler.py:322-387, linna/util.py:68-94)."""
import os
from copy import deepcopy

import numpy as np
import pytest

from linna.util import NN_samplerv1, chisqcut_all, generate_training_point
from linna_b200 import orchestrate, sampler


def test_lhs_center_is_a_latin_hypercube():
    x = orchestrate.lhs_center(5, 40, seed=3)
    assert x.shape == (40, 5)
    for j in range(5):
        assert np.array_equal(np.sort(np.floor(x[:, j] * 40).astype(int)), np.arange(40))   # one point per bin
        np.testing.assert_allclose(np.sort(x[:, j]), (np.arange(40) + 0.5) / 40)            # at the bin centres
    assert np.array_equal(x, orchestrate.lhs_center(5, 40, seed=3))


def test_gensample_flat_covers_the_prior_box():
    pr = [[-2.0, 2.0], [0.5, 1.5], [10.0, 30.0]]
    s = NN_samplerv1("unused", pr).gensample_flat(200)
    assert s.shape == (200, 3)
    for j, (lo, hi) in enumerate(pr):
        assert s[:, j].min() >= lo and s[:, j].max() <= hi
        assert s[:, j].min() < lo + 0.02 * (hi - lo) and s[:, j].max() > hi - 0.02 * (hi - lo)
    cut = NN_samplerv1("unused", pr).gensample_flat(100, omegab2cut=[0, 1, -1.0, 1.0])
    ombh2 = cut[:, 0] * cut[:, 1] ** 2
    assert cut.shape == (100, 3) and np.all((ombh2 > -1.0) & (ombh2 < 1.0))


def test_gensample_flat_log_samples_a_tiny_second_parameter():
    s = NN_samplerv1("unused", [[0.1, 0.5], [1e-9, 5e-9]]).gensample_flat(500)
    assert np.all((s[:, 1] >= 1e-9) & (s[:, 1] <= 5e-9))
    assert abs(np.median(np.log(s[:, 1])) - 0.5 * (np.log(1e-9) + np.log(5e-9))) < 0.05   # uniform in log, not linear


def test_gensample_chain_variants():
    rng = np.random.default_rng(0)
    chain = rng.multivariate_normal([0.3, -0.2], [[0.04, 0.01], [0.01, 0.09]], size=4000)
    ns = NN_samplerv1("unused", [[-2.0, 2.0], [-2.0, 2.0]])
    r = ns.gensample_chain_randomsample(300, chain, 3)
    assert r.shape == (300, 2) and all(any(np.array_equal(row, c) for c in chain[:4000]) for row in r[:5])
    x = ns.gensample_chain(300, chain, 3)
    assert x.shape == (300, 2)
    np.testing.assert_allclose(x.mean(axis=0), chain.mean(axis=0), atol=0.05)
    # half-width of the hypercube along the principal axes = 3 sigma
    w, v = np.linalg.eigh(np.cov(chain, rowvar=False))
    proj = (x - chain.mean(axis=0)) @ v
    assert np.all(np.abs(proj) <= 3 * np.sqrt(w) * (1 + 1e-9))
    assert np.all(np.max(np.abs(proj), axis=0) > 2.9 * np.sqrt(w))


def _theory(x, outdirs):
    return deepcopy(x[1]) * 2.0


def test_generate_training_point_layout(tmp_path):
    out = str(tmp_path / "iter_0") + "/"
    ns = NN_samplerv1(out, [[-1.0, 1.0], [-1.0, 1.0]])
    data, icov = np.zeros(2), np.eye(2)
    generate_training_point(_theory, ns, None, out, 30, 7, data, icov, None)
    for f in ("train_samples_x.txt", "val_samples_x.txt", "train_samples_y.npy", "val_samples_y.npy"):
        assert os.path.isfile(os.path.join(out, f)), f
    assert os.path.isdir(os.path.join(out, "train")) and os.path.isdir(os.path.join(out, "val"))
    x, y = np.loadtxt(os.path.join(out, "train_samples_x.txt")), np.load(os.path.join(out, "train_samples_y.npy"))
    assert x.shape == (30, 2) and np.allclose(y, 2 * x)
    assert np.load(os.path.join(out, "val_samples_y.npy")).shape == (7, 2)
    # second call is a no-op (files exist); the chi^2 cut then removes rows from both files consistently
    generate_training_point(_theory, ns, None, out, 30, 7, data, icov, None, chisqcut=1.0)
    x2, y2 = np.loadtxt(os.path.join(out, "train_samples_x.txt")), np.load(os.path.join(out, "train_samples_y.npy"))
    assert len(x2) == len(y2) < 30 and np.all(np.sum(y2 ** 2, axis=1) < 1.0) and np.allclose(y2, 2 * x2)


def test_generate_training_point_with_optimizer_points(tmp_path):
    out = str(tmp_path / "iter_0") + "/"
    ns = NN_samplerv1(out, [[-1.0, 1.0], [-1.0, 1.0]])
    data, icov = np.array([0.4, -0.2]), np.diag([100.0, 25.0])

    def negloglike(x):
        d = data - _theory([-1, x], None)
        return d.dot(icov.dot(d))
    generate_training_point(_theory, ns, None, out, 20, 10, data, icov, None, negloglike=negloglike, nbest_in=40)
    bx, by = np.loadtxt(os.path.join(out, "best_samples_x.txt")), np.load(os.path.join(out, "best_samples_y.npy"))
    assert bx.shape == (40, 2) and by.shape == (40, 2)
    np.testing.assert_allclose(bx.mean(axis=0), data / 2, atol=0.05)          # scattered around the best fit
    assert np.load(os.path.join(out, "best_samples_y_val.npy")).shape == (20, 2)


def test_integrated_time_of_an_ar1_chain():
    rng = np.random.default_rng(1)
    rho, n, w = 0.9, 20000, 8
    x = np.zeros((n, w, 1))
    e = rng.standard_normal((n, w, 1))
    for t in range(1, n):
        x[t] = rho * x[t - 1] + e[t]
    tau = sampler.integrated_time(x)
    assert abs(tau[0] - (1 + rho) / (1 - rho)) < 2.5          # analytic tau = 19


def test_checkmeanstd_and_chain_store(tmp_path):
    rng = np.random.default_rng(2)
    good = rng.standard_normal((400, 6, 3))
    assert sampler.checkmeanstd(good, 0.2, 0.15)
    drift = good + np.linspace(0, 3, 400)[:, None, None]
    assert not sampler.checkmeanstd(drift, 0.2, 0.15)
    # a host tensor takes the same host path (the device path is the hand-written reduction kernel, tests/test_gpu_main.py)
    import torch
    for x in (good, drift, rng.standard_normal((41, 5, 4))):
        assert sampler.checkmeanstd(torch.from_numpy(x), 0.2, 0.15) == sampler.checkmeanstd(x, 0.2, 0.15)
    # the chain store appends: blocks go to raw files as they come, nothing is rewritten, and an unfinished run can be
    # picked up from them; finalize() writes the single-file .npz with the reference's dataset names
    name = str(tmp_path / "chemcee_256.h5")
    st = sampler.ChainStore(name, transform=lambda c: 2.0 * c)
    st.extend(good[:100], np.zeros((100, 6)))
    size1 = os.path.getsize(str(tmp_path / "chemcee_256.chain.f32"))
    st.extend(good[100:250], np.ones((150, 6)))
    st.save()
    assert os.path.getsize(str(tmp_path / "chemcee_256.chain.f32")) == size1 * 250 // 100
    assert not os.path.exists(str(tmp_path / "chemcee_256.npz"))
    st2 = sampler.ChainStore(name)                       # resume of an unfinished run: from the raw files
    assert st2.exists() and st2.iteration == 250
    np.testing.assert_allclose(st2.get_value("chain_transformed"), 2.0 * good[:250])
    np.testing.assert_allclose(st2.get_last_sample(), good[249])
    st2.extend(good[250:300], np.ones((50, 6)))          # ... and continued
    st2.finalize()
    z = np.load(str(tmp_path / "chemcee_256.npz"))
    assert set(z.files) == {"chain", "chain_transformed", "log_prob"} and z["chain"].shape == (300, 6, 3)
    st3 = sampler.ChainStore(name)                       # a finished chain is read from the .npz
    assert st3.iteration == 300
    st3.extend(good[300:350], np.ones((50, 6)))          # continuing a finished chain converts it back
    assert st3.iteration == 350 and np.allclose(st3.chain[299], good[299]) and np.allclose(st3.chain[349], good[349])
    fresh = sampler.ChainStore(name, fresh=True)
    assert not fresh.exists()
    st = sampler.ChainStore(name, transform=lambda c: 2.0 * c)
    st.extend(good[:250], np.concatenate([np.zeros((100, 6)), np.ones((150, 6))]))
    st.finalize()
    st2 = sampler.ChainStore(name)
    assert st2.get_log_prob(flat=True).shape == (1500,)
    sub = sampler.thin_for_tau(rng.standard_normal((50, 5000, 3)), max_walkers=100)
    assert sub.shape == (50, 100, 3)
    chain, lp, reader = sampler.read_chain_and_cut(name, nk=2, ntimes=20)
    assert chain.shape[1] == 3 and len(chain) % 6 == 0 and reader.iteration == 250
    with pytest.raises(FileNotFoundError):
        sampler.read_chain_and_cut(str(tmp_path / "missing.h5"), 2)
