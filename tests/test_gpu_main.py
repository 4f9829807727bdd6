"""GPU end-to-end tests of ``ml_sampler_core`` (linna/main.py:77-334): the reference's own smoke test
(tests/test_main.py:8-43 -- 2-D Gaussian, identity theory, ntrain=20) and a posterior-moment check against the
analytic Gaussian posterior (docs/notebooks/multivariate_gaussian_distribution.ipynb cells 8-9; SURVEY 8c: the
shipped chain cannot be read without emcee/h5py, so moments are pinned analytically)."""
import os
from copy import deepcopy

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna.main import ml_sampler_core
from linna.nn import ChtoModelv2

pytestmark = pytest.mark.gpu


def theory(x, outdirs):
    return deepcopy(x[1])


def _priors(ndim, lo, hi):
    return [{"param": "test_{0}".format(i), "dist": "flat", "arg1": lo, "arg2": hi} for i in range(ndim)]


def test_main_reference_smoke(tmp_path):
    np.random.seed(0)
    ndim = 2
    init = np.random.uniform(size=ndim)
    cov, means = np.diag([0.5, 0.2]), np.array([0.1, 1.0])
    params = {"trainingoption": 1, "num_epochs": 10, "batch_size": 5}
    outdir = str(tmp_path / "2dgaussian_Fulltconn") + "/"
    chain, logprob = ml_sampler_core([20], [5], [1], [2], [0.5], [100], [100], outdir, theory, _priors(ndim, -2.0, 2.0), means,
                                     cov, init, None, 4, "cuda", None, False, [1.0], omegab2cut=None, docuda=False, tsize=1,
                                     gpunode=None, nnmodel_in=ChtoModelv2, params=params, method="emcee")
    it = os.path.join(outdir, "iter_0")
    for f in ("train_samples_x.txt", "train_samples_y.npy", "val_samples_x.txt", "val_samples_y.npy", "X_transform.pkl",
              "y_transform.pkl", "y_invtransform_data.pkl", "model_pickle.pkl", "model_args.pkl", "best.pth.tar", "finish.pkl",
              "chemcee_256.npz"):
        assert os.path.isfile(os.path.join(it, f)), f
    assert chain.ndim == 2 and chain.shape[1] == ndim and np.all(np.isfinite(chain))
    assert np.all(np.abs(chain) <= 2.0)                       # flat prior box
    assert len(np.asarray(logprob).reshape(-1)) >= len(chain)
    # a second call finds everything on disk and only reads the chain back
    chain2, _ = ml_sampler_core([20], [5], [1], [2], [0.5], [100], [100], outdir, theory, _priors(ndim, -2.0, 2.0), means, cov,
                                init, None, 4, "cuda", None, False, [1.0], params=params, method="emcee")
    assert np.array_equal(chain, chain2)


def test_main_posterior_moments(tmp_path):
    """One iteration at T = 1 on a 3-D Gaussian with an identity theory: the emulator-driven chain has to reproduce
    the analytic posterior N(means, cov) (flat priors far away)."""
    np.random.seed(1)
    torch.manual_seed(1)
    ndim = 3
    means = np.array([0.3, -0.5, 0.8])
    cov = np.diag([0.04, 0.09, 0.0225])
    params = {"trainingoption": 1, "num_epochs": 300, "batch_size": 200}
    outdir = str(tmp_path / "gauss3") + "/"
    chain, logprob = ml_sampler_core([3000], [200], [8], [25], [0.05], [0.2], [0.2], outdir, theory, _priors(ndim, -3.0, 3.0),
                                     means, cov, means + 0.05, None, 32, "cuda", None, False, [1.0], params=params,
                                     method="emcee")
    assert len(chain) > 3000
    sd = np.sqrt(np.diag(cov))
    assert np.all(np.abs(chain.mean(axis=0) - means) < 0.35 * sd), (chain.mean(axis=0), means)
    assert np.all(np.abs(chain.std(axis=0) / sd - 1.0) < 0.15), (chain.std(axis=0), sd)


def test_main_second_iteration_trains_on_the_previous_chain(tmp_path):
    """Two iterations (T = 4, then 1): iteration 1 draws its training parameters from iteration 0's chain
    (params['trainingoption'] = 1, linna/util.py:865-897) and trains on the union of both sets."""
    np.random.seed(2)
    torch.manual_seed(2)
    ndim = 2
    means, cov = np.array([0.1, 1.0]), np.diag([0.5, 0.2])
    params = {"trainingoption": 1, "num_epochs": 20, "batch_size": 50}
    outdir = str(tmp_path / "two_iter") + "/"
    chain, logprob = ml_sampler_core([200, 200], [20, 20], [1, 1], [2, 2], [0.5, 0.5], [100, 100], [100, 100], outdir, theory,
                                     _priors(ndim, -2.0, 2.0), means, cov, means, None, 8, "cuda", None, False, [2.0, 1.0],
                                     params=params, method="emcee")
    prev = np.load(os.path.join(outdir, "iter_0", "chemcee_256.npz"))["chain_transformed"].reshape(-1, ndim)
    x1 = np.loadtxt(os.path.join(outdir, "iter_1", "train_samples_x.txt"))
    assert x1.shape == (200, ndim)
    assert np.all(x1.min(axis=0) >= prev.min(axis=0) - 1e-6) and np.all(x1.max(axis=0) <= prev.max(axis=0) + 1e-6)
    assert os.path.isfile(os.path.join(outdir, "iter_1", "best.pth.tar"))
    assert np.all(np.isfinite(chain)) and chain.shape[1] == ndim


def test_integrated_time_gpu_matches_host():
    """The batched-FFT autocorrelation time on the device equals the per-walker host loop (emcee's estimator)."""
    from linna_b200 import sampler
    rng = np.random.default_rng(3)
    n, w, d = 3000, 12, 4
    x = np.zeros((n, w, d))
    e = rng.standard_normal((n, w, d))
    rho = np.array([0.5, 0.8, 0.9, 0.95])
    for t in range(1, n):
        x[t] = rho * x[t - 1] + e[t]
    host = sampler.integrated_time(x)
    dev = sampler.integrated_time(torch.from_numpy(x).cuda())
    np.testing.assert_allclose(dev, host, rtol=1e-8)
    assert np.all(np.abs(host / ((1 + rho) / (1 - rho)) - 1) < 0.25)


def test_stretch_move_kernels_sample_a_gaussian():
    """The on-device stretch move (csrc/sampler_kernels.cu) leaves a correlated 5-D Gaussian invariant: moments of
    the chain against the analytic ones, proposals inside the stretch-move envelope, acceptance near emcee's."""
    from linna_b200 import engine as E, sampler
    d, W = 5, 512
    rng = np.random.default_rng(0)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + np.eye(d)
    icov = torch.from_numpy(np.linalg.inv(cov).astype(np.float32)).cuda()
    mu = torch.from_numpy(rng.standard_normal(d).astype(np.float32)).cuda()

    def lnp(x):
        r = x - mu
        return -0.5 * torch.einsum("ij,jk,ik->i", r, icov, r)
    x = torch.from_numpy(rng.standard_normal((W, d)).astype(np.float32)).cuda()
    first = torch.arange(0, W // 2, device="cuda")
    second = torch.arange(W // 2, W, device="cuda")
    y, z = E.stretch_propose(x, first, second, 2.0, 7, 0)
    zz = z.cpu().numpy()
    assert zz.min() >= 0.5 - 1e-6 and zz.max() <= 2.0 + 1e-6 and abs(zz.mean() - 7.0 / 6.0) < 0.1   # E[z] under g(z), a = 2
    # every proposal lies on the line through its walker and ONE walker of the complementary set
    yy, xx = y.cpu().numpy(), x.cpu().numpy()
    for i in (0, 17, 255):
        partner = (yy[i] - zz[i] * xx[i]) / (1 - zz[i]) if abs(1 - zz[i]) > 1e-3 else None
        if partner is not None:
            assert np.min(np.max(np.abs(xx[W // 2:] - partner), axis=1)) < 1e-3
    es = sampler.EnsembleSampler(W, d, lnp, seed=3)
    xb, _ = es.run_mcmc(x, 300, store=False)      # burn-in
    es.reset()
    es.run_mcmc(xb, 600)
    chain = es.get_chain().reshape(-1, d)
    np.testing.assert_allclose(chain.mean(axis=0), mu.cpu().numpy(), atol=0.08)
    np.testing.assert_allclose(np.cov(chain, rowvar=False), cov, atol=0.15)
    assert 0.35 < es.acceptance_fraction.mean() < 0.75
