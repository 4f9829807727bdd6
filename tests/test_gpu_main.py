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


def test_checkmeanstd_device_reduction_matches_numpy():
    """The half-chain mean / std shift test on the GPU (two-pass float64 reduction kernel, csrc/sampler_kernels.cu)
    against numpy, float32 and float64 chains, ragged shapes."""
    from linna_b200 import engine as E, sampler
    rng = np.random.default_rng(2)
    for shape, dt in (((400, 6, 3), np.float64), ((41, 5, 4), np.float32), ((301, 1000, 30), np.float32), ((64, 7, 300), np.float64)):
        x = (rng.standard_normal(shape) * 1.7 + 3.0).astype(dt)
        xd = torch.from_numpy(x).cuda()
        flat = x.reshape(-1, shape[-1]).astype(np.float64)
        r0, r1 = 5 * shape[1], flat.shape[0] - 3
        m, sd = E.column_moments(xd.reshape(-1, shape[-1]), r0, r1)
        np.testing.assert_allclose(m, flat[r0:r1].mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(sd, flat[r0:r1].std(0), rtol=1e-10)
        half = shape[0] // 2
        a, b = flat[:half * shape[1]], flat[half * shape[1]:]
        ref = (np.median(np.abs(a.mean(0) - b.mean(0)) / b.std(0)), np.median((a.std(0) - b.std(0)) / b.std(0)))
        np.testing.assert_allclose(sampler._halves_shift_device(xd), ref, rtol=1e-8, atol=1e-12)
    good = rng.standard_normal((400, 64, 3))
    drift = good + np.linspace(0, 3, 400)[:, None, None]
    assert sampler.checkmeanstd(torch.from_numpy(good).cuda(), 0.2, 0.15) and not sampler.checkmeanstd(torch.from_numpy(drift).cuda(), 0.2, 0.15)


def test_hmc_kernels_follow_the_reference_leapfrog():
    """linna_hmc_begin / _step / _end against the reference's leapfrog written out in torch (linna/HMCSampler.py:25-59) on a
    Gaussian target: same momenta (read back from the kernel), same trajectory, same Hamiltonians, and an accept decision
    consistent with them for every chain."""
    from linna_b200 import engine as E
    C, d, eps, L = 300, 37, 0.07, 4
    rng = np.random.default_rng(1)
    A = rng.standard_normal((d, d))
    icov = torch.from_numpy((np.linalg.inv(A @ A.T / d + np.eye(d))).astype(np.float32)).cuda()
    mass = torch.from_numpy(rng.uniform(0.5, 2.0, d).astype(np.float32)).cuda()

    def vg(x):
        g = -(x @ icov)
        return 0.5 * (x * g).sum(-1), g
    x = torch.from_numpy(rng.standard_normal((C, d)).astype(np.float32)).cuda()
    lnp, grad = vg(x)
    x_k, lnp_k, grad_k = x.clone(), lnp.clone(), grad.clone()
    p, xn, H0 = E.hmc_begin(x_k, lnp_k, grad_k, mass, eps, 99, 0)
    p0 = p - 0.5 * eps * grad                                  # the momentum the kernel drew
    assert abs(float((p0 / mass.sqrt()).mean())) < 0.05 and abs(float((p0 / mass.sqrt()).std()) - 1.0) < 0.05
    np.testing.assert_allclose(H0.cpu().numpy(), ((0.5 * p0 * p0 / mass).sum(-1) - lnp).cpu().numpy(), rtol=2e-5, atol=1e-4)
    # reference leapfrog from the same momentum
    pr = p0 + 0.5 * eps * grad
    xr = x + eps * pr / mass
    np.testing.assert_allclose(xn.cpu().numpy(), xr.cpu().numpy(), rtol=1e-5, atol=1e-6)
    for i in range(L):
        l, g = vg(xn)
        lr_, gr = vg(xr)
        if i + 1 < L:
            E.hmc_step(p, xn, g, mass, eps)
            pr = pr + eps * gr
            xr = xr + eps * pr / mass
    pr = pr + 0.5 * eps * gr
    H1 = (0.5 * pr * pr / mass).sum(-1) - lr_
    nacc = torch.zeros(C, device="cuda")
    E.hmc_end(x_k, lnp_k, grad_k, xn, l.contiguous(), g.contiguous(), p, mass, H0, eps, 99, 4, nacc)
    np.testing.assert_allclose(xn.cpu().numpy(), xr.cpu().numpy(), rtol=2e-4, atol=2e-5)
    acc = nacc.cpu().numpy() > 0
    dH = (H0 - H1).cpu().numpy()
    assert np.all(acc[dH > 1e-3])                              # an energy decrease is always accepted
    assert 0.5 < acc.mean() <= 1.0
    moved = np.all(x_k.cpu().numpy() == xn.cpu().numpy(), axis=1)
    stayed = np.all(x_k.cpu().numpy() == x.cpu().numpy(), axis=1)
    assert np.array_equal(moved, acc) and np.array_equal(stayed, ~acc)           # position, lnP and gradient follow ONE decision
    np.testing.assert_array_equal(lnp_k.cpu().numpy()[acc], l.cpu().numpy()[acc])
    np.testing.assert_array_equal(lnp_k.cpu().numpy()[~acc], lnp.cpu().numpy()[~acc])
    np.testing.assert_array_equal(grad_k.cpu().numpy()[acc], g.cpu().numpy()[acc])


def test_run_mcmc_hmc_method(tmp_path):
    """run_mcmc(method="hmc") (linna/util.py:1495; broken at HEAD) reaches the batched device HMC and stores
    the chain under the reference's chhmc file name."""
    import pickle
    import shutil
    import linna.util as U
    from tests.helpers import GOLDEN
    d = str(tmp_path / "iter_0")
    shutil.copytree(os.path.join(GOLDEN, "ref_fixture_iter_0"), d)
    pred, yinv = U.retrieve_model(d, 2, 2)
    with open(os.path.join(d, "model_args.pkl"), "rb") as f:
        args = pickle.load(f)
    priors = [dict(param="x%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)]
    tr = U.Transform(priors)
    lp = U.Log_prob(np.asarray(args[6]), np.asarray(args[2]), pred, yinv, tr, 1.0, U.gaussianlogliklihood, nograd=False)
    dd = U.Ddlnp(np.asarray(args[6]), np.asarray(args[2]), pred, yinv, tr, 1.0)
    ns = U.NN_samplerv1(d + "/", [[-2, 2], [-2, 2]])
    np.random.seed(3)
    store = U.run_mcmc(ns, d + "/", "hmc", 2, 256, np.zeros(2), lp, dlnp=None, ddlnp=dd, transform=tr, ntimes=5, tautol=0.2)
    assert os.path.isfile(os.path.join(d, "chhmc.meta.json")) and store.iteration >= 100
    ch = np.asarray(store.chain_transformed)
    assert ch.shape[1:] == (256, 2) and np.all(np.abs(ch) <= 2.0) and np.all(np.isfinite(np.asarray(store.log_prob)))


def _ref_reading_args():
    """The arguments of the reference's own tests/test_main.py (chto/linna)."""
    ndim = 2
    priors = [dict(param="test_%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(ndim)]
    return dict(ntrainArr=[20], nvalArr=[5], nkeepArr=[1], ntimesArr=[2], ntautolArr=[0.5], meanshiftArr=[100], stdshiftArr=[100],
                priors=priors, data=np.array([0.1, 1.0]), cov=np.diag([0.5, 0.2]), init=np.array([0.3, 0.6]), nwalkers=4,
                temperatureArr=[1.0], params=dict(trainingoption=1, num_epochs=10, batch_size=5))


def test_reading_of_the_reference_output_directory(tmp_path):
    """The reference's test_reading (tests/test_main.py:46-52): ml_sampler_core pointed at the output directory the
    reference ships -- trained emulator, finish.pkl and the emcee chain file -- trains nothing, samples nothing, reads the
    chain back and must return the moments the reference asserts."""
    import shutil
    from linna.main import ml_sampler_core
    from linna.nn import ChtoModelv2
    from tests.helpers import GOLDEN
    out = tmp_path / "2dgaussian_Fulltconn"
    shutil.copytree(os.path.join(GOLDEN, "ref_fixture_iter_0"), out / "iter_0")
    shutil.copy(os.path.join(GOLDEN, "ref_chain", "chemcee_256.h5"), out / "iter_0" / "chemcee_256.h5")
    a = _ref_reading_args()
    import hashlib
    digest = lambda n: hashlib.sha1(open(out / "iter_0" / n, "rb").read()).hexdigest()
    before = {n: digest(n) for n in ("best.pth.tar", "last.pth.tar", "chemcee_256.h5", "train_samples_y.npy")}

    def theory(x, outdirs):
        raise AssertionError("the finished directory must not evaluate the theory again")
    chain, logprob = ml_sampler_core(a["ntrainArr"], a["nvalArr"], a["nkeepArr"], a["ntimesArr"], a["ntautolArr"], a["meanshiftArr"],
                                     a["stdshiftArr"], str(out) + "/", theory, a["priors"], a["data"], a["cov"], a["init"], None,
                                     a["nwalkers"], "cuda", None, False, a["temperatureArr"], omegab2cut=None, docuda=False, tsize=1,
                                     gpunode=None, nnmodel_in=ChtoModelv2, params=a["params"], method="emcee")
    np.testing.assert_almost_equal(np.mean(chain), 0.15151080063411168, decimal=5)
    np.testing.assert_almost_equal(np.std(chain), 0.9633211647095377, decimal=5)
    assert chain.shape == (56, 2) and np.asarray(logprob).shape == (800,)
    assert {n: digest(n) for n in before} == before, "nothing may be retrained or resampled in a finished reference directory"
    assert not any(f.startswith("chemcee_256.") and f != "chemcee_256.h5" for f in os.listdir(out / "iter_0"))


def test_posterior_moments_match_the_reference_chain(tmp_path):
    """Posterior moments against the chain the reference sampled with ITS emulator path (emcee + per-walker
    Log_prob.__call__ on the CPU, tests/golden/ref_chain): the same emulator files, likelihood and priors, sampled here
    by the GPU ensemble sampler over the fused kernel.  The shipped chain is short (200 iterations of 4 walkers,
    tau ~ 14: about 50 independent samples after burn-in), so the bar is its own standard error -- 4 sigma on the mean,
    4 sigma on the standard deviation -- and the stored log-probabilities of its walkers must be reproduced by the kernel
    to the likelihood tolerance."""
    import pickle
    import shutil
    import linna.util as U
    from linna_b200.sampler import ChainStore, EnsembleSampler, integrated_time
    from tests.helpers import GOLDEN, lnp_tol
    d = str(tmp_path / "iter_0")
    shutil.copytree(os.path.join(GOLDEN, "ref_fixture_iter_0"), d)
    ref = ChainStore(os.path.join(GOLDEN, "ref_chain", "chemcee_256.h5"))
    pred, yinv = U.retrieve_model(d, 2, 2)
    with open(os.path.join(d, "model_args.pkl"), "rb") as f:
        args = pickle.load(f)
    priors = [dict(param="test_%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)]
    tr = U.Transform(priors)
    lp = U.Log_prob(np.asarray(args[6]), np.asarray(args[2]), pred, yinv, tr, 1.0, U.gaussianlogliklihood, nograd=True)
    # (1) every log-probability the reference stored for its own walkers, re-evaluated by the kernel on the stored positions
    u_ref = np.asarray(ref.chain).reshape(-1, 2)
    got = lp(u_ref.astype(np.float32)).numpy().astype(np.float64)
    want = np.asarray(ref.log_prob).reshape(-1)
    assert np.all(np.abs(got - want) <= np.maximum(lnp_tol(want), 2e-5)), np.abs(got - want).max()
    # (2) moments of a long GPU chain against the reference chain after its burn-in
    tau = integrated_time(ref.chain)
    burn = int(3 * tau.max())
    tail = np.asarray(ref.chain_transformed)[burn:].reshape(-1, 2)
    n_eff = tail.shape[0] / tau.max()                       # walkers of one ensemble are correlated: a generous count
    torch.manual_seed(0)
    np.random.seed(0)
    samp = EnsembleSampler(512, 2, lp, seed=5)
    x0 = 0.5 * np.random.standard_normal((512, 2))
    samp.run_mcmc(x0, 400)
    u = samp.get_chain()[150:].reshape(-1, 2)
    th = np.asarray(tr(u.astype(np.float64)))
    mean, std = th.mean(0), th.std(0)
    se_mean = tail.std(0) / np.sqrt(n_eff)
    se_std = tail.std(0) / np.sqrt(2 * n_eff)
    assert np.all(np.abs(tail.mean(0) - mean) < 4 * se_mean), (tail.mean(0), mean, se_mean)
    assert np.all(np.abs(tail.std(0) - std) < 4 * se_std), (tail.std(0), std, se_std)
    assert np.all(np.abs(th) <= 2.0)
