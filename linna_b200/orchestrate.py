"""The callers on either side of the hot path: training-set generation, the per-iteration training hand-off
and the MCMC launch of ``ml_sampler``.

Reference: ``NN_samplerv1`` methods (linna/util.py:749-951), ``generate_training_point``
(:1166-1254), ``chisqcut_all`` (:1256-1270), ``run_mcmc`` (:1472-1504).  What changes:

  * the emulator is trained in-process on the GPU (``train_gpu.main``) instead of through ``srun`` /
    ``os.system`` + ``finish.pkl`` polling (linna/main.py:199-257); the files written are the same;
  * the MCMC runs on the on-GPU ensemble sampler of ``linna_b200.sampler`` (one fused likelihood launch per
    half-ensemble), not on emcee / zeus;
  * two third-party dependencies of the reference are absent here and un-vendored there -- ``pyDOE2``
    (``lhs(..., criterion="center")``) and ``sample_generator`` (latin hypercube in the principal-axis frame of
    a chain).  Their published algorithms are restated below; parity with them is unpinned.
"""
import glob
import os
import tempfile
from copy import deepcopy

import numpy as np


# ------------------------------------------------------------------------------------- latin hypercubes
def lhs_center(ndim, nsamples, seed):
    """Centred latin hypercube on [0, 1]^ndim: every column is a random permutation of the bin centres
    (pyDOE2 ``lhs(n, samples, criterion="center")``)."""
    rng = np.random.RandomState(seed)
    centres = (np.arange(nsamples) + 0.5) / nsamples
    out = np.empty((nsamples, ndim))
    for j in range(ndim):
        out[:, j] = rng.permutation(centres)
    return out


def lhs_around_chain(chain, nsamples, scale, seed):
    """Latin hypercube of half-width ``scale`` standard deviations in the principal-axis frame of ``chain``
    (what ``sample_generator.SampleGenerator(chain, scale).get_samples(n, "LH")`` produces)."""
    chain = np.asarray(chain, np.float64)
    mean = chain.mean(axis=0)
    cov = np.atleast_2d(np.cov(chain, rowvar=False))
    w, v = np.linalg.eigh(cov)
    w = np.clip(w, 0.0, None)
    cube = 2.0 * lhs_center(chain.shape[1], nsamples, seed) - 1.0
    return mean + (cube * scale * np.sqrt(w)) @ v.T


# ------------------------------------------------------------------------------------- NN_samplerv1 methods
def _omegab2_keep(samples, omegab2cut):
    """linna/util.py:805-812."""
    ombh2 = samples[:, omegab2cut[0]] * samples[:, omegab2cut[1]] ** 2
    keep = (ombh2 > omegab2cut[2]) & (ombh2 < omegab2cut[3])
    if len(omegab2cut) > 4:
        keep &= (samples[:, omegab2cut[4]] > omegab2cut[5]) & (samples[:, omegab2cut[4]] < omegab2cut[6])
    if len(omegab2cut) > 7:
        keep &= (samples[:, omegab2cut[7]] > omegab2cut[8]) & (samples[:, omegab2cut[7]] < omegab2cut[9])
    return keep


def nnsampler_generate_training_data(self, samples, model, pool=None, args=None, kwargs=None):
    """theory(x, outdirs) over the rows of ``samples`` (an iterable of (index, params)), through ``pool.map``
    when a pool is given (linna/util.py:751-777).  The scratch directory args[0] is emptied before and after."""
    from .util import _FunctionWrapper
    m = _FunctionWrapper(model, args, kwargs)
    scratch = args[0] if args else None

    def clean():
        if scratch is not None and isinstance(scratch, str):
            for f in glob.glob(os.path.join(scratch + "/", "*")):
                os.remove(f)
    clean()
    out = np.array(list(pool.map(m, samples) if pool is not None else map(m, samples)))
    clean()
    return out


def nnsampler_gensample_flat(self, Nsamples, omegab2cut=None):
    """Latin hypercube over the prior box (linna/util.py:778-815); the second parameter is sampled in log when
    its upper limit is below 1e-5 (A_s)."""
    samples = []
    n_in = Nsamples
    while len(samples) < Nsamples:
        samples = 2.0 * (lhs_center(len(self.prior_range), int(n_in), self.seed) - 0.5)
        for ind, prior in enumerate(self.prior_range):
            prior = np.asarray(prior, np.float64)
            log_as = ind == 1 and self.prior_range[1][1] < 1e-5
            if log_as:
                prior = np.log(prior)
            samples[:, ind] = samples[:, ind] * (prior[1] - prior[0]) / 2 + (prior[1] + prior[0]) / 2
            if log_as:
                samples[:, ind] = np.exp(samples[:, ind])
        if omegab2cut is not None:
            samples = samples[_omegab2_keep(samples, omegab2cut)]
        n_in += 1000
    return samples[:Nsamples]


def nnsampler_gensample_chain(self, Nsamples, chain_in, nsigma, omegab2cut=None):
    """Latin hypercube in the ``nsigma`` region of a chain, cut to the prior box (linna/util.py:816-862)."""
    chain = deepcopy(np.asarray(chain_in, np.float64))
    prior_in = deepcopy([list(p) for p in self.prior_range])
    Nsamples = int(Nsamples)
    log_as = prior_in[1][1] < 1e-5 if len(prior_in) > 1 else False
    if log_as:
        chain[:, 1] = np.log(1e10 * chain[:, 1])
        prior_in[1] = [np.log(1e10 * prior_in[1][0]), np.log(1e10 * prior_in[1][1])]
    total, n_factor = 0, 1
    x = chain[:0]
    while total < Nsamples:
        x = lhs_around_chain(chain, int(n_factor * Nsamples), nsigma, self.seed)
        if omegab2cut is not None:
            x = x[_omegab2_keep(x, omegab2cut)]
        for i in range(x.shape[1]):
            x = x[(x[:, i] > prior_in[i][0]) & (x[:, i] < prior_in[i][1])]
        if log_as:
            x[:, 1] = np.exp(x[:, 1]) / 1e10
        n_factor += 1
        total = x.shape[0]
        if n_factor > 64:
            raise RuntimeError("gensample_chain: the chain region does not intersect the prior box")
    return x[:Nsamples]


def nnsampler_gensample_chain_randomsample(self, Nsamples, chain_in, nsigma, omegab2cut=None):
    """Random draws (with replacement) from the part of the chain inside the prior box (linna/util.py:865-897)."""
    chain = deepcopy(np.asarray(chain_in))
    if omegab2cut is not None:
        chain = chain[_omegab2_keep(chain, omegab2cut)]
    for i in range(chain.shape[1]):
        chain = chain[(chain[:, i] > self.prior_range[i][0]) & (chain[:, i] < self.prior_range[i][1])]
    np.random.seed(self.seed)
    return chain[np.random.randint(0, len(chain), int(Nsamples))]


def nnsampler_emcee_sample(self, log_prob, ndim, nwalkers, init, pool, transform, ntimes=50, tautol=0.01, dlnp=None,
                           ddlnp=None, meanshift=0.1, stdshift=0.1, nk=1):
    """linna/util.py:899-919 (ensemble sampler until the autocorrelation / mean-std tests pass)."""
    from . import sampler
    x0 = init + 0.1 * np.random.randn(nwalkers, ndim)
    samp = sampler.HMCSampler(log_prob, dlnp, ddlnp, ndim, nwalkers, x0=x0, m=None, transform=transform)
    return samp.sample(pool, 1000000, 0, 0, outdir=self.outdir, overwrite=False, ntimes=ntimes, method="emcee",
                       incremental=True, progress=False, tautol=tautol, meanshift=meanshift, stdshift=stdshift, nk=nk)


def nnsampler_Zeus_sample(self, log_prob, ndim, nwalkers, init, pool, transform, ntimes=50, tautol=0.01, dlnp=None,
                          ddlnp=None, meanshift=0.1, stdshift=0.1, nk=1):
    """linna/util.py:921-940."""
    from . import sampler
    x0 = init + 0.001 * np.random.randn(nwalkers, ndim)
    samp = sampler.ZeusSampler(log_prob, ndim, nwalkers, x0=x0, transform=transform)
    return samp.sample(pool, 1000000, outdir=self.outdir, overwrite=False, ntimes=ntimes, incremental=True,
                       progress=False, tautol=tautol, meanshift=meanshift, stdshift=stdshift, nk=nk)


def nnsampler__HMC_sample(self, log_prob, dlnp, ddlnp, ndim, nwalkers, init, pool, transform, samp_steps, samp_eps):
    """Intended behaviour of linna/util.py:940-945: ``nwalkers`` HMC chains started around ``init``, ``samp_steps`` leapfrog
    steps of size ``samp_eps`` per sample, mass matrix from the Hessian of lnP (linna/sampler.py:408-456), chain in
    ``chhmc.h5``'s datasets.  The chains run batched on the GPU (``sampler.HMCSampler.sample(method="hmc")``); the mass
    is the DIAGONAL of -Hessian at ``init`` (the reference rotates into the Hessian's eigenbasis after a Nelder-Mead +
    BFGS search for the MAP -- a host-side optimisation outside the likelihood path)."""
    from . import sampler
    x0 = init + 0.1 * np.random.randn(nwalkers, ndim)
    mass = None
    if ddlnp is not None:
        try:
            h = -np.diag(np.asarray(ddlnp(np.asarray(init, np.float64).reshape(-1)), np.float64))
            mass = np.where(np.isfinite(h) & (h > 1e-6), h, 1.0).astype(np.float32)
        except Exception:
            mass = None
    samp = sampler.HMCSampler(log_prob, dlnp, ddlnp, ndim, nwalkers, x0=x0, m=mass, transform=transform)
    return samp.sample(pool, 1000000, samp_steps, samp_eps, outdir=self.outdir, overwrite=True, ntimes=50, method="hmc",
                       incremental=True, progress=True)


# ------------------------------------------------------------------------------------- training points
def _hessian(f, x, rel=1e-4):
    """Central-difference Hessian (stands in for numdifftools.Hessian, linna/util.py:1237)."""
    x = np.asarray(x, np.float64)
    n = len(x)
    h = rel * np.maximum(np.abs(x), 1.0)
    H = np.empty((n, n))
    for i in range(n):
        for j in range(i, n):
            ei, ej = np.zeros(n), np.zeros(n)
            ei[i], ej[j] = h[i], h[j]
            H[i, j] = H[j, i] = (f(x + ei + ej) - f(x + ei - ej) - f(x - ei + ej) + f(x - ei - ej)) / (4 * h[i] * h[j])
    return H


def _make_positive_definite(H):
    w, v = np.linalg.eigh(0.5 * (H + H.T))
    return (v * np.maximum(w, 1e-10 * np.max(np.abs(w)))) @ v.T


def chisqcut_all(data, invcov, chisqcut, fnamey, fnamex):
    """Drop training rows whose chi^2 against the data exceeds ``chisqcut`` (linna/util.py:1256-1270)."""
    y, x = np.load(fnamey), np.loadtxt(fnamex)
    d = y[:, :len(data)] - data
    keep = np.einsum("ij,jk,ik->i", d, invcov, d) < chisqcut
    np.save(fnamey, y[keep])
    np.savetxt(fnamex, x[keep])


def generate_training_point(theory, nnsampler, pool, outdir, ntrain, nval, data, invcov, chain, nsigma=3, omegab2cut=None,
                            options=0, negloglike=None, nbest_in=None, chisqcut=None):
    """Write train/val parameter sets and their theory vectors under ``outdir`` (linna/util.py:1166-1254): from
    the prior box in the first iteration, from the previous chain afterwards."""
    if pool is not None and not pool.is_master():
        return
    os.makedirs(outdir, exist_ok=True)

    def draw(n):
        if chain is None:
            return nnsampler.gensample_flat(n, omegab2cut=omegab2cut)
        if options == 0:
            return nnsampler.gensample_chain(n, chain, nsigma, omegab2cut=omegab2cut)
        if options == 1:
            return nnsampler.gensample_chain_randomsample(n, chain, nsigma, omegab2cut=omegab2cut)
        print("options : {0} not recognized".format(options))
        assert 0
    for name, n in (("train", ntrain), ("val", nval)):
        fx = os.path.join(outdir, name + "_samples_x.txt")
        if not os.path.isfile(fx):
            np.savetxt(fx, draw(n))
        scratch = os.path.join(outdir, name + "/")
        os.makedirs(scratch, exist_ok=True)
        fy = os.path.join(outdir, name + "_samples_y.npy")
        if not os.path.isfile(fy):
            x = np.atleast_2d(np.loadtxt(fx))
            np.save(fy, nnsampler.generate_training_data(zip(range(len(x)), x), theory, pool=pool, args=[scratch]))
    if negloglike is not None:
        from scipy.optimize import minimize
        from scipy.stats import multivariate_normal
        fbx = os.path.join(outdir, "best_samples_x.txt")
        if not os.path.isfile(fbx):
            train_x = np.atleast_2d(np.loadtxt(os.path.join(outdir, "train_samples_x.txt")))
            best = minimize(negloglike, train_x[0], method="Nelder-Mead", tol=1e-6).x
            inv_hess = np.linalg.inv(_make_positive_definite(_hessian(negloglike, best)))
            np.savetxt(fbx, np.atleast_2d(multivariate_normal.rvs(mean=best, cov=inv_hess, size=nbest_in)))
            np.savetxt(os.path.join(outdir, "best_samples_x_val.txt"),
                       np.atleast_2d(multivariate_normal.rvs(mean=best, cov=inv_hess, size=max(int(nbest_in / ntrain * nval), 1))))
        if not os.path.isfile(os.path.join(outdir, "best_samples_y.npy")):
            for fx, fy in (("best_samples_x.txt", "best_samples_y.npy"), ("best_samples_x_val.txt", "best_samples_y_val.npy")):
                bx = np.atleast_2d(np.loadtxt(os.path.join(outdir, fx)))
                with tempfile.TemporaryDirectory() as tmp:
                    np.save(os.path.join(outdir, fy),
                            nnsampler.generate_training_data(zip(range(len(bx)), bx), theory, pool=pool, args=[tmp]))
    if chisqcut is not None:
        names = [("train_samples_y.npy", "train_samples_x.txt"), ("val_samples_y.npy", "val_samples_x.txt")]
        if negloglike is not None:
            names += [("best_samples_y.npy", "best_samples_x.txt"), ("best_samples_y_val.npy", "best_samples_x_val.txt")]
        for fy, fx in names:
            chisqcut_all(data, invcov, chisqcut, os.path.join(outdir, fy), os.path.join(outdir, fx))


def run_mcmc(nnsampler, outdir, method, ndim, nwalkers, init, log_prob, dlnp=None, ddlnp=None, pool=None, transform=None,
             ntimes=50, tautol=0.01, meanshift=0.1, stdshift=0.1, nk=2):
    """linna/util.py:1472-1504.  ``emcee`` and ``zeus`` run on the GPU ensemble sampler, ``hmc`` (and, as in the reference,
    any other name except "nuts") on the batched GPU HMC chains with the reference's defaults of 5 leapfrog steps of size
    0.1 (the reference's own hmc branch passes its arguments in the wrong order, SURVEY Q3: this is the intended call);
    "nuts" is a stub in the reference (linna/sampler.py:14-21) and raises here."""
    samp_steps, samp_eps = 5, 0.1
    if method == "nuts":
        raise NotImplementedError("nuts: stop_criterion / leapfrog / build_tree are NotImplementedError stubs in the reference "
                                  "(linna/sampler.py:14-21)")
    if method == "emcee":
        return nnsampler.emcee_sample(log_prob, ndim, nwalkers, init, pool, ntimes=ntimes, tautol=tautol, transform=transform,
                                      dlnp=dlnp, ddlnp=ddlnp, meanshift=meanshift, stdshift=stdshift, nk=nk)
    if method == "zeus":
        return nnsampler.Zeus_sample(log_prob, ndim, nwalkers, init, pool, ntimes=ntimes, tautol=tautol, transform=transform,
                                     dlnp=dlnp, ddlnp=ddlnp, meanshift=meanshift, stdshift=stdshift, nk=nk)
    return nnsampler._HMC_sample(log_prob, dlnp, ddlnp, ndim, nwalkers, init, pool, transform, samp_steps, samp_eps)
