"""``ml_sampler`` / ``ml_sampler_core``: the iterative train-then-sample loop of LINNA (linna/main.py:22-334)
around the fused GPU likelihood.

Per iteration (temperature schedule T = temperatureArr[i]^2, main.py:153): draw training parameters (prior box
first, previous chain afterwards), evaluate ``theory(x, outdirs)`` on them, train the emulator on the GPU with
the fused kernels, then sample lnP(u) with the on-GPU ensemble sampler until the autocorrelation and mean/std
tests pass.  Files under ``outdir/iter_<k>/`` follow the reference layout; chains are written as
``<name>.npz`` (+ ``.h5`` when h5py is importable) with the reference's dataset names."""
import gc
import os
import pickle
import tempfile
from copy import deepcopy

import numpy as np
import torch

from .nn import ChtoModelv2
from .orchestrate import generate_training_point, run_mcmc
from .sampler import read_chain_and_cut
from .util import (LogPrior, Log_prob, NN_samplerv1, Transform, gaussianlogliklihood, invTransform, logp_theory_data,
                   retrieve_model, train_NN)
from . import train_gpu


def ml_sampler(outdir, theory, priors, data, cov, init, pool, nwalkers, gpunode, omegab2cut=None, nepoch=4500, method="zeus",
               nbest=None, chisqcut=None, loglikelihoodfunc=None):
    """LINNA with the hyper-parameters of To et al. 2022 (linna/main.py:22-75)."""
    ntrainArr = [10000, 10000, 10000, 10000]
    nvalArr = [500, 500, 500, 500]
    if method == "emcee":
        nkeepArr, ntimesArr = [2, 2, 5, 4], [5, 5, 10, 15]
    elif method == "zeus":
        nkeepArr, ntimesArr = [2, 2, 5, 5], [5, 5, 10, 50]
    else:
        raise NotImplementedError(method)
    ntautolArr = [0.03, 0.03, 0.02, 0.01]
    temperatureArr = [4.0, 2.0, 1.0, 1.0]
    meanshiftArr = [0.2, 0.2, 0.2, 0.2]
    stdshiftArr = [0.15, 0.15, 0.15, 0.15]
    params = {"trainingoption": 1, "num_epochs": nepoch, "batch_size": 500}
    return ml_sampler_core(ntrainArr, nvalArr, nkeepArr, ntimesArr, ntautolArr, meanshiftArr, stdshiftArr, outdir, theory,
                           priors, data, cov, init, pool, nwalkers, "cuda", None, False, temperatureArr, omegab2cut, False, 1,
                           gpunode, ChtoModelv2, params, method, nbest=nbest, chisqcut=chisqcut,
                           loglikelihoodfunc=loglikelihoodfunc)


def _chain_file(outdir_iter, filename):
    base = os.path.join(outdir_iter, filename[:-3])
    for ext in (".npz", ".h5"):
        if os.path.isfile(base + ext):
            return base + ext
    return None


def ml_sampler_core(ntrainArr, nvalArr, nkeepArr, ntimesArr, ntautolArr, meanshiftArr, stdshiftArr, outdir, theory, priors, data,
                    cov, init, pool, nwalkers, device, dolog10index, ypositive, temperatureArr, omegab2cut=None, docuda=False,
                    tsize=1, gpunode=None, nnmodel_in=None, params=None, method="emcee", nbest=None, chisqcut=None,
                    loglikelihoodfunc=None, nsigma=3, externalloglike=None):
    """linna/main.py:77-334.  ``device``, ``docuda``, ``tsize`` and ``gpunode`` are accepted for signature
    compatibility: training and sampling always run on the local GPU."""
    params = dict(params or {})
    nnmodel_in = ChtoModelv2 if nnmodel_in is None else nnmodel_in
    data, cov = np.asarray(data, np.float64), np.asarray(cov, np.float64)
    ndim = len(init)
    sigma = np.sqrt(np.diag(cov))
    inv_cov = np.linalg.inv(cov)
    prior_range = []
    for item in priors:
        if item["dist"] == "flat":
            prior_range.append([item["arg1"], item["arg2"]])
        elif item["dist"] == "gauss":
            prior_range.append([item["arg1"] - 5 * item["arg2"], item["arg1"] + 5 * item["arg2"]])
        else:
            print("not implement dist : {0}".format(item["dist"]), flush=True)
            assert 0
    transform = Transform(priors)
    init = np.asarray(invTransform(priors)(init))
    if method == "emcee":
        filename = "chemcee_256.h5"
    elif method == "zeus":
        filename = "zeus_256.h5"
    else:
        raise NotImplementedError(method)
    nk = ntimes = None
    schedule = zip(ntrainArr, nvalArr, nkeepArr, ntimesArr, ntautolArr, temperatureArr, meanshiftArr, stdshiftArr)
    for i, (nt, nv, nk, ntimes, tautol, temperature, meanshift, stdshift) in enumerate(schedule):
        nbest_in = nbest[i] if isinstance(nbest, list) else nbest
        if nbest_in is not None and nbest_in <= 0:
            nbest_in = None
        negloglike = None
        if nbest_in is not None:
            tempdir = tempfile.TemporaryDirectory()

            def negloglike(x, _t=tempdir):
                d = data - theory([-1, x], _t.name)
                return d.dot(inv_cov.dot(d))
        temperature = temperature ** 2
        print("#" * 100)
        print("iteration: {0}".format(i), flush=True)
        print("#" * 100)
        outdir_in = os.path.join(outdir, "iter_{0}/".format(i))
        chain = None
        if i > 0:
            prev = _chain_file(os.path.join(outdir, "iter_{0}/".format(i - 1)), filename)
            if prev is not None:
                chain, _, _ = read_chain_and_cut(prev, nk, ntimes, method=method)
            else:
                chain = np.loadtxt(os.path.join(outdir, "iter_{0}/".format(i - 1), filename[:-3] + ".txt"))[-100000:, :-1]
        nnsampler = NN_samplerv1(outdir_in, prior_range)
        generate_training_point(theory, nnsampler, pool, outdir_in, nt, nv, data, inv_cov, chain, nsigma=nsigma,
                                omegab2cut=omegab2cut, options=params.get("trainingoption", 0), negloglike=negloglike,
                                nbest_in=nbest_in, chisqcut=chisqcut)
        del chain
        gc.collect()
        if pool is None or pool.is_master():
            outdir_list = [os.path.join(outdir, "iter_{0}/".format(m)) for m in range(i + 1)]
            with open(os.path.join(outdir_in, "model_pickle.pkl"), "wb") as f:
                pickle.dump(train_NN, f)
            with open(os.path.join(outdir_in, "model_args.pkl"), "wb") as f:
                pickle.dump([nnsampler, cov, inv_cov, sigma, outdir_in, outdir_list, data, dolog10index, ypositive, False, 2,
                             temperature, True, None, 1, nnmodel_in, params, nbest_in is not None], f)
            if not os.path.isfile(os.path.join(outdir_in, "finish.pkl")):
                train_gpu.main(outdir_in)
        model, y_invtransform_data = retrieve_model(outdir_in, len(init), len(data), nnmodel_in)
        cf = _chain_file(outdir_in, filename)
        if cf is not None and (os.path.isfile(os.path.join(outdir_in, "mcmc_done.pkl")) or
                               (cf.endswith(".h5") and not os.path.isfile(cf[:-3] + ".meta.json"))):
            # a finished chain of this package, or a chain file written by the reference itself (an HDF5 file with none of
            # this package's bookkeeping beside it): linna/main.py:273-274 skips the MCMC whenever the chain file exists
            continue
        invcov_new = torch.from_numpy(inv_cov.astype(np.float32))
        data_new = torch.from_numpy(data.astype(np.float32))
        log_prob = Log_prob(data_new, invcov_new, model, y_invtransform_data, transform, temperature, nograd=True,
                            loglikelihoodfunc=gaussianlogliklihood if loglikelihoodfunc is None else loglikelihoodfunc,
                            externalloglike=externalloglike)
        if pool is not None:
            pool.noduplicate = True
        run_mcmc(nnsampler, outdir_in, method, ndim, nwalkers, init, log_prob, dlnp=None, ddlnp=None, pool=pool,
                 transform=transform, ntimes=ntimes, tautol=tautol, meanshift=meanshift, stdshift=stdshift, nk=nk)
        with open(os.path.join(outdir_in, "mcmc_done.pkl"), "wb") as f:
            pickle.dump([True], f)
        if pool is not None and hasattr(pool, "noduplicate_close"):
            pool.noduplicate_close()
    last = os.path.join(outdir, "iter_{0}/".format(len(ntrainArr) - 1))
    chain_name = _chain_file(last, filename)
    if chain_name is not None:
        chain, log_prob_samples_x, reader = read_chain_and_cut(chain_name, nk, ntimes, method=method)
        log_prob_samples_x = reader.get_log_prob(discard=0, flat=True, thin=1)
    else:
        txt = np.loadtxt(os.path.join(last, filename[:-3] + ".txt"))
        chain, log_prob_samples_x = txt[-100000:, :-1], txt[-100000:, -1]

    # optional importance sampling against the true theory (main.py:301-333)
    if "nimp" in params:
        if not os.path.isfile(os.path.join(outdir, "samples_im.npy")):
            chain, log_prob_samples_x, _ = read_chain_and_cut(chain_name, nk, ntimes, method=method, flat=True)
            select = np.random.randint(0, len(chain), params["nimp"])
            chain, log_prob_samples_x = chain[select], np.asarray(log_prob_samples_x).reshape(-1)[select]
            np.save(os.path.join(outdir, "samples_im.npy"), chain)
            np.save(os.path.join(outdir, "log_prob_samples_x.npy"), log_prob_samples_x)
        else:
            chain = np.load(os.path.join(outdir, "samples_im.npy"))
            log_prob_samples_x = np.load(os.path.join(outdir, "log_prob_samples_x.npy"))
        outimp = os.path.join(outdir, "imp/")
        os.makedirs(outimp, exist_ok=True)
        nnsampler = NN_samplerv1(outimp, prior_range)
        if not os.path.isfile(os.path.join(outdir, "theory.npy")):
            th = nnsampler.generate_training_data(zip(range(len(chain)), chain), theory, pool=pool, args=[outimp])
            np.save(os.path.join(outdir, "theory.npy"), th)
        else:
            th = np.load(os.path.join(outdir, "theory.npy"))
        log_prob_samples_x = np.asarray(log_prob_samples_x).flatten()
        logp = np.asarray(logp_theory_data(chain, th, data, inv_cov, LogPrior(priors)), np.float64)
        w = np.exp(logp - log_prob_samples_x)
        lw = np.log(np.maximum(w, 1e-300))
        w[np.abs(lw - np.mean(lw)) > 2 * np.std(lw)] = 0
        w = w / np.sum(w)
        np.save(os.path.join(outdir, "weight_im.npy"), [log_prob_samples_x.flatten(), logp, w])
    return chain, log_prob_samples_x


for _f in (ml_sampler, ml_sampler_core):
    _f.__module__ = "linna.main"
del _f
