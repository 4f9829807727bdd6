"""Samplers that drive the fused likelihood: an affine-invariant ensemble sampler whose whole state
(walker positions, lnP, RNG, accept/reject) lives on the GPU, the reference's convergence logic, and
chain storage with the reference's dataset names.

Reference: ``linna/sampler.py`` -- ``HMCSampler.sample`` (:456-554: emcee ensemble with a burn-in
selection, autocorrelation / mean-std convergence test every 100 iterations), ``ZeusSampler``
(:699-737), ``checkmeanstd`` (:370-387), the HDF5 back-ends (:322-368, :556-630).  emcee, zeus and
h5py are third-party and not vendored in the reference; this module re-implements the pieces the
path needs:

  * stretch move of Goodman & Weare (what ``emcee.EnsembleSampler`` runs by default): ONE batched
    ``Log_prob`` launch per half-ensemble instead of one pool task per walker
    (``linna/sampler.py:495``, no ``vectorize``);
  * integrated autocorrelation time with Sokal's automatic window (emcee's ``integrated_time``);
  * chains are written as ``<name>.npz`` with the arrays ``chain`` / ``chain_transformed`` /
    ``log_prob`` (the reference's HDF5 dataset names); when ``h5py`` is importable an
    ``<name>.h5`` with the same datasets under group ``mcmc`` is written as well.
"""
import os

import numpy as np
import torch

from . import engine as _engine


# ---------------------------------------------------------------------------------------- statistics
def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def _autocorr_1d(x):
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0] if acf[0] != 0 else acf


def _mean_acf_numpy(chain):
    """Host twin of _mean_acf_torch: one batched FFT over every (walker, parameter) series."""
    nstep = chain.shape[0]
    n = _next_pow_two(nstep)
    x = chain - chain.mean(axis=0, keepdims=True)
    f = np.fft.rfft(x, n=2 * n, axis=0)
    acf = np.fft.irfft(f * np.conjugate(f), n=2 * n, axis=0)[:nstep]
    a0 = acf[0:1]
    acf = np.where(a0 != 0, acf / np.where(a0 != 0, a0, 1.0), acf)
    return acf.mean(axis=1)


def _mean_acf_torch(chain):
    """Walker-averaged normalised autocorrelation function per parameter, batched FFT on the GPU:
    chain [steps, walkers, ndim] (device tensor) -> [steps, ndim] float64 on the host."""
    nstep = chain.shape[0]
    n = _next_pow_two(nstep)
    x = chain.to(torch.float64)
    x = x - x.mean(dim=0, keepdim=True)
    f = torch.fft.rfft(x, n=2 * n, dim=0)
    acf = torch.fft.irfft(f * f.conj(), n=2 * n, dim=0)[:nstep]
    a0 = acf[0:1]
    acf = torch.where(a0 != 0, acf / torch.where(a0 != 0, a0, torch.ones_like(a0)), acf)
    return acf.mean(dim=1).cpu().numpy()


def integrated_time(chain, c=5.0):
    """Integrated autocorrelation time per parameter of a chain [steps, walkers, ndim]
    (autocorrelation averaged over walkers, Sokal window M >= c*tau; always returns an estimate).
    Device tensors -- and large host chains when a GPU is present -- go through one batched FFT on the GPU
    (the reference recomputes this every 100 iterations over the whole chain, linna/sampler.py:532-550)."""
    on_gpu = torch.is_tensor(chain) and chain.is_cuda
    if not on_gpu:
        chain = np.asarray(chain, np.float64)
        if chain.ndim == 2:
            chain = chain[:, :, None]
        if chain.size >= (1 << 20) and torch.cuda.is_available():
            chain, on_gpu = torch.from_numpy(chain).cuda(), True
    elif chain.dim() == 2:
        chain = chain[:, :, None]
    nstep, nwalk, ndim = chain.shape
    if on_gpu:
        macf = _mean_acf_torch(chain)
    else:
        macf = _mean_acf_numpy(chain)
    tau = np.empty(ndim)
    for d in range(ndim):
        taus = 2.0 * np.cumsum(macf[:, d]) - 1.0
        m = np.arange(len(taus)) < c * taus
        window = np.argmin(m) if np.any(~m) else len(taus) - 1
        tau[d] = taus[window]
    return tau


def _halves_shift_torch(x):
    """(median |mean_a - mean_b| / std_b, median (std_a - std_b) / std_b) of the two halves of x [steps, walkers, ndim],
    reductions in float64 on the tensor's device; two scalars come back."""
    half = x.shape[0] // 2
    a = x[:half].reshape(-1, x.shape[-1]).to(torch.float64)
    b = x[half:].reshape(-1, x.shape[-1]).to(torch.float64)
    sa, sb = a.std(dim=0, unbiased=False), b.std(dim=0, unbiased=False)
    dm = _median_np(torch.abs(a.mean(dim=0) - b.mean(dim=0)) / sb)
    ds = _median_np((sa - sb) / sb)
    return float(dm), float(ds)


def _median_np(v):
    """numpy's median (mean of the two middle values for an even count) of a 1-D tensor."""
    s, _ = torch.sort(v)
    n = s.numel()
    return (s[(n - 1) // 2] + s[n // 2]) / 2


def checkmeanstd(samples, meanshift, stdshift):
    """Median shift of the mean (in sigma) and of the std (fractional) between the two halves of
    ``samples`` [steps, walkers, ndim] (linna/sampler.py:370-387).  Device tensors -- and large host chains when
    a GPU is present -- are reduced on the GPU (10^5 walkers x 100 steps x 30 parameters is 240 MB: the reference
    re-reads it on the host at every convergence check, linna/sampler.py:540-548)."""
    on_gpu = torch.is_tensor(samples) and samples.is_cuda
    if not on_gpu and not torch.is_tensor(samples) and np.size(samples) >= (1 << 22) and torch.cuda.is_available():
        samples, on_gpu = torch.from_numpy(np.ascontiguousarray(samples)).cuda(), True
    if torch.is_tensor(samples):
        dm, ds = _halves_shift_torch(samples)
        print(dm, ds, flush=True)
        return bool((dm < meanshift) & (ds < stdshift))
    half = int(len(samples) / 2)
    a = samples[:half].reshape(-1, samples.shape[-1])
    b = samples[half:].reshape(-1, samples.shape[-1])
    sb = np.std(b, axis=0)
    dm = np.median(np.abs(np.mean(a, axis=0) - np.mean(b, axis=0)) / sb)
    ds = np.median((np.std(a, axis=0) - sb) / sb)
    print(dm, ds, flush=True)
    return bool((dm < meanshift) & (ds < stdshift))


# ---------------------------------------------------------------------------------------- storage
class ChainStore:
    """Chain file with the reference's dataset names (linna/sampler.py:330-339, :572-578)."""

    def __init__(self, filename, transform=None):
        self.base = filename[:-3] if filename.endswith(".h5") else filename[:-4] if filename.endswith(".npz") else filename
        self.transform = transform
        self.chain = None           # [steps, walkers, ndim] latent positions
        self.chain_transformed = None
        self.log_prob = None
        if os.path.isfile(self.base + ".npz"):
            z = np.load(self.base + ".npz")
            self.chain, self.chain_transformed, self.log_prob = z["chain"], z["chain_transformed"], z["log_prob"]

    @property
    def iteration(self):
        return 0 if self.chain is None else len(self.chain)

    def exists(self):
        return self.chain is not None

    def extend(self, coords, log_prob):
        coords = np.asarray(coords, np.float64)
        tr = coords if self.transform is None else np.stack(
            [np.atleast_2d(self.transform(c.astype(np.float32))) for c in coords]).astype(np.float64)
        if self.chain is None:
            self.chain, self.chain_transformed, self.log_prob = coords, tr, np.asarray(log_prob, np.float64)
        else:
            self.chain = np.concatenate([self.chain, coords])
            self.chain_transformed = np.concatenate([self.chain_transformed, tr])
            self.log_prob = np.concatenate([self.log_prob, np.asarray(log_prob, np.float64)])

    def save(self):
        np.savez(self.base + ".npz", chain=self.chain, chain_transformed=self.chain_transformed, log_prob=self.log_prob)
        try:
            import h5py
            with h5py.File(self.base + ".h5", "w") as f:
                g = f.create_group("mcmc")
                g.create_dataset("chain", data=self.chain)
                g.create_dataset("chain_transformed", data=self.chain_transformed)
                g.create_dataset("log_prob", data=self.log_prob)
        except ImportError:
            pass

    def get_last_sample(self):
        return self.chain[-1]

    def get_value(self, name, discard=0, flat=False, thin=1):
        v = {"chain": self.chain, "chain_transformed": self.chain_transformed, "samples": self.chain_transformed,
             "log_prob": self.log_prob}[name][discard::thin]
        return v.reshape((-1,) + v.shape[2:]) if flat else v

    def get_log_prob(self, discard=0, flat=False, thin=1):
        return self.get_value("log_prob", discard, flat, thin)

    def get_autocorr_time(self, quiet=True, **kw):
        return integrated_time(self.chain)


def read_chain_and_cut(chainname, nk, ntimes=20, walkercut=False, method="emcee", flat=False):
    """Last ``nk`` autocorrelation times of a stored chain, flattened over walkers
    (linna/util.py:68-94)."""
    reader = ChainStore(chainname)
    if not reader.exists():
        raise FileNotFoundError(chainname)
    if nk > ntimes:
        print("Error: keep number greater then chain samples. nk: {0}, ntimes: {1}. This will lead to inclusion of "
              "all burn in step".format(nk, ntimes))
    nkeep = max(int(np.nanmedian(reader.get_autocorr_time()) * nk), 1)
    chain = reader.get_value("chain_transformed")
    lp = reader.get_log_prob()
    chain = chain[-nkeep:].reshape(-1, chain.shape[-1])
    lp = lp[-nkeep:]
    if flat:
        lp = lp.reshape(-1, 1)
    return chain, lp, reader


# ---------------------------------------------------------------------------------------- ensemble sampler
class EnsembleSampler:
    """Goodman-Weare stretch move with every array on the GPU.  ``log_prob`` is a ``linna.util.Log_prob``
    (fused: one kernel launch per half-ensemble) or any callable mapping a CUDA tensor [n, ndim] to [n]."""

    def __init__(self, nwalkers, ndim, log_prob, a=2.0, seed=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("EnsembleSampler: no CUDA device -- linna_b200 has no CPU fallback")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.gen = torch.Generator(device=self.device)
        self.seed = int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
        self.gen.manual_seed(self.seed)
        self._calls = 0               # Philox offsets 8c (proposal: 4 numbers) and 8c + 4 (acceptance: 1 number)
        self._lp = log_prob
        self.reset()

    def reset(self):
        self.iteration = 0
        self._chain, self._lnp = [], []
        self.naccepted = torch.zeros(self.nwalkers, device=self.device)

    def _eval(self, x):
        lp = self._lp
        if getattr(lp, "fused", False) and getattr(lp, "externalloglike", None) is None:
            out = lp.engine().lnp(x.contiguous())
        else:
            out = lp(x)
            out = torch.as_tensor(out, dtype=torch.float32, device=self.device)
        return torch.where(torch.isnan(out), torch.full_like(out, -torch.inf), out)

    @torch.no_grad()
    def sample(self, x0, iterations, store=True, lnp0=None):
        """Generator over iterations (like ``emcee.EnsembleSampler.sample``); yields (x, lnp) device tensors."""
        x = torch.as_tensor(np.asarray(x0), dtype=torch.float32).to(self.device).clone() if not torch.is_tensor(x0) \
            else x0.to(self.device, torch.float32).clone()
        x = x.contiguous()
        lnp = (self._eval(x) if lnp0 is None else lnp0.clone()).to(torch.float32).contiguous()
        W, d, a = self.nwalkers, self.ndim, self.a
        half = W // 2
        for _ in range(int(iterations)):
            perm = torch.randperm(W, device=self.device, generator=self.gen)
            for first, second in ((perm[:half], perm[half:]), (perm[half:], perm[:half])):
                # proposal and accept/reject are one CUDA kernel each (csrc/sampler_kernels.cu)
                first, second = first.contiguous(), second.contiguous()
                y, z = _engine.stretch_propose(x, first, second, a, self.seed, 8 * self._calls)
                lnp_y = self._eval(y).contiguous()
                _engine.stretch_accept(x, lnp, self.naccepted, first, y, lnp_y, z, self.seed, 8 * self._calls + 4)
                self._calls += 1
            self.iteration += 1
            if store:
                self._chain.append(x.clone())
                self._lnp.append(lnp.clone())
            yield x, lnp

    def run_mcmc(self, x0, nsteps, store=True):
        x = lnp = None
        for x, lnp in self.sample(x0, nsteps, store=store):
            pass
        return x, lnp

    def get_chain(self, flat=False, start=0):
        """Stored positions [steps, walkers, ndim] on the host, from iteration ``start`` on."""
        part = self._chain[start:]
        c = (torch.stack(part).cpu().numpy().astype(np.float64) if part else np.zeros((0, self.nwalkers, self.ndim)))
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False, start=0):
        part = self._lnp[start:]
        l = torch.stack(part).cpu().numpy().astype(np.float64) if part else np.zeros((0, self.nwalkers))
        return l.reshape(-1) if flat else l

    def get_autocorr_time(self, tol=0, **kw):
        return integrated_time(self.get_chain())

    @property
    def acceptance_fraction(self):
        return (self.naccepted / max(self.iteration, 1)).cpu().numpy()


class HMCSampler:
    """The reference's sampler wrapper (linna/sampler.py:389-554), ``method="emcee"`` branch: burn-in
    with re-selection of the best positions, then sampling until  tau*ntimes < n,  |d tau|/tau < tautol
    and the half-chain mean/std test pass (checked every 100 iterations)."""
    FILENAME = "chemcee_256.h5"

    def __init__(self, lnp, dlnp, ddlnp, ndim, nwalkers, x0=None, m=None, transform=None, torchspeed=False):
        self.lnp, self.dlnp, self.ddlnp = lnp, dlnp, ddlnp
        self.nparams, self.nwalkers = ndim, nwalkers
        self.x0, self.m, self.transform = x0, m, transform
        self.torchspeed = torchspeed
        self.sampler = None

    def sample(self, pool, nsamp, samp_steps=0, samp_eps=0, Madapt=1000, outdir="./", progress=False, overwrite=False,
               ntimes=10, tautol=0.01, method="emcee", incremental=True, meanshift=0.1, stdshift=0.1, nk=2,
               check_every=100, burnin=100):
        if method not in ("emcee", "zeus"):
            raise NotImplementedError("method %r: only the ensemble samplers are implemented here; for HMC use "
                                      "linna.HMCSampler.HMCSampler.sample_chains" % (method,))
        filename = os.path.join(outdir, self.FILENAME)
        store = ChainStore(filename, self.transform)
        x0 = self.x0
        resume = False
        if store.exists():
            if overwrite:
                store = ChainStore.__new__(ChainStore)
                store.base, store.transform = filename[:-3], self.transform
                store.chain = store.chain_transformed = store.log_prob = None
            else:
                print("init from previous")
                x0, resume = store.get_last_sample(), True
        self.sampler = EnsembleSampler(self.nwalkers, self.nparams, self.lnp)
        print("start", flush=True)
        if not incremental:
            self.sampler.run_mcmc(x0, nsamp)
            chain = self.sampler.get_chain(flat=True)
            return chain if self.transform is None else np.array([self.transform(c.astype(np.float32)) for c in chain])
        if not resume:
            print("burnin...", flush=True)
            burn = EnsembleSampler(self.nwalkers, self.nparams, self.lnp)
            burn.run_mcmc(x0, burnin)
            flat, lp = burn.get_chain(flat=True), burn.get_log_prob(flat=True)
            pos = flat[np.argsort(lp)[::-1][:int(50 * self.nwalkers)]]
            x0 = pos[np.random.randint(0, len(pos), self.nwalkers), :]
            print("burnin done...", flush=True)
        old_tau = np.inf
        saved = 0
        for _ in self.sampler.sample(x0, int(nsamp)):
            it = self.sampler.iteration
            if it % check_every:
                continue
            store.extend(self.sampler.get_chain(start=saved), self.sampler.get_log_prob(start=saved))   # only the new steps
            saved = it
            store.save()
            tau = integrated_time(store.chain)
            if np.isnan(np.sum(tau)) and it > 10:
                break
            converged = np.all(tau * ntimes < store.iteration)
            converged &= np.all(np.abs(old_tau - tau) / tau < tautol)
            keep = max(int(nk * np.mean(tau)), 2)
            converged &= checkmeanstd(store.chain[-keep:], meanshift=meanshift, stdshift=stdshift)
            print("max, min tau diff, max tau, ninter: {0}, {1}, {2}, {3}\n".format(
                np.max(np.abs(old_tau - tau) / tau), np.min(np.abs(old_tau - tau) / tau), np.max(tau), store.iteration),
                flush=True)
            if converged:
                break
            old_tau = tau
        if self.sampler.iteration > saved:
            store.extend(self.sampler.get_chain(start=saved), self.sampler.get_log_prob(start=saved))
            store.save()
        self.sampler = None
        return store


class ZeusSampler(HMCSampler):
    """Interface of the reference's zeus wrapper (linna/sampler.py:699-737).  zeus' ensemble slice move is not
    re-implemented: the same target is sampled with the stretch move, stored under the zeus file name."""
    FILENAME = "zeus_256.h5"

    def __init__(self, lnp, ndim, nwalkers, x0=None, transform=None):
        super().__init__(lnp, None, None, ndim, nwalkers, x0=x0, m=None, transform=transform)

    def sample(self, pool, nsamp, outdir="./", progress=False, overwrite=False, ntimes=10, tautol=0.01,
               incremental=True, meanshift=0.1, stdshift=0.1, nk=2, **kw):
        return super().sample(pool, nsamp, outdir=outdir, progress=progress, overwrite=overwrite, ntimes=ntimes,
                              tautol=tautol, method="zeus", incremental=incremental, meanshift=meanshift,
                              stdshift=stdshift, nk=nk, **kw)


Zeusbackend = ChainStore

for _c in (ChainStore, EnsembleSampler, HMCSampler, ZeusSampler):
    _c.__module__ = "linna.sampler"
del _c
