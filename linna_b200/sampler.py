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
            chain, on_gpu = torch.from_numpy(np.ascontiguousarray(chain).copy() if not chain.flags.writeable else chain).cuda(), True
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


def _halves_shift_device(x):
    """(median |mean_a - mean_b| / std_b, median (std_a - std_b) / std_b) of the two halves of x [steps, walkers, ndim]
    on the GPU: per-parameter mean and std of each half by the two-pass float64 reduction kernel
    (``linna_column_moments``, csrc/sampler_kernels.cu); 4 x ndim numbers come back."""
    nd = x.shape[-1]
    half = x.shape[0] // 2
    rows_per_step = int(np.prod(x.shape[1:-1])) if x.dim() > 2 else 1
    flat = x.reshape(-1, nd)
    if flat.dtype not in (torch.float32, torch.float64):
        flat = flat.to(torch.float32)
    flat = flat.contiguous()
    ma, sa = _engine.column_moments(flat, 0, half * rows_per_step)
    mb, sb = _engine.column_moments(flat, half * rows_per_step, flat.shape[0])
    return float(np.median(np.abs(ma - mb) / sb)), float(np.median((sa - sb) / sb))


def checkmeanstd(samples, meanshift, stdshift):
    """Median shift of the mean (in sigma) and of the std (fractional) between the two halves of
    ``samples`` [steps, walkers, ndim] (linna/sampler.py:370-387).  Device tensors -- and large host chains when
    a GPU is present -- are reduced on the GPU (10^5 walkers x 100 steps x 30 parameters is 240 MB: the reference
    re-reads it on the host at every convergence check, linna/sampler.py:540-548)."""
    on_gpu = torch.is_tensor(samples) and samples.is_cuda
    if not on_gpu and not torch.is_tensor(samples) and np.size(samples) >= (1 << 22) and torch.cuda.is_available():
        samples, on_gpu = torch.from_numpy(np.ascontiguousarray(samples)).cuda(), True
    if on_gpu:
        dm, ds = _halves_shift_device(samples)
        print(dm, ds, flush=True)
        return bool((dm < meanshift) & (ds < stdshift))
    samples = samples.detach().cpu().numpy() if torch.is_tensor(samples) else np.asarray(samples)
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
    """Chain file with the reference's dataset names (linna/sampler.py:330-339, :572-578).

    The chain is APPENDED, never rewritten: every flush adds its block to three raw files
    (``<name>.chain.f32`` / ``.chain_transformed.f32`` / ``.log_prob.f32`` + ``<name>.meta.json``; float32 is what the GPU
    sampler produces -- runs started by an older build keep their ``.f64`` files) -- what the reference's
    HDF5 back-end does with resizable datasets -- and the arrays are read back as memory maps, so neither the host memory
    nor the I/O per convergence check grows with the length of the run.  ``finalize()`` writes the ``<name>.npz`` (and,
    when h5py is importable, ``<name>.h5`` with group ``mcmc``) once, at the end.  A finished ``.npz`` is what is read on
    resume; an unfinished run resumes from the raw files."""
    NAMES = ("chain", "chain_transformed", "log_prob")

    def __init__(self, filename, transform=None, fresh=False):
        self.base = filename[:-3] if filename.endswith(".h5") else filename[:-4] if filename.endswith(".npz") else filename
        self.transform = transform
        self.nsteps, self.nwalkers, self.ndim = 0, 0, 0
        self._frozen = None          # arrays of a finished .npz
        self._maps = {}
        self._raw = "f32"            # element type of the raw append files
        if fresh:
            self._remove_files()
        elif os.path.isfile(self.base + ".npz"):
            z = np.load(self.base + ".npz")
            self._frozen = {k: z[k] for k in self.NAMES}
            self.nsteps, self.nwalkers, self.ndim = self._frozen["chain"].shape
        elif os.path.isfile(self.base + ".h5") and not os.path.isfile(self.base + ".meta.json"):
            self._frozen = self._load_h5(self.base + ".h5")    # a chain written by the reference (emcee HDFBackend)
            self.nsteps, self.nwalkers, self.ndim = self._frozen["chain"].shape
        elif os.path.isfile(self.base + ".meta.json"):
            import json
            with open(self.base + ".meta.json") as f:
                meta = json.load(f)
            self.nsteps, self.nwalkers, self.ndim = int(meta["nsteps"]), int(meta["nwalkers"]), int(meta["ndim"])
            self._raw = meta.get("dtype", "f64")

    @staticmethod
    def _load_h5(path, group="mcmc"):
        """The datasets of an emcee-style HDF5 chain (linna/sampler.py:330-368), cut at the attribute ``iteration``
        (emcee grows its datasets ahead of the samples it has written)."""
        from .h5read import H5File
        f = H5File(path)
        names = f.keys("/" + group)
        it = f.attrs("/" + group).get("iteration")
        out = {}
        for n in ("chain", "chain_transformed", "log_prob"):
            src = n if n in names else ("chain" if n == "chain_transformed" else None)   # linna/util.py:81-84
            if src is None:
                raise KeyError("%s: no dataset %r in group %r" % (path, n, group))
            v = f.dataset("/%s/%s" % (group, src))
            out[n] = np.ascontiguousarray(v[:int(it)] if it is not None else v, np.float64)
        return out

    def _remove_files(self):
        for ext in (".npz", ".h5", ".meta.json") + tuple("." + n + e for n in self.NAMES for e in (".f64", ".f32")):
            if os.path.isfile(self.base + ext):
                os.remove(self.base + ext)

    def _path(self, name):
        return self.base + "." + name + "." + self._raw

    @property
    def _dtype(self):
        return np.float32 if self._raw == "f32" else np.float64

    def _shape(self, name):
        return (self.nsteps, self.nwalkers) if name == "log_prob" else (self.nsteps, self.nwalkers, self.ndim)

    def _array(self, name):
        if self._frozen is not None:
            return self._frozen[name]
        if self.nsteps == 0:
            return None
        mm = self._maps.get(name)
        if mm is None or mm.shape[0] != self.nsteps:
            mm = np.memmap(self._path(name), dtype=self._dtype, mode="r", shape=self._shape(name))
            self._maps[name] = mm
        return mm

    chain = property(lambda self: self._array("chain"))
    chain_transformed = property(lambda self: self._array("chain_transformed"))
    log_prob = property(lambda self: self._array("log_prob"))

    @property
    def iteration(self):
        return self.nsteps

    def exists(self):
        return self.nsteps > 0

    def extend(self, coords, log_prob, transformed=None):
        """Append a block [steps, walkers, ndim].  ``transformed``: the block already mapped to physical parameters (the
        samplers do that on the GPU, one elementwise pass over the device block, instead of one host call per step)."""
        coords = np.asarray(coords)
        if coords.shape[0] == 0:
            return
        if self._frozen is not None:     # continuing a finished chain: back to the appendable form first
            frozen, self._frozen, self.nsteps = self._frozen, None, 0
            for n in self.NAMES:
                with open(self._path(n), "wb") as f:
                    np.ascontiguousarray(frozen[n], self._dtype).tofile(f)
            self.nsteps = frozen["chain"].shape[0]
        if transformed is not None:
            tr = np.asarray(transformed).reshape(coords.shape)
        else:
            tr = coords if self.transform is None else np.stack(
                [np.atleast_2d(self.transform(c.astype(np.float32))) for c in coords])
        if self.nsteps == 0:
            self.nwalkers, self.ndim = coords.shape[1], coords.shape[2]
        for n, block in (("chain", coords), ("chain_transformed", tr), ("log_prob", np.asarray(log_prob))):
            with open(self._path(n), "ab") as f:
                for step in block:       # step by step: no second copy of a GB-sized block when the element type differs
                    np.ascontiguousarray(step, self._dtype).tofile(f)
        self.nsteps += coords.shape[0]
        self._maps = {}

    def save(self):
        """Make the appended blocks durable: a few bytes of metadata (the blocks themselves are already on disk)."""
        import json
        with open(self.base + ".meta.json", "w") as f:
            json.dump({"nsteps": self.nsteps, "nwalkers": self.nwalkers, "ndim": self.ndim, "dtype": self._raw}, f)

    def finalize(self, max_bytes=2 << 30):
        """End of the run: the portable single-file forms with the reference's dataset names."""
        self.save()
        if self.nsteps == 0 or self._frozen is not None:
            return
        if self.nsteps * self.nwalkers * (2 * self.ndim + 1) * 8 > max_bytes:
            return                       # a chain this large stays in its raw appendable form
        f64 = lambda a: np.asarray(a, np.float64)        # the reference's chain files hold float64 (emcee's default dtype)
        np.savez(self.base + ".npz", chain=f64(self.chain), chain_transformed=f64(self.chain_transformed), log_prob=f64(self.log_prob))
        try:
            import h5py
            with h5py.File(self.base + ".h5", "w") as f:
                g = f.create_group("mcmc")
                for n in self.NAMES:
                    g.create_dataset(n, data=f64(self._array(n)))
        except ImportError:
            pass

    def get_last_sample(self):
        return np.asarray(self.chain[-1])

    def get_value(self, name, discard=0, flat=False, thin=1):
        v = {"chain": self.chain, "chain_transformed": self.chain_transformed, "samples": self.chain_transformed,
             "log_prob": self.log_prob}[name][discard::thin]
        v = np.asarray(v)
        return v.reshape((-1,) + v.shape[2:]) if flat else v

    def get_log_prob(self, discard=0, flat=False, thin=1):
        return self.get_value("log_prob", discard, flat, thin)

    def get_autocorr_time(self, quiet=True, **kw):
        return integrated_time(thin_for_tau(self.chain))


def transform_block_device(transform, xb):
    """chain_transformed of a device block [steps, walkers, ndim]: the prior map (linna/util.py:323-347) applied on the
    GPU in one elementwise pass, or None when ``transform`` is not one of this package's prior maps (then ChainStore
    falls back to one host call per step, as the reference's back-end does, linna/sampler.py:356)."""
    if transform is None or not hasattr(transform, "priors") or not hasattr(transform, "_col") or not xb.is_cuda:
        return None
    nd = xb.shape[-1]
    flat = xb.reshape(-1, nd).to(torch.float32)
    cols = [transform._col(flat[:, i], p) for i, p in enumerate(transform.priors)]
    return torch.stack(cols, dim=1).reshape(xb.shape)


_PINNED = {}


def to_host_pinned(name, t):
    """Device block -> numpy through a cached pinned staging tensor (a fresh pageable ``.cpu()`` of a 1.2 GB block moves
    at ~2 GB/s; the pinned copy at PCIe rate).  The returned array aliases the staging buffer: it is valid until the next
    call with the same ``name``."""
    key = (name, t.device.index)
    buf = _PINNED.get(key)
    if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
        buf = torch.empty(int(t.numel() * 1.25) + 16, dtype=t.dtype, pin_memory=True)
        _PINNED[key] = buf
    dst = buf[:t.numel()].view(t.shape)
    dst.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return dst.numpy()


def thin_for_tau(chain, max_walkers=2048, max_bytes=1 << 30):
    """The part of a stored chain [steps, walkers, ndim] that feeds the autocorrelation estimate: the walker-averaged
    autocorrelation function of a few thousand walkers is as good as that of 10^5, and the FFT buffers are several
    times the size of what goes in -- so the walkers are sub-sampled (evenly) until the block fits ``max_bytes``."""
    steps, nw, nd = chain.shape
    keep = min(nw, max_walkers, max(1, int(max_bytes // max(steps * nd * 8, 1))))
    if keep >= nw:
        return np.asarray(chain)
    idx = np.linspace(0, nw - 1, keep).astype(np.int64)
    return np.asarray(chain[:, idx, :])


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
        self._first = 0               # iteration index of self._chain[0] (earlier snapshots were dropped after a flush)
        self.naccepted = torch.zeros(self.nwalkers, device=self.device)

    def drop_before(self, iteration):
        """Forget the device snapshots of iterations < ``iteration`` (they have been flushed to the chain store): the
        sampler then holds ``check_every`` iterations on the GPU, not the whole run."""
        k = max(0, min(int(iteration) - self._first, len(self._chain)))
        del self._chain[:k], self._lnp[:k]
        self._first += k

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

    def device_block(self, start=0):
        """(positions [steps, walkers, ndim], lnP [steps, walkers]) of the iterations >= ``start`` still held, on the GPU."""
        k = max(0, int(start) - self._first)
        if len(self._chain) <= k:
            return (torch.zeros((0, self.nwalkers, self.ndim), device=self.device),
                    torch.zeros((0, self.nwalkers), device=self.device))
        return torch.stack(self._chain[k:]), torch.stack(self._lnp[k:])

    def get_chain(self, flat=False, start=0):
        """Stored positions [steps, walkers, ndim] on the host, from iteration ``start`` on."""
        c = self.device_block(start)[0].cpu().numpy().astype(np.float64)
        return c.reshape(-1, self.ndim) if flat else c

    def get_log_prob(self, flat=False, start=0):
        l = self.device_block(start)[1].cpu().numpy().astype(np.float64)
        return l.reshape(-1) if flat else l

    def get_autocorr_time(self, tol=0, **kw):
        return integrated_time(self.get_chain())

    @property
    def acceptance_fraction(self):
        return (self.naccepted / max(self.iteration, 1)).cpu().numpy()


class HMCSampler:
    """The reference's sampler wrapper (linna/sampler.py:389-554).

    ``method="emcee"`` / ``"zeus"``: burn-in with re-selection of the best positions, then ensemble sampling until
    tau*ntimes < n,  |d tau|/tau < tautol  and the half-chain mean/std test pass (checked every 100 iterations).
    ``method="hmc"``: ``nwalkers`` independent HMC chains advanced together on the GPU
    (``linna.HMCSampler.HMCSampler.sample_chains``: ``samp_steps`` leapfrog steps of size ``samp_eps`` per sample), same
    storage and convergence logic.

    Under ``torch.distributed`` (one process per GPU, world > 1) the walkers / chains are sharded over the ranks: every
    rank runs its own sub-ensemble of nwalkers/world walkers (independent units: no collective inside an iteration), the
    blocks of all ranks are gathered at every convergence check, rank 0 stores the chain and takes the decision, which is
    broadcast.  This is the GPU form of the reference's one-walker-per-MPI-rank farm (linna/util.py:100-256)."""
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
        from . import parallel
        if method not in ("emcee", "zeus", "hmc"):
            raise NotImplementedError("method %r: 'emcee' / 'zeus' (ensemble stretch move) and 'hmc' (batched chains) are "
                                      "implemented; NUTS is a stub in the reference too (linna/sampler.py:14-21)" % (method,))
        rank, world = parallel.world()
        dev = torch.device("cuda", torch.cuda.current_device())
        filename = os.path.join(outdir, "chhmc.h5" if method == "hmc" else self.FILENAME)     # linna/sampler.py:466-471
        store = ChainStore(filename, self.transform, fresh=bool(overwrite) and rank == 0)
        if world > 1:
            torch.distributed.barrier()
            if rank != 0:
                store = ChainStore(filename, self.transform)
        x0 = np.asarray(self.x0, np.float64)
        resume = store.exists()
        if resume:
            print("init from previous")
            x0 = store.get_last_sample()
        lo, hi = parallel.shard_rows(self.nwalkers, rank, world)
        nloc = hi - lo
        if method != "hmc" and nloc < 2 * self.nparams and world > 1:
            raise ValueError("%d walkers over %d ranks leaves %d per rank: the stretch move needs >= 2 x ndim = %d walkers "
                             "in every sub-ensemble" % (self.nwalkers, world, nloc, 2 * self.nparams))

        def gathered(xb, lb):
            """[steps, local walkers, ...] blocks of every rank -> [steps, all walkers, ...] (the same on all ranks)."""
            if world == 1:
                return xb, lb
            xg = parallel.gather_rows(xb.permute(1, 0, 2).contiguous()).permute(1, 0, 2).contiguous()
            lg = parallel.gather_rows(lb.permute(1, 0).contiguous()).permute(1, 0).contiguous()
            return xg, lg

        print("start", flush=True)
        if method == "hmc":
            return self._sample_hmc(store, x0, lo, hi, rank, world, nsamp, samp_steps, samp_eps, ntimes, tautol, meanshift, stdshift,
                                    nk, check_every, gathered, dev)
        self.sampler = EnsembleSampler(nloc, self.nparams, self.lnp, seed=None if world == 1 else 1000003 * (rank + 1) + 17)
        if not incremental:
            self.sampler.run_mcmc(x0[lo:hi], nsamp)
            xb, _ = gathered(*self.sampler.device_block(0))
            chain = xb.reshape(-1, self.nparams).cpu().numpy().astype(np.float64)
            return chain if self.transform is None else np.array([self.transform(c.astype(np.float32)) for c in chain])
        if not resume:
            print("burnin...", flush=True)
            burn = EnsembleSampler(nloc, self.nparams, self.lnp, seed=None if world == 1 else 7 * (rank + 1) + 1)
            burn.run_mcmc(x0[lo:hi], burnin)
            flat, lp = burn.get_chain(flat=True), burn.get_log_prob(flat=True)
            pos = flat[np.argsort(lp)[::-1][:int(50 * nloc)]]
            x0_loc = pos[np.random.randint(0, len(pos), nloc), :]
            print("burnin done...", flush=True)
        else:
            x0_loc = x0[lo:hi]
        old_tau = np.inf
        saved = 0

        def flush_and_check(final=False):
            nonlocal old_tau, saved
            xb, lb = gathered(*self.sampler.device_block(saved))
            saved = self.sampler.iteration
            self.sampler.drop_before(saved)       # the flushed snapshots leave the GPU
            stop = False
            if rank == 0:
                tb = transform_block_device(self.transform, xb)
                store.extend(to_host_pinned("x", xb), to_host_pinned("lnp", lb),   # only the new steps are appended to the files
                             transformed=None if tb is None else to_host_pinned("theta", tb))
                store.save()
                if not final:
                    tau = integrated_time(thin_for_tau(store.chain))
                    if np.isnan(np.sum(tau)) and saved > 10:
                        stop = True
                    else:
                        converged = np.all(tau * ntimes < store.iteration)
                        converged &= np.all(np.abs(old_tau - tau) / tau < tautol)
                        keep = max(int(nk * np.mean(tau)), 2)
                        converged &= checkmeanstd(np.asarray(store.chain[-keep:]), meanshift=meanshift, stdshift=stdshift)
                        print("max, min tau diff, max tau, ninter: {0}, {1}, {2}, {3}\n".format(
                            np.max(np.abs(old_tau - tau) / tau), np.min(np.abs(old_tau - tau) / tau), np.max(tau), store.iteration),
                            flush=True)
                        stop = bool(converged)
                        old_tau = tau
            if world > 1:
                box = [stop]
                torch.distributed.broadcast_object_list(box, src=0)
                stop = bool(box[0])
            return stop

        for _ in self.sampler.sample(x0_loc, int(nsamp)):
            if self.sampler.iteration % check_every:
                continue
            if flush_and_check():
                break
        if self.sampler.iteration > saved:
            flush_and_check(final=True)
        if rank == 0:
            store.finalize()
        if world > 1:
            torch.distributed.barrier()
            store = ChainStore(filename, self.transform)     # every rank (rank 0 too) reads the finished chain the same way
        self.sampler = None
        return store

    def _sample_hmc(self, store, x0, lo, hi, rank, world, nsamp, samp_steps, samp_eps, ntimes, tautol, meanshift, stdshift, nk,
                    check_every, gathered, dev):
        """``method="hmc"``: batched device HMC in blocks of ``check_every`` samples with the ensemble branch's storage
        and convergence logic (the reference's hmc branch calls private helpers with mismatched signatures)."""
        from .HMCSampler import HMCSampler as _TorchHMC
        if not getattr(self.lnp, "fused", False):
            raise NotImplementedError("method='hmc' needs the built-in Gaussian likelihood (fused lnP + gradient kernel)")
        steps = int(samp_steps) if samp_steps else 5
        eps = float(samp_eps) if samp_eps else 0.1
        mass = torch.ones(self.nparams) if self.m is None else torch.as_tensor(np.asarray(self.m, np.float32)).reshape(-1)
        hmc = _TorchHMC(self.lnp, torch.as_tensor(x0[lo:hi], dtype=torch.float32), mass, device="cuda")
        old_tau, done = np.inf, 0
        while done < int(nsamp):
            nblk = int(min(check_every, int(nsamp) - done))
            xs, ls, _ = hmc.sample_chains(nblk, steps, eps, seed=12345 + 104729 * (done // max(check_every, 1)), distributed=False)
            done += nblk
            xb, lb = gathered(xs, ls)
            stop = False
            if rank == 0:
                tb = transform_block_device(self.transform, xb)
                store.extend(to_host_pinned("x", xb), to_host_pinned("lnp", lb),
                             transformed=None if tb is None else to_host_pinned("theta", tb))
                store.save()
                tau = integrated_time(thin_for_tau(store.chain))
                if not (np.isnan(np.sum(tau)) and done > 10):
                    converged = np.all(tau * ntimes < store.iteration) and np.all(np.abs(old_tau - tau) / tau < tautol)
                    keep = max(int(nk * np.mean(tau)), 2)
                    converged = converged and checkmeanstd(np.asarray(store.chain[-keep:]), meanshift=meanshift, stdshift=stdshift)
                    stop, old_tau = bool(converged), tau
                else:
                    stop = True
            if world > 1:
                box = [stop]
                torch.distributed.broadcast_object_list(box, src=0)
                stop = bool(box[0])
            if stop:
                break
        if rank == 0:
            store.finalize()
        if world > 1:
            torch.distributed.barrier()
            store = ChainStore(os.path.join(os.path.dirname(store.base), os.path.basename(store.base)), self.transform)
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
