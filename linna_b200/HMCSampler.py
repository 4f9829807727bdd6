"""Hamiltonian Monte Carlo on the emulator posterior.

``HMCSampler`` keeps the reference's standalone sampler interface and algorithm
(``linna/HMCSampler.py:6-68``): ``HMCSampler(lnP, x0, m, transform=None, device='cpu')
.sample(num_samps, num_steps, step_size) -> list[dict]`` -- leapfrog with ``num_steps + 1``
lnP+gradient evaluations per sample and one Metropolis test.

Two things are different:
  * when ``lnP`` is a ``linna.util.Log_prob`` the value and gradient come from ONE fused kernel
    launch (``Log_prob.value_and_grad``) instead of a forward pass plus ``torch.autograd.grad``;
  * ``x0`` may be ``[C, n]``: C independent chains advanced together, each with its own
    accept/reject (the reference sums the Hamiltonian over everything, so a batched ``x0`` there
    would be accepted or rejected jointly -- SURVEY 3.2).  ``sample_chains`` is the fully
    device-resident form used for the LSST-shaped HMC config (C4): hand-written leapfrog /
    Metropolis kernels around the fused lnP+gradient launch, chains sharded over the GPUs of a
    ``torchrun`` job.
"""
import numpy as np
import torch
from tqdm.auto import tqdm


class HMCSampler:
    def __init__(self, lnP, x0, m, transform=None, device="cpu"):
        self.lnP = lnP
        self.x0 = x0.to(dtype=torch.float32, device=device)
        self.x = self.x0.clone()
        self.m = m.to(dtype=torch.float32, device=device)
        self.device = device
        self.transform = transform if transform is not None else (lambda x: x)

    # -- lnP and gradient ------------------------------------------------------------------
    def _value_and_grad(self, x):
        vg = getattr(self.lnP, "value_and_grad", None)
        if vg is not None and getattr(self.lnP, "fused", False) and getattr(self.lnP, "externalloglike", None) is None:
            xd = x.detach().reshape(-1, x.shape[-1]).to("cuda", torch.float32)
            lnp, g = vg(xd)
            lnp, g = lnp.to(x.device), g.to(x.device).reshape(x.shape)
            return (lnp if x.dim() > 1 else lnp.reshape(())), g
        x = x.detach().clone().requires_grad_()
        lnp = self.lnP(x)
        g = torch.autograd.grad(lnp.sum(), x)[0]
        return lnp.detach(), g

    def _kinetic(self, p):
        k = 0.5 * torch.square(p) / self.m
        return k.sum(dim=-1) if p.dim() > 1 else k.sum()

    def sample(self, num_samps, num_steps, step_size):
        """Reference-shaped entry point: list of dicts {'x','lnP','accpet_ratio','accept_prob','accepted'}."""
        chain = []
        batched = self.x.dim() > 1
        for _ in tqdm(range(num_samps)):
            x = self.x.detach().clone()
            p = torch.randn(x.shape, device=self.device) * torch.sqrt(self.m)
            lnP, grad = self._value_and_grad(x)
            prev_lnP = lnP.clone()
            H_init = (self._kinetic(p) - lnP).cpu().numpy()
            # leapfrog: half kick, (num_steps-1) x [drift, kick], drift, half kick
            p = p + 0.5 * grad * step_size
            x = x + (p / self.m) * step_size
            lnP, grad = self._value_and_grad(x)
            for _i in range(1, num_steps):
                p = p + grad * step_size
                x = x + (p / self.m) * step_size
                lnP, grad = self._value_and_grad(x)
            p = p + 0.5 * grad * step_size
            H_prime = (self._kinetic(p) - lnP).cpu().numpy()
            accept_ratio = np.exp(np.minimum(H_init - H_prime, 0))
            if not batched:
                accept_prob = min(accept_ratio, 1)
                if np.random.uniform() < accept_prob:
                    self.x = x
                    chain.append({"x": self.transform(x).detach().cpu().numpy(), "lnP": lnP.cpu().numpy(),
                                  "accpet_ratio": accept_ratio, "accept_prob": accept_prob, "accepted": True})
                else:
                    chain.append({"x": self.transform(self.x).detach().cpu().numpy(), "lnP": prev_lnP.cpu().numpy(),
                                  "accpet_ratio": accept_ratio, "accept_prob": accept_prob, "accepted": False})
            else:
                accept_prob = np.minimum(accept_ratio, 1)
                acc = np.random.uniform(size=accept_prob.shape) < accept_prob
                acc_t = torch.from_numpy(acc).to(self.x.device)
                self.x = torch.where(acc_t[:, None], x, self.x)
                lnp_out = torch.where(acc_t, lnP, prev_lnP)
                chain.append({"x": self.transform(self.x).detach().cpu().numpy(), "lnP": lnp_out.cpu().numpy(),
                              "accpet_ratio": accept_ratio, "accept_prob": accept_prob, "accepted": acc})
        return chain

    @torch.no_grad()
    def sample_chains(self, num_samps, num_steps, step_size, generator=None, thin=1, seed=None, distributed=None):
        """All chains resident on the GPU: positions, momenta, Metropolis test and RNG stay on the device.  One sample of
        every chain is ``num_steps`` fused lnP+gradient launches (``linna_lnp_grad``) plus ONE sampler kernel per leapfrog
        step (``csrc/sampler_kernels.cu``: momentum draw + Hamiltonian + half kick + drift; kick + drift; half kick +
        Hamiltonian + per-chain Metropolis select) -- the reference's leapfrog (linna/HMCSampler.py:23-66) batched over
        chains, each chain with its own accept / reject.

        Under ``torch.distributed`` (one process per GPU, world > 1; ``distributed=False`` switches it off) the chains are
        sharded over the ranks -- they are independent, no collective touches the sampling loop -- and the returned
        samples are the chains of ALL ranks, gathered once at the end.

        Returns (samples [num_samps/thin, C, n] latent positions, lnP [.., C], acceptance fraction)."""
        from . import engine as _engine
        from . import parallel
        dev = torch.device("cuda", torch.cuda.current_device())
        x_all = self.x.detach().to(dev, torch.float32).reshape(-1, self.x.shape[-1]).contiguous()
        rank, world = parallel.world()
        if distributed is False:
            rank, world = 0, 1
        lo, hi = parallel.shard_rows(x_all.shape[0], rank, world)
        x = x_all[lo:hi].contiguous()
        m = self.m.to(dev, torch.float32).reshape(-1).contiguous()
        if m.numel() == 1:
            m = m.expand(x.shape[1]).contiguous()
        vg = self.lnP.value_and_grad
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,), generator=generator, device=generator.device if generator is not None else "cpu").item())
        seed = int(seed) + 7919 * rank                                  # every rank its own Philox streams
        lnP, grad = vg(x)
        lnP, grad = lnP.contiguous(), grad.contiguous()
        nacc = torch.zeros(x.shape[0], dtype=torch.float32, device=dev)
        keep_x, keep_l = [], []
        for s in range(num_samps):
            p, xn, H0 = _engine.hmc_begin(x, lnP, grad, m, step_size, seed, 8 * s)
            for i in range(num_steps):
                l, g = vg(xn)
                if i + 1 < num_steps:
                    _engine.hmc_step(p, xn, g, m, step_size)
            _engine.hmc_end(x, lnP, grad, xn, l.contiguous(), g.contiguous(), p, m, H0, step_size, seed, 8 * s + 4, nacc)
            if (s + 1) % thin == 0:
                keep_x.append(x.clone())
                keep_l.append(lnP.clone())
        xs, ls = torch.stack(keep_x), torch.stack(keep_l)
        accfrac = nacc / max(num_samps, 1)
        if world > 1:      # final chain assembly: the only collective of the run
            xs = parallel.gather_rows(xs.permute(1, 0, 2).contiguous()).permute(1, 0, 2).contiguous()
            ls = parallel.gather_rows(ls.permute(1, 0).contiguous()).permute(1, 0).contiguous()
            accfrac = parallel.gather_rows(accfrac)
            x_all = parallel.gather_rows(x)
        else:
            x_all = x
        self.x = x_all
        return xs, ls, float(accfrac.mean())


HMCSampler.__module__ = "linna.HMCSampler"
