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
    device-resident form used for the LSST-shaped HMC config (C4).
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
    def sample_chains(self, num_samps, num_steps, step_size, generator=None, thin=1):
        """All chains resident on the GPU: positions, momenta, Metropolis test and RNG stay on the
        device; per sample only ``num_steps + 1`` fused lnP+grad launches and a handful of
        elementwise updates are issued.  Returns (samples [num_samps/thin, C, n] latent positions,
        lnP [.., C], acceptance fraction)."""
        dev = torch.device("cuda", torch.cuda.current_device())
        x = self.x.detach().to(dev, torch.float32).reshape(-1, self.x.shape[-1]).contiguous()
        m = self.m.to(dev)
        vg = self.lnP.value_and_grad
        lnP, grad = vg(x)
        keep_x, keep_l, nacc = [], [], torch.zeros((), device=dev)
        for s in range(num_samps):
            p = torch.randn(x.shape, device=dev, generator=generator) * torch.sqrt(m)
            H0 = (0.5 * p.square() / m).sum(-1) - lnP
            xn, g, l = x, grad, lnP
            p = p + 0.5 * step_size * g
            for i in range(num_steps):
                xn = xn + step_size * (p / m)
                l, g = vg(xn)
                p = p + (step_size if i + 1 < num_steps else 0.5 * step_size) * g
            H1 = (0.5 * p.square() / m).sum(-1) - l
            acc = torch.rand(x.shape[0], device=dev, generator=generator) < torch.exp(torch.clamp(H0 - H1, max=0.0))
            acc = acc & torch.isfinite(l)
            x = torch.where(acc[:, None], xn, x)
            grad = torch.where(acc[:, None], g, grad)
            lnP = torch.where(acc, l, lnP)
            nacc += acc.float().mean()
            if (s + 1) % thin == 0:
                keep_x.append(x.clone())
                keep_l.append(lnP.clone())
        self.x = x
        return torch.stack(keep_x), torch.stack(keep_l), float(nacc / max(num_samps, 1))


HMCSampler.__module__ = "linna.HMCSampler"
