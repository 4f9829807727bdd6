"""Device-side training state for one emulator: the flat parameter / AdamW-moment vectors and the
fused optimiser step (``csrc/fused_ffma.cu`` PROG_TRAIN + ``csrc/train_kernels.cu``).

Replaces the body of the reference's training inner loop -- ``zero_grad``, forward, ``loss_fn``,
``backward``, ``AdamW.step`` and the per-step ``loss.item()`` sync
(``linna/predictor_gpu.py:273-288``) -- by 3 kernel launches with no host synchronisation.
Data-parallel training adds one NCCL all-reduce of the flat gradient between the weight-gradient
launch and the stand-alone AdamW kernel.
"""
import os

import numpy as np
import torch

from . import arch
from . import engine as _engine


def flatten_module(model):
    """Parameters of a linna.nn model as one flat float32 vector in state_dict order."""
    sd = model.state_dict()
    keys = [k for k, _ in arch.state_dict_shapes(model.KIND, model.in_size, model.out_size)]
    return torch.cat([sd[k].detach().reshape(-1).to(torch.float32).cpu() for k in keys])


def unflatten_into_module(model, flat):
    """Write a flat parameter vector back into the module's tensors (in place)."""
    flat = flat.detach().cpu()
    sd = model.state_dict()
    o = 0
    with torch.no_grad():
        for k, shp in arch.state_dict_shapes(model.KIND, model.in_size, model.out_size):
            n = int(np.prod(shp))
            sd[k].copy_(flat[o:o + n].reshape(shp).to(sd[k].device))
            o += n


class FusedTrainer:
    """Owns the packed emulator on one GPU plus flat (params, exp_avg, exp_avg_sq, grad) vectors."""

    def __init__(self, model, X_transform, y_transform, aux, max_batch, device_index=None, lr=1e-3,
                 weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=1):
        from .predictor_gpu import _transform_constants
        if not torch.cuda.is_available():
            raise RuntimeError("FusedTrainer: no CUDA device -- linna_b200 has no CPU fallback")
        self.model = model
        self.dev = torch.cuda.current_device() if device_index is None else int(device_index)
        self.device = torch.device("cuda", self.dev)
        data_hat, icov_hat, sigma, y_mean, y_std, ypos = aux.constants()
        xm, xs, log10, _, _, _ = _transform_constants(X_transform, None, model.in_size, model.out_size)
        sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
        self.engine = _engine.Engine(model.KIND, model.in_size, model.out_size, sd, xm, xs, y_mean, y_std,
                                     dolog10index=log10, ypositive=ypos, sigma=sigma, device=self.dev)
        self.engine.train_setup(data_hat, icov_hat, int(max_batch))
        self.n_out = model.out_size
        self.p = flatten_module(model).to(self.device)
        assert self.p.numel() == self.engine.n_params
        self.m = torch.zeros_like(self.p)
        self.v = torch.zeros_like(self.p)
        self.g = torch.zeros_like(self.p)
        self.t = 0
        self.lr, self.weight_decay, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        self.pg, self.world = process_group, int(world_size)
        self._peer = None
        if self.world > 1:
            self._setup_peer_reduce()
        self._loss_mean = torch.zeros(1, dtype=torch.float32, device=self.device)

    # -- data ------------------------------------------------------------------------------
    def chisq_md(self, X, Y):
        """max(chi2(target, data), n_out/2) per row -- depends on the targets only, so it is computed
        once per data set instead of once per step (linna/util.py:1080-1086)."""
        return torch.clamp(self.engine.train_chisq(X, Y, 1), min=0.5 * self.n_out)

    # -- one optimiser step ----------------------------------------------------------------
    def step(self, X, Y, cmd, loss_out=None):
        """X [B, n_in] physical parameters, Y [B, n_out] physical targets, cmd [B]; all CUDA.
        Returns the mean loss as a 1-element device tensor (no sync)."""
        self.t += 1
        out = self._loss_mean if loss_out is None else loss_out
        if self.world > 1 and self._peer is not None:
            # gradient-out step into this rank's half of the symmetric buffer, then ONE kernel that waits for the peers'
            # gradients, averages them from peer memory (NVLink) and applies AdamW: no NCCL call on the step
            pr = self._peer
            n = self.p.numel()
            off = (pr["k"] & 1) * pr["stride"]
            pr["k"] += 1
            g = pr["buf"][off:off + n]
            self.engine.train_step(X, Y, cmd, None, None, None, g, self.t, self.lr, self.betas, self.eps,
                                   self.weight_decay, fuse_adam=False, loss_mean=out)
            self.engine.train_adamw_peer(self.p, self.m, self.v, pr["grad_ptrs"], off, pr["avg"], pr["pad_ptrs"], pr["slot"],
                                         self.world, pr["rank"], self.t, self.lr, self.betas, self.eps, self.weight_decay)
        elif self.world > 1:
            import torch.distributed as dist
            self.engine.train_step(X, Y, cmd, None, None, None, self.g, self.t, self.lr, self.betas, self.eps,
                                   self.weight_decay, fuse_adam=False, loss_mean=out)
            dist.all_reduce(self.g, op=dist.ReduceOp.AVG, group=self.pg)      # NCCL over NVLink
            self.engine.train_adamw(self.p, self.m, self.v, self.g, self.t, self.lr, self.betas, self.eps,
                                    self.weight_decay)
        else:
            self.engine.train_step(X, Y, cmd, self.p, self.m, self.v, None, self.t, self.lr, self.betas, self.eps,
                                   self.weight_decay, fuse_adam=True, loss_mean=out)
        return out

    def _setup_peer_reduce(self):
        """Symmetric gradient buffer (two halves, alternating per step) whose peer pointers every rank holds: the optimiser
        kernel reads the other ranks' gradients through them.  Falls back to the NCCL all-reduce (with a note) when
        symmetric memory cannot be set up -- e.g. the stand-in ranks of a single-GPU test."""
        import torch.distributed as dist
        if os.environ.get("LINNA_DP_NCCL") or not (dist.is_available() and dist.is_initialized()):
            return
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.pg if self.pg is not None else dist.group.WORLD
            if dist.get_world_size(group) != self.world:
                return
            n = self.p.numel()
            stride = (n + 31) // 32 * 32                       # 128-byte aligned halves: 16-byte loads stay aligned
            buf = symm.empty(3 * stride, dtype=torch.float32, device=self.device)   # two gradient halves + the averaged slices
            buf.zero_()
            hdl = symm.rendezvous(buf, group)
            slot = int(hdl.signal_pad_size) // 4 - 64          # the last words of the pad: clear of torch's own barriers
            if slot < 0 or self.world > 16:
                return
            hdl.barrier()
            two_phase = (self.world > 2 or bool(os.environ.get("LINNA_DP_TWO_PHASE"))) and not os.environ.get("LINNA_DP_ONE_PHASE")
            self._peer = dict(buf=buf, hdl=hdl, stride=stride, k=0, slot=slot, rank=int(hdl.rank), avg=2 * stride if two_phase else -1,
                              grad_ptrs=int(hdl.buffer_ptrs_dev), pad_ptrs=int(hdl.signal_pad_ptrs_dev))
        except Exception as e:      # noqa: BLE001 -- any failure of the experimental API: keep the NCCL path
            print("linna_b200: peer-memory gradient reduction unavailable (%s: %s); using the NCCL all-reduce" % (type(e).__name__, e),
                  flush=True)
            self._peer = None

    def kernel_path(self):
        """'tc' when the optimiser step runs on the tensor-core (tcgen05) training kernels, else 'ffma'."""
        return self.engine.last_train_kernel() or "none"

    # -- validation metric (Val_metric_fn, linna/util.py:1118-1127) ------------------------
    def val_metric(self, X, Y, cmd):
        mnn = self.engine.train_chisq(X, Y, 0)
        nnd = self.engine.train_chisq(X, Y, 2)
        loss = mnn / cmd
        frac = torch.abs(nnd / cmd - 1)
        return torch.stack([torch.median(loss), torch.max(frac), torch.median(frac)])

    # -- state exchange with the torch module / optimizer ----------------------------------
    def reset_optimizer(self, weight_decay=1e-4):
        """A fresh optimiser, as the reference's `AdamW(lr, weight_decay=1E-4)` after a re-initialisation or a roll-back
        (predictor_gpu.py:329, :358): zero moments, step 0, and the weight decay back at its default (EarlyStopping may
        have halved or doubled it)."""
        self.m.zero_(), self.v.zero_()
        self.t = 0
        if weight_decay is not None:
            self.weight_decay = float(weight_decay)

    def load_from_module(self):
        """Adopt the module's current parameters (after init_weight / load_checkpoint)."""
        self.p.copy_(flatten_module(self.model).to(self.device))
        self.engine.train_load_params(self.p)

    def sync_to_module(self):
        unflatten_into_module(self.model, self.p)

    def optim_state_dict(self):
        """An ``AdamW.state_dict()``-shaped dict (what the reference stores as 'optim_dict')."""
        params = list(self.model.parameters())
        state, o = {}, 0
        m, v = self.m.detach().cpu(), self.v.detach().cpu()
        for i, prm in enumerate(params):
            n = prm.numel()
            state[i] = {"step": torch.tensor(float(self.t)), "exp_avg": m[o:o + n].reshape(prm.shape).clone(),
                        "exp_avg_sq": v[o:o + n].reshape(prm.shape).clone()}
            o += n
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": True,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_optim_state_dict(self, od):
        params = list(self.model.parameters())
        st = od.get("state", {})
        if len(st) != len(params):
            self.reset_optimizer()
            return
        m = torch.cat([st[i]["exp_avg"].reshape(-1).float().cpu() for i in range(len(params))])
        v = torch.cat([st[i]["exp_avg_sq"].reshape(-1).float().cpu() for i in range(len(params))])
        self.m.copy_(m.to(self.device)), self.v.copy_(v.to(self.device))
        self.t = int(float(st[0]["step"]))

    def commit(self):
        """End of training: parameters back into the module and into the engine's host copies."""
        self.sync_to_module()
        self.engine.train_commit(self.p.detach().cpu().numpy())


class _LossTerms(torch.autograd.Function):
    """loss rows with d loss / d y_pred attached (the other two outputs carry no gradient to y_pred that the reference's
    Loss_fn uses: chisqMd depends on the targets only, chisqnnd is a diagnostic)."""

    @staticmethod
    def forward(ctx, y_pred, y_target, consts):
        data_hat, icov, sigma, y_mean, y_std, ypos = consts
        want = y_pred.requires_grad
        loss, md, nnd, g = _engine.loss_terms(y_pred.detach().contiguous(), y_target.detach().contiguous(), data_hat, icov, sigma,
                                              y_mean, y_std, ypos, want_grad=want)
        if want:
            ctx.save_for_backward(g)
        ctx.mark_non_differentiable(md, nnd)
        return loss, md, nnd

    @staticmethod
    def backward(ctx, gl, gmd, gnnd):
        (g,) = ctx.saved_tensors
        return gl.unsqueeze(-1) * g, None, None


def loss_terms(aux, y_pred, y_target):
    """``Auxilleryfunc.__call__`` on free-standing tensors (linna/util.py:1070-1088): (loss, chisqMd, chisqnnd) per row,
    differentiable with respect to ``y_pred`` -- one kernel launch (``linna_loss_terms``).  CUDA tensors stay on their
    device; host tensors are evaluated on the current GPU and the results returned on the host, like every other
    evaluation in this package (there is no CPU arithmetic path)."""
    if not torch.cuda.is_available():
        raise RuntimeError("Loss_fn / Val_metric_fn: no CUDA device -- linna_b200 has no CPU fallback")
    src = y_pred.device
    dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    key = str(dev)
    cache = aux.__dict__.setdefault("_dev_consts", {})
    if key not in cache:
        data_hat, icov, sigma, y_mean, y_std, ypos = aux.constants()
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)
        cache[key] = (f(data_hat), f(icov), f(sigma), f(y_mean), f(y_std), ypos)
    yp = y_pred.to(dev, torch.float32)
    yt = y_target.to(dev, torch.float32)
    one = yp.dim() == 1
    if one:
        yp, yt = yp.reshape(1, -1), yt.reshape(1, -1)
    loss, md, nnd = _LossTerms.apply(yp, yt, cache[key])
    return loss.to(src), md.to(src), nnd.to(src)


class _EmulatorFunction(torch.autograd.Function):
    """y = emulator(x) with the kernel-computed vector-Jacobian products attached: what torch.autograd does for the
    reference's ``model(X_transform(X))`` / ``Predictor.predict(X, no_grad=False)`` (linna/predictor_gpu.py:279-283,
    :495-496).  Forward = ``linna_predict``; backward = ONE fused forward + backward-data launch (``linna_predict_vjp``)
    plus, when parameters require grad, the weight-gradient launch."""

    @staticmethod
    def forward(ctx, x, eng, out_kind, params_owner, *params):
        ctx.eng, ctx.out_kind, ctx.nparams = eng, out_kind, len(params)
        ctx.want_params = any(p.requires_grad for p in params)
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.pdevs = [p.device for p in params]
        ctx.x_dev = x.device
        xd = x.detach().to(torch.device("cuda", eng.device), torch.float32).contiguous()
        ctx.save_for_backward(xd)
        ctx.need_x = x.requires_grad
        return eng.predict(xd, out_kind).to(x.device)

    @staticmethod
    def backward(ctx, gy):
        (xd,) = ctx.saved_tensors
        gth, gp = ctx.eng.predict_vjp(xd, gy.to(xd.device, torch.float32).contiguous(), ctx.out_kind, want_params=ctx.want_params)
        grads = [None] * ctx.nparams
        if gp is not None:
            o = 0
            for i, shp in enumerate(ctx.shapes):
                n = int(np.prod(shp))
                grads[i] = gp[o:o + n].reshape(shp).to(ctx.pdevs[i])
                o += n
        return (gth.to(ctx.x_dev) if ctx.need_x else None, None, None, None) + tuple(grads)


def emulator_forward_autograd(model, x):
    """``model(x)`` under autograd (x [B, n_in] on the GPU; input and / or parameters may require grad)."""
    eng = model.bare_engine(x.device.index)
    params = [p for _, p in model.named_parameters()]
    if any(p.requires_grad for p in params):
        if getattr(eng, "_vjp_batch", 0) < x.shape[0]:       # row-major activation store for the weight-gradient kernel
            eng.train_setup(np.zeros(model.out_size, np.float32), np.eye(model.out_size, dtype=np.float32), int(x.shape[0]))
            eng._vjp_batch = int(x.shape[0])
    return _EmulatorFunction.apply(x, eng, _engine.LINNA_OUT_YHAT, model, *params)


def predict_with_grad(pred, X):
    """``Predictor.predict(X, no_grad=False)`` for a tensor that requires grad: d predict / d X through the kernel
    (the emulator weights are constants here, as in the reference's sampling-time use, linna/util.py:1012)."""
    eng = pred._get_engine()
    one = X.dim() == 1
    X2 = X.reshape(1, -1) if one else X
    y = _EmulatorFunction.apply(X2, eng, _engine.LINNA_OUT_Y, None)
    return y.view(-1) if one else y
