"""``Predictor`` and ``EarlyStopping`` with the reference's interface
(``linna/predictor_gpu.py:19-150``, ``:153-505``), evaluated by the fused sm_100a kernels.

``Predictor.predict`` = X_transform -> emulator -> y_transform in ONE kernel launch for the whole
batch (reference: three Python-level stages, ``predictor_gpu.py:479-500``).  ``Predictor.train`` keeps
the reference's host-side heuristics (LR halving, weight-decay doubling, re-initialisation on a
stalled loss, early stopping, per-epoch checkpoints) around a fused forward/loss/backward/AdamW
device step (``train.FusedTrainer``).
"""
import copy
import os

import numpy as np
import torch

from . import nnutils
from .nn import ChtoModelsimple, ChtoModelv2, ChtoModelv2_linear, ResBlock_batchnorm  # noqa: F401
from .nnutils import load_checkpoint, save_checkpoint  # noqa: F401


class EarlyStopping(object):
    """Decides when training should cool the learning rate (1), stop (2) or raise the weight decay
    (3); 0 = carry on.  Behavioural restatement of linna/predictor_gpu.py:19-150."""

    def __init__(self, mode="min", min_delta=0, patience=10, nqueue=200, percentage=False):
        if mode not in {"min", "max"}:
            raise ValueError("mode " + mode + " is unknown!")
        self.mode, self.min_delta, self.patience, self.nqueue = mode, min_delta, patience, nqueue
        self.percentage = percentage
        self.best = None
        self.best_t = None
        self.num_bad_epochs = 0
        self.cooling = 0
        self.cooling_weight_decay = 0
        self.queue_t, self.queue_v = [], []

    def is_better(self, a, best):
        if self.patience == 0:
            return True
        margin = best * self.min_delta / 100 if self.percentage else self.min_delta
        return a < best - margin if self.mode == "min" else a > best + margin

    @staticmethod
    def _halves(q):
        h = int(0.5 * len(q))
        return float(np.median(q[:h])), float(np.median(q[h:]))

    def step(self, metrics, metrics_t):
        if self.patience == 0:
            return False
        metrics_t = float(metrics_t)
        self.queue_t.append(metrics_t)
        self.queue_v.append(float(metrics))
        self.queue_t = self.queue_t[-self.nqueue:]
        self.queue_v = self.queue_v[-self.nqueue:]
        trend = None
        if len(self.queue_t) > 2:
            t1, t2 = self._halves(self.queue_t)
            v1, v2 = self._halves(self.queue_v)
            trend = (t2 - t1, v2 - v1)
        if self.best is None:
            self.best, self.best_t, self.num_bad_epochs = metrics, metrics_t, 0
            return 0
        if np.isnan(metrics):
            print("nan metric", flush=True)
            self.num_bad_epochs += 1
            return 0
        if self.is_better(metrics, self.best):
            self.num_bad_epochs = 0
            self.cooling = 0
            self.cooling_weight_decay = 0
            self.best, self.best_t = metrics, metrics_t
        else:
            self.num_bad_epochs += 1
            if self.patience * 0.9 <= self.num_bad_epochs < self.patience:
                # close to giving up: ask once for a smaller learning rate, then hold the counter while cooling
                if self.cooling == 0:
                    self.cooling = 1
                    return 1
                if self.cooling > 500:
                    self.cooling = 0
                    self.num_bad_epochs += 5
                else:
                    self.num_bad_epochs -= 1
                    self.cooling += 1
                return 0
            # training loss still falling while validation rises: over-fitting -> more weight decay
            if trend is not None and len(self.queue_t) > 0.5 * self.nqueue and trend[0] < 0 and trend[1] > 0:
                if self.cooling_weight_decay == 0:
                    self.cooling_weight_decay = 1
                    return 3
                if self.cooling_weight_decay > 1000:
                    self.cooling_weight_decay = 0
                    return 0
                self.queue_t, self.queue_v = [], []
                self.cooling_weight_decay += 1
                return 3 if self.cooling_weight_decay % 50 == 0 else 0
        if self.num_bad_epochs >= self.patience:
            return 2
        return 0


def _transform_constants(X_transform, y_transform, in_size, out_size):
    """Pull the diagonal constants out of the reference transform objects (or identity defaults)."""
    xm = getattr(X_transform, "X_mean", None)
    if xm is None:
        x_mean, x_std, log10 = np.zeros(in_size, np.float32), np.ones(in_size, np.float32), None
    else:
        x_mean = xm.detach().cpu().numpy().astype(np.float32).reshape(-1)
        x_std = X_transform.X_std.detach().cpu().numpy().astype(np.float32).reshape(-1)
        log10 = getattr(X_transform, "dolog10index", None)
    ym = getattr(y_transform, "y_mean", None)
    if ym is None:
        y_mean, y_std, ypos = np.zeros(out_size, np.float32), np.ones(out_size, np.float32), False
    else:
        y_mean = ym.detach().cpu().numpy().astype(np.float32).reshape(-1)
        y_std = y_transform.y_std.detach().cpu().numpy().astype(np.float32).reshape(-1)
        ypos = bool(getattr(y_transform, "ypositive", False))
    return x_mean, x_std, log10, y_mean, y_std, ypos


class _Identity:
    def __call__(self, x):
        return x


class Predictor:
    """Emulator + its input/output normalisation (linna/predictor_gpu.py:153-505)."""

    def __init__(self, in_size=None, out_size=None, model=None, optim=None, X_transform=None,
                 y_transform=None, device="cpu", scheduler=None, outdir=None):
        self.in_size = in_size
        self.out_size = out_size
        self.device = device
        self.best_val_loss = float("inf")
        self.outdir = outdir
        if model is None:
            raise NameError("Predictor(model=None): the reference falls back to an undefined `ChtoModel` "
                            "(SURVEY Q9); pass a model from linna.nn")
        self.model = model.to(device)
        self.scheduler = scheduler
        self.optim = optim if optim is not None else torch.optim.AdamW(self.model.parameters())
        self.X_transform = X_transform if X_transform is not None else _Identity()
        self.y_transform = y_transform if y_transform is not None else _Identity()
        self.MKLDNN = False        # kept for API compatibility; there is no oneDNN path here
        self.MKLDNNMODEL = False
        self._engine = None
        self._engine_key = None
        self._sigma = None

    # ---- engine plumbing ----------------------------------------------------------------
    def param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def make_engine(self, device_index, sigma=None):
        """A fresh Engine holding this emulator with its transforms (+ sigma of
        ``Y_invtransform_data`` when the caller wants data-space vectors)."""
        from . import engine as _engine
        xm, xs, log10, ym, ys, ypos = _transform_constants(self.X_transform, self.y_transform, self.in_size,
                                                           self.out_size)
        sd = {k: v.detach().cpu().numpy() for k, v in self.model.state_dict().items()}
        sg = None if sigma is None else sigma.detach().cpu().numpy().astype(np.float32).reshape(-1)
        return _engine.Engine(self.model.KIND, self.in_size, self.out_size, sd, xm, xs, ym, ys, dolog10index=log10,
                              ypositive=ypos, sigma=sg, device=device_index)

    def set_output_scale(self, sigma):
        """Remember sigma so that ``predict_data_vector`` can return m = y*sigma from the same launch."""
        self._sigma = sigma
        self._engine = None

    def _get_engine(self):
        if not torch.cuda.is_available():
            raise RuntimeError("Predictor: no CUDA device -- linna_b200 has no CPU fallback")
        dev = torch.cuda.current_device()
        key = (dev, self.param_key())
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = self.make_engine(dev, sigma=self._sigma)
            self._engine_key = key
        return self._engine

    def _run(self, X, out_kind):
        eng = self._get_engine()
        was_np = isinstance(X, np.ndarray)
        Xt = torch.from_numpy(np.asarray(X, np.float32)) if was_np else X
        one_input = Xt.dim() == 1
        X2 = Xt.reshape(1, -1) if one_input else Xt
        if X2.is_cuda:
            y = eng.predict(X2, out_kind)
        else:
            y = torch.from_numpy(eng.predict(X2.detach().cpu().numpy(), out_kind))
        return y.view(-1) if one_input else y

    # ---- reference API ------------------------------------------------------------------
    def predict(self, X, no_grad=True):
        """y_transform(model(X_transform(X))); 1-D in -> 1-D out (linna/predictor_gpu.py:461-504)."""
        from . import engine as _engine
        self.model.eval()
        if not no_grad and torch.is_tensor(X) and X.requires_grad:
            from .train import predict_with_grad
            return predict_with_grad(self, X)
        return self._run(X, _engine.LINNA_OUT_Y)

    def predict_data_vector(self, X, no_grad=True):
        """y_invtransform_data(predict(X)) = data-space model vector (needs ``set_output_scale``)."""
        from . import engine as _engine
        if self._sigma is None:
            raise RuntimeError("call set_output_scale(sigma) first")
        return self._run(X, _engine.LINNA_OUT_M)

    def load_checkpoint(self, ismpi=False):
        path = os.path.join(self.outdir, "best.pth.tar")
        if not os.path.isfile(path):
            return False
        optim = self.optim if isinstance(self.optim, torch.optim.Optimizer) else None
        load_checkpoint(path, self.model, optim, device=self.device, ismpi=ismpi)
        return True

    def train(self, dataset, num_epochs, loss_fn, val_dataset=None, val_metric_fn=None, initfrombest=False,
              pool=None, nocpu=False, rank=0, size=1):
        """Training loop of linna/predictor_gpu.py:201-449 around the fused device step."""
        from .trainer import run_training
        return run_training(self, dataset, num_epochs, loss_fn, val_dataset, val_metric_fn, initfrombest, pool,
                            nocpu, rank, size)

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engine"], d["_engine_key"] = None, None
        return d


for _c in (EarlyStopping, Predictor):
    _c.__module__ = "linna.predictor_gpu"
del _c
