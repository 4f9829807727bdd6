"""Emulator network classes with the reference's names, constructor signatures, parameter
names (``state_dict`` keys) and initialisation -- ``linna/nn.py`` in the reference:
``ResBlock_batchnorm`` (:11-56), ``ChtoModelv2`` (:59-133), ``ChtoModelv2_linear`` (:136-198),
``ChtoModelsimple`` (:300-374).

The modules are *parameter containers*: they exist so that reference checkpoints
(``best.pth.tar``) load unchanged, ``torch.optim`` state round-trips and the pickled
``model_args.pkl`` (which names the class) resolves.  ``forward`` does not run torch math: it
hands the batch to the fused sm_100a kernel through the C ABI (``engine.Engine``) and fails
loudly on a CPU tensor -- there is no fallback path.
"""
import numpy as np
import torch
from torch import nn

from . import arch


def _reset_linear(lin):
    """Xavier-uniform weight, bias 1e-2 (reference ``init_weight``, linna/nn.py:38-42, :96-99)."""
    nn.init.xavier_uniform_(lin.weight)
    if lin.bias is not None:
        lin.bias.data.fill_(1e-2)


class ResBlock_batchnorm(nn.Module):
    """h = relu(layer1 x); y = relu(0.1*layer2 h + skip x).  No batch-norm despite the name."""

    def __init__(self, in_size, channel, out_size):
        super().__init__()
        self.layer1 = nn.Linear(in_size, channel)
        self.layer2 = nn.Linear(channel, out_size)
        self.skip_layer = nn.Identity() if in_size == out_size else nn.Linear(in_size, out_size, bias=False)
        self.init_weight()

    def init_weight(self):
        for mod in self.modules():
            if type(mod) == nn.Linear:
                _reset_linear(mod)
        if isinstance(self.skip_layer, nn.Linear):   # the reference would crash on Identity
            nn.init.zeros_(self.skip_layer.weight)

    def forward(self, x):
        raise RuntimeError("ResBlock_batchnorm is evaluated inside the fused emulator kernel; call the "
                           "enclosing model")


class _ChtoBase(nn.Module):
    KIND = None

    def __init__(self, in_size, out_size, linearmodel, docpu=False):
        super().__init__()
        self.in_size, self.out_size = int(in_size), int(out_size)
        self.channel = 4 if self.KIND == "ChtoModelsimple" else 16
        for op in arch.chto_ops(self.KIND, self.in_size, self.out_size):
            if op["kind"] == "linear":
                setattr(self, op["name"], nn.Linear(op["in"], op["out"]))
            else:
                setattr(self, op["name"], ResBlock_batchnorm(op["in"], op["mid"], op["out"]))
        if self.KIND == "ChtoModelv2_linear":
            self.linearlayer = nn.Linear(self.in_size, self.out_size)
        self.init_weight()
        if self.KIND == "ChtoModelv2_linear":
            self.linearlayer.bias.data.fill_(0)
            self.linearlayer.weight.data.fill_(1e-5)
        if linearmodel is not None:
            raise NotImplementedError("linearmodel add-on is disabled in the reference (util.py:634) and "
                                      "unsupported here")
        self.linearmodel = None
        self.docpu = docpu
        self._engine = None
        self._engine_key = None

    def init_weight(self):
        """Same traversal as the reference (linna/nn.py:91-108): every Linear is Xavier'd, and each
        res-block re-initialises itself when visited, so skip weights end up Xavier as well."""
        for mod in self.modules():
            if type(mod) == nn.Linear:
                _reset_linear(mod)
            elif type(mod) == ResBlock_batchnorm:
                mod.init_weight()
            elif mod is self or isinstance(mod, (nn.Identity, nn.modules.batchnorm.BatchNorm1d)):
                pass    # (the reference asserts on Identity; unreachable for its width tables)
            else:
                print(type(mod), flush=True)
                assert 0

    # ---- bridge to the CUDA engine ------------------------------------------------------
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def bare_engine(self, device_index):
        """Engine with identity input/output transforms (yhat = model(xhat)); rebuilt when any
        parameter has been modified in place."""
        from . import engine as _engine
        key = (device_index, self._param_key())
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            sd = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}
            ones, zeros = np.ones, np.zeros
            self._engine = _engine.Engine(self.KIND, self.in_size, self.out_size, sd,
                                          zeros(self.in_size, np.float32), ones(self.in_size, np.float32),
                                          zeros(self.out_size, np.float32), ones(self.out_size, np.float32),
                                          device=device_index)
            self._engine_key = key
        return self._engine

    def forward(self, s):
        if not s.is_cuda:
            raise RuntimeError("%s.forward: input is on %s; linna_b200 evaluates emulators on the GPU only "
                               "(no CPU fallback)" % (self.KIND, s.device))
        from . import engine as _engine
        squeeze = s.dim() == 1
        x = s.reshape(1, -1) if squeeze else s
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .train import emulator_forward_autograd
            y = emulator_forward_autograd(self, x)
        else:
            y = self.bare_engine(x.device.index).predict(x, _engine.LINNA_OUT_YHAT)
        return y.reshape(-1) if squeeze else y

    def __getstate__(self):
        d = self.__dict__.copy()
        d["_engine"], d["_engine_key"] = None, None
        return d


class ChtoModelv2(_ChtoBase):
    """Default LINNA emulator (linna/main.py:70, yamlfile/training_3x2pt.yaml:38)."""
    KIND = "ChtoModelv2"


class ChtoModelv2_linear(_ChtoBase):
    """ChtoModelv2 + 1e-3 * Linear(xhat) (linna/nn.py:193)."""
    KIND = "ChtoModelv2_linear"


class ChtoModelsimple(_ChtoBase):
    """Narrow variant: channel 4, layer6 h/8 -> h/8 (linna/nn.py:300-374)."""
    KIND = "ChtoModelsimple"


for _c in (ResBlock_batchnorm, ChtoModelv2, ChtoModelv2_linear, ChtoModelsimple):
    _c.__module__ = "linna.nn"   # pickles written here resolve under the reference's module path
del _c
