"""ctypes binding of the linna_b200 C ABI (``include/linna_b200.h``) and the ``Engine``
object the LINNA-facing classes (``Predictor``, ``Log_prob``, ``HMCSampler``) sit on.

PyTorch is used for device memory and streams only; every number on the hot path is
produced by the hand-written sm_100a kernels in ``csrc/``.  There is no CPU path: if
the shared library is missing or no B200 is visible, construction raises.
"""
import ctypes
import os

import numpy as np

from . import arch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "liblinna_b200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_i32_p = ctypes.POINTER(ctypes.c_int32)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)

LINNA_OUT_YHAT, LINNA_OUT_Y, LINNA_OUT_M = 0, 1, 2
LINNA_QUAD_CHOL, LINNA_QUAD_DENSE = 0, 1
_PRIOR = {"gauss": 0, "flat": 1}


class OpDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("in_dim", ctypes.c_int32), ("mid_dim", ctypes.c_int32),
                ("out_dim", ctypes.c_int32), ("act", ctypes.c_int32), ("alpha", ctypes.c_float),
                ("w", c_float_p), ("b", c_float_p), ("w2", c_float_p), ("b2", c_float_p), ("ws", c_float_p)]


class ModelDesc(ctypes.Structure):
    _fields_ = [("n_in", ctypes.c_int32), ("n_out", ctypes.c_int32), ("n_ops", ctypes.c_int32),
                ("ops", ctypes.POINTER(OpDesc)),
                ("x_mean", c_float_p), ("x_std", c_float_p), ("log10_flag", c_u8_p),
                ("y_mean", c_float_p), ("y_std", c_float_p), ("ypositive", ctypes.c_int32),
                ("sigma", c_float_p), ("extra_linear_w", c_float_p), ("extra_linear_b", c_float_p),
                ("extra_linear_scale", ctypes.c_float)]


class LikeDesc(ctypes.Structure):
    _fields_ = [("prior_kind", c_i32_p), ("prior_arg1", c_float_p), ("prior_arg2", c_float_p),
                ("data", c_float_p), ("quad", c_float_p), ("quad_kind", ctypes.c_int32),
                ("temperature", ctypes.c_float)]


class TrainDesc(ctypes.Structure):
    _fields_ = [("data_hat", c_float_p), ("icov_hat", c_float_p), ("max_batch", ctypes.c_int32)]


class LinnaError(RuntimeError):
    pass


_lib = None


def load_library():
    """Load ``liblinna_b200.so`` (built in-tree by ``__graft_entry__.build()``).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LinnaError("linna_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                         "there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.linna_abi_version.restype = ctypes.c_int
    lib.linna_last_error.restype = ctypes.c_char_p
    lib.linna_launch_count.restype = ctypes.c_int64
    lib.linna_model_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.c_int, ctypes.POINTER(vp)]
    lib.linna_model_destroy.argtypes = [vp]
    lib.linna_model_destroy.restype = None
    lib.linna_model_set_likelihood.argtypes = [vp, ctypes.POINTER(LikeDesc)]
    lib.linna_model_set_weights.argtypes = [vp, ctypes.POINTER(OpDesc), i32, c_float_p, c_float_p]
    lib.linna_predict.argtypes = [vp, vp, i64, vp, i32, vp]
    lib.linna_lnp.argtypes = [vp, vp, i64, vp, vp]
    lib.linna_lnp_grad.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.linna_predict_vjp.argtypes = [vp, vp, i64, vp, i32, vp, vp, vp, vp]
    lib.linna_predict_host.argtypes = [vp, vp, i64, vp, i32]
    lib.linna_lnp_host.argtypes = [vp, vp, i64, vp]
    lib.linna_lnp_grad_host.argtypes = [vp, vp, i64, vp, vp]
    lib.linna_model_info.argtypes = [vp, c_i32_p, c_i32_p, ctypes.POINTER(i64), c_i32_p]
    lib.linna_model_set_tile_rows.argtypes = [vp, i32]
    lib.linna_model_set_path.argtypes = [vp, i32, i64]
    lib.linna_model_set_fold.argtypes = [vp, i32]
    lib.linna_debug_tc_counters.argtypes = [vp, vp, i32]
    lib.linna_model_last_kernel.argtypes = [vp]
    lib.linna_debug_tg_counters.argtypes = [vp, vp, i32]
    lib.linna_debug_cluster_counters.argtypes = [vp, i32]
    lib.linna_column_median_mad.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp]
    u64 = ctypes.c_uint64
    lib.linna_stretch_propose.argtypes = [vp, i32, vp, vp, i64, i64, ctypes.c_float, u64, u64, vp, vp, vp]
    lib.linna_stretch_accept.argtypes = [vp, vp, vp, i32, vp, i64, vp, vp, vp, u64, u64, vp]
    f32 = ctypes.c_float
    lib.linna_hmc_begin.argtypes = [vp, vp, vp, vp, i32, i64, f32, u64, u64, vp, vp, vp, vp]
    lib.linna_hmc_step.argtypes = [vp, vp, vp, vp, i32, i64, f32, vp]
    lib.linna_hmc_end.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i64, f32, u64, u64, vp, vp]
    lib.linna_column_moments.argtypes = [vp, i32, i64, i64, i32, vp, vp, vp]
    lib.linna_train_setup.argtypes = [vp, ctypes.POINTER(TrainDesc)]
    lib.linna_train_num_params.argtypes = [vp]
    lib.linna_train_num_params.restype = ctypes.c_int64
    lib.linna_train_chisq.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.linna_train_step.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp, vp, vp]
    lib.linna_train_adamw.argtypes = [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp]
    lib.linna_train_load_params.argtypes = [vp, vp, vp]
    lib.linna_train_commit.argtypes = [vp, vp]
    lib.linna_loss_terms.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.linna_train_set_path.argtypes = [vp, i32]
    lib.linna_train_adamw_peer.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp, i32, i32, i32, i64, f32, f32, f32, f32, f32, vp]
    lib.linna_train_last_kernel.argtypes = [vp]
    if lib.linna_abi_version() != 1:
        raise LinnaError("linna_b200: ABI version mismatch")
    _lib = lib
    return lib


def stretch_propose(x, first, second, a, seed, offset):
    """Stretch-move proposal for the walkers ``first`` against the complementary set ``second`` (device tensors:
    x [W, d] float32, index tensors int64).  Returns (y [ns, d], z [ns]).  linna/sampler.py:493-503 via emcee."""
    import torch
    lib = load_library()
    ns, d = int(first.numel()), int(x.shape[1])
    y = torch.empty((ns, d), dtype=torch.float32, device=x.device)
    z = torch.empty(ns, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.linna_stretch_propose(x.data_ptr(), d, first.data_ptr(), second.data_ptr(), ns, int(second.numel()), float(a),
                                       int(seed), int(offset), y.data_ptr(), z.data_ptr(),
                                       ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise LinnaError("linna_stretch_propose failed (%d)" % rc)
    return y, z


def stretch_accept(x, lnp, naccepted, first, y, lnp_y, z, seed, offset):
    """Accept / reject the proposals in place (x, lnp, naccepted are updated for the accepted walkers)."""
    import torch
    lib = load_library()
    with torch.cuda.device(x.device):
        rc = lib.linna_stretch_accept(x.data_ptr(), lnp.data_ptr(), naccepted.data_ptr(), int(x.shape[1]), first.data_ptr(),
                                      int(first.numel()), y.data_ptr(), lnp_y.data_ptr(), z.data_ptr(), int(seed), int(offset),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise LinnaError("linna_stretch_accept failed (%d)" % rc)


def _cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def hmc_begin(x, lnp, grad, mass, eps, seed, offset):
    """Momentum draw, initial Hamiltonian, first half kick and drift of every chain (linna/HMCSampler.py:25-36).
    Returns (p, xn, H0) device tensors."""
    import torch
    lib = load_library()
    nc, d = int(x.shape[0]), int(x.shape[1])
    p, xn = torch.empty_like(x), torch.empty_like(x)
    H0 = torch.empty(nc, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.linna_hmc_begin(x.data_ptr(), lnp.data_ptr(), grad.data_ptr(), mass.data_ptr(), d, nc, float(eps), int(seed),
                                 int(offset), p.data_ptr(), xn.data_ptr(), H0.data_ptr(), _cur_stream())
    if rc != 0:
        raise LinnaError("linna_hmc_begin failed (%d)" % rc)
    return p, xn, H0


def hmc_step(p, xn, grad_n, mass, eps):
    """Inner leapfrog step in place: p += eps grad(xn); xn += eps p / m (linna/HMCSampler.py:43-46)."""
    import torch
    lib = load_library()
    with torch.cuda.device(p.device):
        rc = lib.linna_hmc_step(p.data_ptr(), xn.data_ptr(), grad_n.data_ptr(), mass.data_ptr(), int(p.shape[1]), int(p.shape[0]),
                                float(eps), _cur_stream())
    if rc != 0:
        raise LinnaError("linna_hmc_step failed (%d)" % rc)


def hmc_end(x, lnp, grad, xn, lnp_n, grad_n, p, mass, H0, eps, seed, offset, naccepted):
    """Last half kick, Hamiltonian and the per-chain Metropolis select, in place (linna/HMCSampler.py:51-59)."""
    import torch
    lib = load_library()
    with torch.cuda.device(x.device):
        rc = lib.linna_hmc_end(x.data_ptr(), lnp.data_ptr(), grad.data_ptr(), xn.data_ptr(), lnp_n.data_ptr(), grad_n.data_ptr(),
                               p.data_ptr(), mass.data_ptr(), H0.data_ptr(), int(x.shape[1]), int(x.shape[0]), float(eps), int(seed),
                               int(offset), naccepted.data_ptr(), _cur_stream())
    if rc != 0:
        raise LinnaError("linna_hmc_end failed (%d)" % rc)


def column_moments(x2d, r0, r1):
    """(mean, std) per column of rows [r0, r1) of a CUDA matrix (float32 / float64), float64 numpy arrays."""
    import torch
    lib = load_library()
    assert x2d.is_cuda and x2d.dim() == 2 and x2d.is_contiguous() and x2d.dtype in (torch.float32, torch.float64)
    d = int(x2d.shape[1])
    mean, std = np.empty(d, np.float64), np.empty(d, np.float64)
    with torch.cuda.device(x2d.device):
        rc = lib.linna_column_moments(x2d.data_ptr(), int(x2d.dtype == torch.float64), int(r0), int(r1), d, mean.ctypes.data,
                                      std.ctypes.data, _cur_stream())
    if rc != 0:
        raise LinnaError("linna_column_moments failed (%d)" % rc)
    return mean, std


def loss_terms(y_pred, y_target, data_hat, icov_hat, sigma, y_mean, y_std, ypositive, want_grad=False):
    """(loss, chisqMd, chisqnnd[, d loss/d y_pred]) per row: Auxilleryfunc.__call__ (linna/util.py:1070-1088) on CUDA
    tensors, one kernel launch."""
    import torch
    lib = load_library()
    n, n_out = int(y_pred.shape[0]), int(y_pred.shape[1])
    dev = y_pred.device
    loss = torch.empty(n, dtype=torch.float32, device=dev)
    md, nnd = torch.empty_like(loss), torch.empty_like(loss)
    g = torch.empty_like(y_pred) if want_grad else None
    with torch.cuda.device(dev):
        rc = lib.linna_loss_terms(y_pred.data_ptr(), y_target.data_ptr(), n, n_out, data_hat.data_ptr(), icov_hat.data_ptr(),
                                  sigma.data_ptr() if sigma is not None else None, y_mean.data_ptr(), y_std.data_ptr(),
                                  int(bool(ypositive)), loss.data_ptr(), md.data_ptr(), nnd.data_ptr(),
                                  g.data_ptr() if g is not None else None, _cur_stream())
    if rc != 0:
        raise LinnaError("linna_loss_terms failed (%d)" % rc)
    return loss, md, nnd, g


def column_median_mad(Y, sigma=None, take_log=False):
    """(median, MAD) per column of the CUDA tensor Y [n, d] after v = Y / sigma (log(v) when take_log): the training-set
    statistics of train_NN (linna/util.py:1440-1450) by radix selection on the device; lower medians, as torch.median."""
    import torch
    lib = load_library()
    Y = Y.contiguous()
    n, d = int(Y.shape[0]), int(Y.shape[1])
    med = torch.empty(d, dtype=torch.float32, device=Y.device)
    mad = torch.empty_like(med)
    sg = None if sigma is None else torch.as_tensor(sigma, dtype=torch.float32, device=Y.device).contiguous()
    with torch.cuda.device(Y.device):
        rc = lib.linna_column_median_mad(Y.data_ptr(), n, d, sg.data_ptr() if sg is not None else None, int(bool(take_log)),
                                         med.data_ptr(), mad.data_ptr(), _cur_stream())
    if rc != 0:
        raise LinnaError("linna_column_median_mad failed (%d)" % rc)
    return med, mad


def cluster_counters():
    """Per-step cycle counters of the small-batch cluster kernel since the last read (needs LINNA_CLUSTER_DEBUG=1):
    array [64, 4] = k-loop, k-lane reduction, epilogue + broadcast, cluster barrier."""
    buf = np.zeros((128, 4), np.int64)
    load_library().linna_debug_cluster_counters(buf.ctypes.data_as(ctypes.c_void_p), 512)
    return buf


def launch_count():
    return int(load_library().linna_launch_count())


def _f32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def build_op_descs(kind, n_in, n_out, state_dict):
    """(ctypes OpDesc array, keep-alive list, extra-linear (w, b) or None) from a reference-layout
    ``state_dict`` (keys of SURVEY 8b)."""
    ops = arch.chto_ops(kind, n_in, n_out)
    arr = (OpDesc * len(ops))()
    keep = []

    def g(key):
        a = _f32(state_dict[key])
        keep.append(a)
        return a

    for i, op in enumerate(ops):
        d = arr[i]
        nm = op["name"]
        d.in_dim, d.out_dim = op["in"], op["out"]
        if op["kind"] == "linear":
            d.kind, d.mid_dim, d.alpha = 0, 0, 1.0
            d.act = 1 if op["act"] == "relu" else 0
            w, b = g(nm + ".weight"), g(nm + ".bias")
            assert w.shape == (op["out"], op["in"]) and b.shape == (op["out"],), nm
            d.w, d.b = _fp(w), _fp(b)
        else:
            d.kind, d.mid_dim, d.alpha, d.act = 1, op["mid"], op["alpha"], 1
            w, b = g(nm + ".layer1.weight"), g(nm + ".layer1.bias")
            w2, b2 = g(nm + ".layer2.weight"), g(nm + ".layer2.bias")
            assert w.shape == (op["mid"], op["in"]) and w2.shape == (op["out"], op["mid"]), nm
            d.w, d.b, d.w2, d.b2 = _fp(w), _fp(b), _fp(w2), _fp(b2)
            if op["in"] != op["out"]:
                ws = g(nm + ".skip_layer.weight")
                assert ws.shape == (op["out"], op["in"]), nm
                d.ws = _fp(ws)
    extra = None
    if kind == "ChtoModelv2_linear":
        extra = (g("linearlayer.weight"), g("linearlayer.bias"))
    return arr, keep, extra


def cholesky_of_inverse(inv_cov):
    """L (lower, float64) with inv_cov = L L^T, or None when the matrix is not positive definite."""
    a = np.asarray(inv_cov, np.float64)
    a = 0.5 * (a + a.T)
    try:
        return np.linalg.cholesky(a)
    except np.linalg.LinAlgError:
        return None


class Engine:
    """One packed emulator (+ optionally its likelihood constants) on one GPU."""

    def __init__(self, kind, n_in, n_out, state_dict, X_mean, X_std, y_mean, y_std, dolog10index=None,
                 ypositive=False, sigma=None, device=0):
        self.lib = load_library()
        self.kind, self.n_in, self.n_out = str(kind), int(n_in), int(n_out)
        self.device = int(device)
        self.handle = ctypes.c_void_p()
        self._has_like = False
        arr, keep, extra = build_op_descs(self.kind, self.n_in, self.n_out, state_dict)
        d = ModelDesc()
        d.n_in, d.n_out, d.n_ops, d.ops = self.n_in, self.n_out, len(arr), arr
        xm, xs, ym, ys = _f32(X_mean), _f32(X_std), _f32(y_mean), _f32(y_std)
        assert xm.shape == (n_in,) and xs.shape == (n_in,) and ym.shape == (n_out,) and ys.shape == (n_out,)
        d.x_mean, d.x_std, d.y_mean, d.y_std = _fp(xm), _fp(xs), _fp(ym), _fp(ys)
        flags = np.zeros(n_in, np.uint8)
        if dolog10index is not None:
            flags[list(dolog10index)] = 1
        d.log10_flag = flags.ctypes.data_as(c_u8_p)
        d.ypositive = int(bool(ypositive))
        sg = _f32(sigma) if sigma is not None else None
        d.sigma = _fp(sg)
        if extra is not None:
            d.extra_linear_w, d.extra_linear_b, d.extra_linear_scale = _fp(extra[0]), _fp(extra[1]), 1e-3
        rc = self.lib.linna_model_create(ctypes.byref(d), self.device, ctypes.byref(self.handle))
        self._check(rc)
        del keep

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise LinnaError("linna_b200 error %d: %s" % (rc, self.lib.linna_last_error().decode()))

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.linna_model_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tile_rows(self, rows):
        self._check(self.lib.linna_model_set_tile_rows(self.handle, int(rows)))

    def set_path(self, path, tc_min_rows=0):
        """'auto' | 'ffma' | 'tc' | 'cluster' -- which kernel serves lnp() / lnp_grad() / predict()."""
        code = {"auto": 0, "ffma": 1, "tc": 2, "cluster": 3}[path]
        self._check(self.lib.linna_model_set_path(self.handle, code, int(tc_min_rows)))

    def last_kernel(self):
        """'ffma' | 'tc' | 'cluster' | None -- the kernel that served the last launch."""
        return {0: None, 1: "ffma", 2: "tc", 3: "cluster"}[int(self.lib.linna_model_last_kernel(self.handle))]

    def tc_counters(self, max_ctas=256):
        """Per-CTA cycle counters of the last tensor-core launch (needs LINNA_TC_DEBUG=1 in the environment)."""
        buf = np.zeros((max_ctas, 128), np.int64)
        n = self.lib.linna_debug_tc_counters(self.handle, buf.ctypes.data_as(ctypes.c_void_p), max_ctas)
        return buf[:n]

    def tg_counters(self, max_steps=64):
        """clock64 stamps of CTA 0 of every layer launch of the last training step (needs LINNA_TG_DEBUG=1)."""
        buf = np.zeros((max_steps, 8), np.int64)
        n = self.lib.linna_debug_tg_counters(self.handle, buf.ctypes.data_as(ctypes.c_void_p), max_steps)
        return buf[:n]

    def set_fold(self, on):
        """Fold last layer + inverse transform + Cholesky product into one GEMM for lnP (default on)."""
        self._check(self.lib.linna_model_set_fold(self.handle, int(bool(on))))

    def set_weights(self, state_dict):
        arr, keep, extra = build_op_descs(self.kind, self.n_in, self.n_out, state_dict)
        ew, eb = (_fp(extra[0]), _fp(extra[1])) if extra is not None else (None, None)
        self._check(self.lib.linna_model_set_weights(self.handle, arr, len(arr), ew, eb))

    def set_likelihood(self, priors, data, inv_cov, temperature=1.0, quad="chol"):
        """priors: list of {'dist','arg1','arg2'} (README.rst:73-80); inv_cov: float64 C^-1 as
        ``np.linalg.inv(cov)`` (linna/main.py:120).  quad='chol' factors it in float64 (chi^2 =
        |L^T d|^2); quad='dense' uses the f32 matrix itself like the reference (util.py:953-955)."""
        n_in, n_out = self.n_in, self.n_out
        if len(priors) != n_in:
            raise ValueError("need %d priors, got %d" % (n_in, len(priors)))
        kinds = []
        for p in priors:
            if p["dist"] not in _PRIOR:
                print("not implement dist : {0}".format(p["dist"]), flush=True)   # linna/main.py:128-129
                assert 0
            kinds.append(_PRIOR[p["dist"]])
        pk = np.asarray(kinds, np.int32)
        a1 = _f32([p["arg1"] for p in priors])
        a2 = _f32([p["arg2"] for p in priors])
        dat = _f32(data)
        inv_cov = np.asarray(inv_cov)
        assert dat.shape == (n_out,) and inv_cov.shape == (n_out, n_out)
        kind = LINNA_QUAD_DENSE
        q = None
        if quad == "chol":
            L = cholesky_of_inverse(inv_cov)
            if L is not None:
                q, kind = _f32(L), LINNA_QUAD_CHOL
        if q is None:
            q = _f32(inv_cov)
        ld = LikeDesc()
        ld.prior_kind = pk.ctypes.data_as(c_i32_p)
        ld.prior_arg1, ld.prior_arg2, ld.data, ld.quad = _fp(a1), _fp(a2), _fp(dat), _fp(q)
        ld.quad_kind, ld.temperature = kind, float(temperature)
        self._check(self.lib.linna_model_set_likelihood(self.handle, ctypes.byref(ld)))
        self._has_like = True
        self.quad_kind = "chol" if kind == LINNA_QUAD_CHOL else "dense"
        self.temperature = float(temperature)

    # -- tensor helpers ------------------------------------------------------------------
    @staticmethod
    def _stream():
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _prep_dev(self, x, width):
        import torch
        if x.device.type != "cuda" or x.device.index != self.device:
            raise LinnaError("tensor must live on cuda:%d" % self.device)
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != width:
            raise ValueError("expected shape [n, %d], got %s" % (width, tuple(x.shape)))
        return x

    # -- the three calls ------------------------------------------------------------------
    def predict(self, theta, out_kind=LINNA_OUT_Y):
        """Batched Predictor.predict.  torch CUDA tensor in -> torch CUDA tensor out;
        numpy in -> numpy out through the host-buffer entry point."""
        import torch
        if isinstance(theta, np.ndarray):
            th = np.ascontiguousarray(theta, np.float32).reshape(-1, self.n_in)
            out = np.empty((th.shape[0], self.n_out), np.float32)
            self._check(self.lib.linna_predict_host(self.handle, th.ctypes.data, th.shape[0], out.ctypes.data, out_kind))
            return out
        th = self._prep_dev(theta, self.n_in)
        out = torch.empty((th.shape[0], self.n_out), dtype=torch.float32, device=th.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_predict(self.handle, th.data_ptr(), th.shape[0], out.data_ptr(), out_kind,
                                               self._stream()))
        return out

    def predict_vjp(self, theta, cot, out_kind=LINNA_OUT_Y, want_params=False):
        """Vector-Jacobian product of ``predict`` for CUDA tensors: returns (d/d theta [n, n_in], flat parameter gradient
        or None).  ``want_params`` needs ``train_setup`` (max_batch >= n)."""
        import torch
        th = self._prep_dev(theta, self.n_in)
        ct = self._prep_dev(cot, self.n_out)
        gth = torch.empty_like(th)
        gp = torch.zeros(self.n_params, dtype=torch.float32, device=th.device) if want_params else None
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_predict_vjp(self.handle, th.data_ptr(), th.shape[0], ct.data_ptr(), int(out_kind), None,
                                                   gth.data_ptr(), gp.data_ptr() if gp is not None else None, self._stream()))
        return gth, gp

    @staticmethod
    def _host_out(out, shape):
        """Caller-provided result buffer (e.g. pinned memory, which skips the staging copy) or a fresh array."""
        if out is None:
            return np.empty(shape, np.float32)
        if out.dtype != np.float32 or out.shape != tuple(shape) or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % (tuple(shape),))
        return out

    def lnp(self, u, out=None):
        import torch
        if isinstance(u, np.ndarray):
            uu = np.ascontiguousarray(u, np.float32).reshape(-1, self.n_in)
            out = self._host_out(out, (uu.shape[0],))
            self._check(self.lib.linna_lnp_host(self.handle, uu.ctypes.data, uu.shape[0], out.ctypes.data))
            return out
        uu = self._prep_dev(u, self.n_in)
        out = torch.empty(uu.shape[0], dtype=torch.float32, device=uu.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_lnp(self.handle, uu.data_ptr(), uu.shape[0], out.data_ptr(), self._stream()))
        return out

    def lnp_grad(self, u, out=None, out_grad=None):
        import torch
        if isinstance(u, np.ndarray):
            uu = np.ascontiguousarray(u, np.float32).reshape(-1, self.n_in)
            out = self._host_out(out, (uu.shape[0],))
            g = self._host_out(out_grad, uu.shape)
            self._check(self.lib.linna_lnp_grad_host(self.handle, uu.ctypes.data, uu.shape[0], out.ctypes.data,
                                                     g.ctypes.data))
            return out, g
        uu = self._prep_dev(u, self.n_in)
        out = torch.empty(uu.shape[0], dtype=torch.float32, device=uu.device)
        g = torch.empty_like(uu)
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_lnp_grad(self.handle, uu.data_ptr(), uu.shape[0], out.data_ptr(), g.data_ptr(),
                                                self._stream()))
        return out, g

    # -- training ---------------------------------------------------------------------------
    def train_setup(self, data_hat, icov_hat, max_batch):
        """Constants of the training loss (Auxilleryfunc.__init__, linna/util.py:1060-1069)."""
        dh, ic = _f32(data_hat), _f32(icov_hat)
        assert dh.shape == (self.n_out,) and ic.shape == (self.n_out, self.n_out)
        d = TrainDesc()
        d.data_hat, d.icov_hat, d.max_batch = _fp(dh), _fp(ic), int(max_batch)
        self._check(self.lib.linna_train_setup(self.handle, ctypes.byref(d)))
        self.n_params = int(self.lib.linna_train_num_params(self.handle))

    def set_train_path(self, path):
        """'auto' | 'ffma' | 'tc' -- which kernels run train_step / train_chisq."""
        self._check(self.lib.linna_train_set_path(self.handle, {"auto": 0, "ffma": 1, "tc": 2}[path]))

    def last_train_kernel(self):
        return {0: None, 1: "ffma", 2: "tc"}[int(self.lib.linna_train_last_kernel(self.handle))]

    def train_chisq(self, X, Y, kind):
        """Per-row chi^2 of the loss (kind 0: target/pred, 1: target/data, 2: pred/data); CUDA tensors."""
        import torch
        X, Y = self._prep_dev(X, self.n_in), self._prep_dev(Y, self.n_out)
        out = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_train_chisq(self.handle, X.data_ptr(), Y.data_ptr(), X.shape[0], int(kind),
                                                   out.data_ptr(), self._stream()))
        return out

    def train_step(self, X, Y, cmd, params, adam_m, adam_v, grads, step, lr, betas=(0.9, 0.999), eps=1e-8,
                   weight_decay=1e-4, fuse_adam=True, loss_rows=None, loss_mean=None):
        import torch
        X, Y = self._prep_dev(X, self.n_in), self._prep_dev(Y, self.n_out)
        B = X.shape[0]
        if loss_rows is None:
            loss_rows = torch.empty(B, dtype=torch.float32, device=X.device)
        if loss_mean is None:
            loss_mean = torch.empty(1, dtype=torch.float32, device=X.device)
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_train_step(self.handle, X.data_ptr(), Y.data_ptr(), cmd.data_ptr(), B, ptr(params),
                                                  ptr(adam_m), ptr(adam_v), ptr(grads), int(step), float(lr),
                                                  float(betas[0]), float(betas[1]), float(eps), float(weight_decay),
                                                  int(bool(fuse_adam)), loss_rows.data_ptr(), loss_mean.data_ptr(),
                                                  self._stream()))
        return loss_mean, loss_rows

    def train_adamw_peer(self, params, adam_m, adam_v, peer_grad_ptrs, grad_offset, avg_offset, signal_pad_ptrs, signal_slot, world,
                         rank, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        """AdamW on the average of the ranks' gradients, read from the peers' memory inside the kernel (linna_train_adamw_peer)."""
        import torch
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_train_adamw_peer(self.handle, params.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr(),
                                                        int(peer_grad_ptrs), int(grad_offset), int(avg_offset), int(signal_pad_ptrs), int(signal_slot),
                                                        int(world), int(rank), int(step), float(lr), float(betas[0]), float(betas[1]),
                                                        float(eps), float(weight_decay), self._stream()))

    def train_adamw(self, params, adam_m, adam_v, grads, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        import torch
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_train_adamw(self.handle, params.data_ptr(), adam_m.data_ptr(), adam_v.data_ptr(),
                                                   grads.data_ptr(), int(step), float(lr), float(betas[0]),
                                                   float(betas[1]), float(eps), float(weight_decay), self._stream()))

    def train_load_params(self, params):
        import torch
        with torch.cuda.device(self.device):
            self._check(self.lib.linna_train_load_params(self.handle, params.data_ptr(), self._stream()))

    def train_commit(self, params_host):
        p = _f32(params_host)
        self._check(self.lib.linna_train_commit(self.handle, p.ctypes.data))

    def info(self):
        n_in, n_out, sms = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        npar = ctypes.c_int64()
        self._check(self.lib.linna_model_info(self.handle, ctypes.byref(n_in), ctypes.byref(n_out),
                                              ctypes.byref(npar), ctypes.byref(sms)))
        return dict(n_in=n_in.value, n_out=n_out.value, n_params=npar.value, num_sms=sms.value)


def engine_from_problem(p, device=0, quad="chol", with_likelihood=True):
    """Engine for a ``linna_b200.synthetic.Problem`` (tests, bench)."""
    e = Engine(p.kind, p.n_in, p.n_out, p.state_dict, p.X_mean, p.X_std, p.y_mean, p.y_std,
               dolog10index=p.dolog10index, ypositive=p.ypositive, sigma=np.asarray(p.sigma, np.float32),
               device=device)
    if with_likelihood and p.data is not None:
        e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, p.temperature, quad=quad)
    return e
