"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

``Oracle`` wraps ``oracle/_build/liblinna_oracle.so`` (plain C restatement of the
reference, see ``linna_oracle_impl.h`` for per-function reference citations) through
ctypes.  ``NumpyPort`` is a batched numpy/BLAS restatement of the same arithmetic used
only as the timed CPU baseline in ``bench.py`` (``cpu_baseline.kind == "port"``).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline and
--impl reference).  The product package ``linna_b200`` never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liblinna_oracle.so")

_PRIOR_KIND = {"gauss": 0, "flat": 1}


def build(force=False):
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB)
            for f in ("linna_oracle.c", "linna_oracle_impl.h")):
        subprocess.check_call(["make", "-C", HERE, "-s", "clean", "all"])
    return LIB


def _ops_from_arch(ops):
    rows, alpha = [], []
    for op in ops:
        if op["kind"] == "linear":
            rows.append([0, op["in"], 0, op["out"], 1 if op["act"] == "relu" else 0])
            alpha.append(1.0)
        else:
            rows.append([1, op["in"], op["mid"], op["out"], 1])
            alpha.append(op["alpha"])
    return np.ascontiguousarray(rows, np.int32), np.asarray(alpha, np.float64)


def flatten_state_dict(sd, shapes):
    """Concatenate tensors in reference ``state_dict()`` order (arch.state_dict_shapes)."""
    return np.concatenate([np.asarray(sd[k], np.float64).reshape(-1) for k, _ in shapes])


def unflatten(flat, shapes):
    out, o = {}, 0
    for k, shp in shapes:
        n = int(np.prod(shp))
        out[k] = np.asarray(flat[o:o + n]).reshape(shp)
        o += n
    return out


class Oracle:
    """CPU oracle for one emulator likelihood.

    ``desc`` needs: kind, n_in, n_out, state_dict, priors, dolog10index, ypositive,
    X_mean, X_std, y_mean, y_std, sigma, data, inv_cov, temperature
    (a ``linna_b200.synthetic.Problem`` or any object with those attributes).
    """

    def __init__(self, desc, arch_module):
        build()
        self.lib = ctypes.CDLL(LIB)
        self.arch = arch_module
        self.kind = str(desc.kind)
        self.n_in, self.n_out = int(desc.n_in), int(desc.n_out)
        self.shapes = arch_module.state_dict_shapes(self.kind, self.n_in, self.n_out)
        self.ops, self.alpha = _ops_from_arch(arch_module.chto_ops(self.kind, self.n_in, self.n_out))
        self.has_linear = int(self.kind == "ChtoModelv2_linear")
        self.w64 = flatten_state_dict(desc.state_dict, self.shapes)
        self.prior_kind = np.asarray([_PRIOR_KIND[p["dist"]] for p in desc.priors], np.int32)
        self.a1 = np.asarray([p["arg1"] for p in desc.priors], np.float64)
        self.a2 = np.asarray([p["arg2"] for p in desc.priors], np.float64)
        lf = np.zeros(self.n_in, np.int32)
        if desc.dolog10index is not None:
            lf[list(desc.dolog10index)] = 1
        self.log10 = lf
        self.ypositive = int(bool(desc.ypositive))
        f32 = lambda a: np.asarray(a, np.float32)   # the reference stores all of these as f32
        self.x_mean, self.x_std = f32(desc.X_mean), f32(desc.X_std)
        self.y_mean, self.y_std = f32(desc.y_mean), f32(desc.y_std)
        self.sigma = f32(desc.sigma)
        self.data = f32(desc.data) if desc.data is not None else np.zeros(self.n_out, np.float32)
        self.invcov = f32(desc.inv_cov) if desc.inv_cov is not None else np.eye(self.n_out, dtype=np.float32)
        self.T = float(desc.temperature)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None

    def lnp(self, u, dtype=np.float32, grad=False, want=()):
        """Returns dict(lnp, [grad], [m], [yhat], [theta])."""
        dt = np.dtype(dtype)
        c = lambda a: np.ascontiguousarray(np.asarray(a).astype(np.float32), dt)  # f32-stored constants
        u = np.ascontiguousarray(u, dt).reshape(-1, self.n_in)
        n = u.shape[0]
        w = np.ascontiguousarray(self.w64.astype(np.float32), dt)
        out = dict(lnp=np.empty(n, dt))
        g = np.empty((n, self.n_in), dt) if grad else None
        m = np.empty((n, self.n_out), dt) if "m" in want else None
        yh = np.empty((n, self.n_out), dt) if "yhat" in want else None
        th = np.empty((n, self.n_in), dt) if "theta" in want else None
        fn = getattr(self.lib, "linna_oracle_lnp_f32" if dt == np.float32 else "linna_oracle_lnp_f64")
        real = ctypes.c_float if dt == np.float32 else ctypes.c_double
        fn.restype = ctypes.c_int
        args = [self._p(self.ops), ctypes.c_int(len(self.ops)), self._p(w), ctypes.c_int(self.has_linear),
                ctypes.c_int(self.n_in), ctypes.c_int(self.n_out)]
        # prior args and alpha are Python floats in the reference: a float64 scalar times an f32
        # tensor is f32 arithmetic with the f32-rounded scalar, and stays float64 in the f64 run.
        keep = [np.ascontiguousarray(self.alpha, dt), self.prior_kind, np.ascontiguousarray(self.a1, dt),
                np.ascontiguousarray(self.a2, dt), self.log10, c(self.x_mean), c(self.x_std), c(self.y_mean),
                c(self.y_std)]
        args += [self._p(k) for k in keep]
        tail = [c(self.sigma), c(self.data), c(self.invcov)]
        args += [ctypes.c_int(self.ypositive)] + [self._p(k) for k in tail] + [real(self.T), self._p(u),
                 ctypes.c_long(n), self._p(out["lnp"]), self._p(g), self._p(m), self._p(yh), self._p(th)]
        rc = fn(*args)
        if rc != 0:
            raise RuntimeError("oracle lnp failed rc=%d" % rc)
        if grad:
            out["grad"] = g
        if m is not None:
            out["m"] = m
        if yh is not None:
            out["yhat"] = yh
        if th is not None:
            out["theta"] = th
        return out

    def train_step(self, w, adam_m, adam_v, step, X, Y, data_n, icov_n, lr, wd=1e-4, beta1=0.9, beta2=0.999,
                   eps=1e-8, dtype=np.float32, do_update=True):
        """One AdamW step; ``w``, ``adam_m``, ``adam_v`` are flat arrays updated in place.
        Returns dict(loss, grads, loss_rows, chisq_md, chisq_nnd, yhat)."""
        dt = np.dtype(dtype)
        assert w.dtype == dt and adam_m.dtype == dt and adam_v.dtype == dt
        c = lambda a: np.ascontiguousarray(np.asarray(a).astype(np.float32), dt)
        X = np.ascontiguousarray(np.asarray(X, np.float32), dt).reshape(-1, self.n_in)
        Y = np.ascontiguousarray(np.asarray(Y, np.float32), dt).reshape(-1, self.n_out)
        B = X.shape[0]
        loss = np.zeros(1, dt)
        grads = np.zeros(w.size, dt)
        rows, cmd, cnd = np.zeros(B, dt), np.zeros(B, dt), np.zeros(B, dt)
        yhat = np.zeros((B, self.n_out), dt)
        fn = getattr(self.lib, "linna_oracle_train_step_f32" if dt == np.float32 else "linna_oracle_train_step_f64")
        fn.restype = ctypes.c_int
        real = ctypes.c_float if dt == np.float32 else ctypes.c_double
        keep = [np.ascontiguousarray(self.alpha, dt), self.log10, c(self.x_mean), c(self.x_std), c(self.y_mean),
                c(self.y_std)]
        tail = [c(self.sigma), c(data_n), c(icov_n), X, Y]
        rc = fn(self._p(self.ops), ctypes.c_int(len(self.ops)), self._p(w), ctypes.c_size_t(w.size),
                ctypes.c_int(self.has_linear), ctypes.c_int(self.n_in), ctypes.c_int(self.n_out),
                *[self._p(k) for k in keep], ctypes.c_int(self.ypositive), *[self._p(k) for k in tail],
                ctypes.c_long(B), self._p(adam_m), self._p(adam_v), ctypes.c_long(step), real(lr), real(beta1),
                real(beta2), real(eps), real(wd), ctypes.c_int(int(do_update)), self._p(loss), self._p(grads),
                self._p(rows), self._p(cmd), self._p(cnd), self._p(yhat))
        if rc != 0:
            raise RuntimeError("oracle train_step failed rc=%d" % rc)
        return dict(loss=float(loss[0]), grads=grads, loss_rows=rows, chisq_md=cmd, chisq_nnd=cnd, yhat=yhat)


def normalised_loss_constants(cov, sigma_f32, y_mean, y_std, data, ypositive=False):
    """Host one-time setup of the training loss (Auxilleryfunc.__init__, linna/util.py:1060-1069):
    C_hat = D^-1 C D^-1 with D = diag(sigma*y_std) in float64, inverted in float64, cast to f32;
    data_hat = ((data/sigma) - y_mean)/y_std (log first if ypositive), NaN -> 1e-30."""
    sig = np.asarray(sigma_f32, np.float32).astype(np.float64)
    ys = np.asarray(y_std, np.float32).astype(np.float64)
    ym = np.asarray(y_mean, np.float32)
    C = np.asarray(cov, np.float64) / np.outer(sig, sig)
    d32 = np.asarray(data, np.float32) / np.asarray(sigma_f32, np.float32)
    if ypositive:
        expected = np.asarray(data, np.float32).astype(np.float64)
        # Y_invtransform_class.transform_cov (util.py:583-588) is applied to the ALREADY sigma-scaled cov
        C0 = C / np.outer(expected, expected)
        C0[C0 <= -1] = 1e-10 - 1
        C = np.log1p(C0) / np.outer(ys, ys)
        with np.errstate(invalid="ignore", divide="ignore"):
            dn = (np.log(d32) - ym) / np.asarray(y_std, np.float32)
    else:
        C = C / np.outer(ys, ys)
        dn = (d32 - ym) / np.asarray(y_std, np.float32)
    icov = np.linalg.inv(C).astype(np.float32)
    dn = np.nan_to_num(dn.astype(np.float32), nan=1e-30)
    return dn, icov


class NumpyPort:
    """Batched numpy/BLAS port of the same lnP arithmetic -- the *charitable* vectorised CPU
    baseline.  Timed by bench.py only; checked against ``Oracle`` in tests."""

    def __init__(self, oracle):
        o = self.o = oracle
        self.sd = unflatten(o.w64.astype(np.float32), o.shapes)
        self.arch_ops = o.arch.chto_ops(o.kind, o.n_in, o.n_out)

    def forward(self, xhat):
        s = xhat
        sd = self.sd
        for op in self.arch_ops:
            nm = op["name"]
            if op["kind"] == "linear":
                s = s @ sd[nm + ".weight"].T + sd[nm + ".bias"]
                if op["act"] == "relu":
                    np.maximum(s, 0, out=s)
            else:
                h = np.maximum(s @ sd[nm + ".layer1.weight"].T + sd[nm + ".layer1.bias"], 0)
                y = (h @ sd[nm + ".layer2.weight"].T + sd[nm + ".layer2.bias"]) * np.float32(op["alpha"])
                y += s @ sd[nm + ".skip_layer.weight"].T if op["in"] != op["out"] else s
                s = np.maximum(y, 0)
        if o_has_linear(self.o):
            s = s + np.float32(1e-3) * (xhat @ sd["linearlayer.weight"].T + sd["linearlayer.bias"])
        return s

    def lnp(self, u):
        from scipy.special import erf
        o = self.o
        u = np.asarray(u, np.float32)
        a1, a2 = o.a1.astype(np.float32), o.a2.astype(np.float32)
        flat = o.prior_kind == 1
        theta = u * a2 + a1
        if flat.any():
            phi = np.float32(0.5) * (1 + erf(u[:, flat] / np.float32(np.sqrt(2)))).astype(np.float32)
            theta[:, flat] = phi * (a2 - a1)[flat] + a1[flat]
        tp = theta.copy()
        lg = o.log10 == 1
        if lg.any():
            tp[:, lg] = np.log10(theta[:, lg])
        xhat = (tp - o.x_mean) / o.x_std
        yhat = self.forward(xhat)
        y = yhat * o.y_std + o.y_mean
        if o.ypositive:
            y = np.exp(y)
        d = y * o.sigma - o.data
        chi2 = np.einsum("ij,ij->i", d @ o.invcov, d)
        lnp = np.float32(-0.5) * chi2 / np.float32(o.T) - np.float32(0.5) * np.sum(u * u, axis=1)
        lnp[np.isnan(lnp)] = -np.inf
        return lnp


def _port_lnp_grad(self, u):
    """lnP and d lnP/du of the batched port: forward with saved activations, then the hand-written
    backward pass (what torch.autograd does for linna/HMCSampler.py:29-32), all in float32 numpy/BLAS."""
    from scipy.special import erf
    o = self.o
    sd = self.sd
    u = np.asarray(u, np.float32)
    a1, a2 = o.a1.astype(np.float32), o.a2.astype(np.float32)
    flat = o.prior_kind == 1
    theta = u * a2 + a1
    dtheta = np.broadcast_to(a2, u.shape).copy()
    if flat.any():
        uf = u[:, flat]
        phi = np.float32(0.5) * (1 + erf(uf / np.float32(np.sqrt(2)))).astype(np.float32)
        theta[:, flat] = phi * (a2 - a1)[flat] + a1[flat]
        dtheta[:, flat] = (a2 - a1)[flat] * np.float32(0.3989422804014327) * np.exp(np.float32(-0.5) * uf * uf)
    tp = theta.copy()
    lg = o.log10 == 1
    dtp = np.ones_like(theta)
    if lg.any():
        tp[:, lg] = np.log10(theta[:, lg])
        dtp[:, lg] = 1.0 / (theta[:, lg] * np.float32(np.log(10.0)))
    xhat = ((tp - o.x_mean) / o.x_std).astype(np.float32)
    # forward, keeping what the backward pass needs
    saved = []
    s = xhat
    for op in self.arch_ops:
        nm = op["name"]
        if op["kind"] == "linear":
            z = s @ sd[nm + ".weight"].T + sd[nm + ".bias"]
            out = np.maximum(z, 0) if op["act"] == "relu" else z
            saved.append((op, s, out, None))
        else:
            h = np.maximum(s @ sd[nm + ".layer1.weight"].T + sd[nm + ".layer1.bias"], 0)
            y = (h @ sd[nm + ".layer2.weight"].T + sd[nm + ".layer2.bias"]) * np.float32(op["alpha"])
            y += s @ sd[nm + ".skip_layer.weight"].T if op["in"] != op["out"] else s
            out = np.maximum(y, 0)
            saved.append((op, s, out, h))
        s = out
    yhat = s
    if o_has_linear(o):
        yhat = yhat + np.float32(1e-3) * (xhat @ sd["linearlayer.weight"].T + sd["linearlayer.bias"])
    y = yhat * o.y_std + o.y_mean
    if o.ypositive:
        y = np.exp(y)
    d = y * o.sigma - o.data
    q = d @ o.invcov
    chi2 = np.einsum("ij,ij->i", q, d)
    lnp = np.float32(-0.5) * chi2 / np.float32(o.T) - np.float32(0.5) * np.sum(u * u, axis=1)
    # backward
    g = (-(q + d @ o.invcov.T) * np.float32(0.5) / np.float32(o.T)) * o.sigma * o.y_std
    if o.ypositive:
        g = g * y
    gx = np.float32(1e-3) * (g @ sd["linearlayer.weight"]) if o_has_linear(o) else 0.0
    for op, sin, out, h in reversed(saved):
        nm = op["name"]
        if op["kind"] == "linear":
            if op["act"] == "relu":
                g = g * (out > 0)
            g = g @ sd[nm + ".weight"]
        else:
            g = g * (out > 0)
            gh = (g @ sd[nm + ".layer2.weight"]) * np.float32(op["alpha"]) * (h > 0)
            gs = g @ sd[nm + ".skip_layer.weight"] if op["in"] != op["out"] else g
            g = gh @ sd[nm + ".layer1.weight"] + gs
    g = (g + gx) / o.x_std * dtp * dtheta - u
    lnp[np.isnan(lnp)] = -np.inf
    return lnp, g.astype(np.float32)


NumpyPort.lnp_grad = _port_lnp_grad


def o_has_linear(o):
    return bool(o.has_linear)
