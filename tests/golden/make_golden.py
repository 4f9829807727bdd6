#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (``/root/reference/linna``) in the build container.

    cd /tmp && python /root/repo/tests/golden/make_golden.py

The reference publishes no known-answer values for the likelihood path
(SURVEY 4, 8c), so the goldens are produced by executing its own code
(``linna.util.Log_prob``, ``Predictor.predict``, ``torch.autograd.grad``,
``Loss_fn``, ``torch.optim.AdamW``, ``linna/HMCSampler.py``) on

  * the fixture it ships (tests/test_data/2dgaussian_Fulltconn/iter_0), and
  * synthetic problems from ``linna_b200.synthetic`` (numpy-RNG, reproducible).

Each case is written twice: float32 exactly as the reference computes it, and a
float64 run of the same reference modules (``model.double()``) used for error
budgeting.  The reference does not exist on the GPU box, so only these files
travel.  Must be run from a cwd that does not contain the repo's ``linna`` shim.
"""
import importlib.util
import io
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path = [p for p in sys.path if os.path.realpath(p or ".") != os.path.realpath(REPO)]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


refstubs = _load("refstubs", os.path.join(REPO, "oracle", "refstubs.py"))
ref = refstubs.import_reference()
import linna.util as U            # noqa: E402  (the reference)
import linna.predictor_gpu as PG  # noqa: E402
import linna.nn as RNN            # noqa: E402
import linna.HMCSampler as RHMC   # noqa: E402

# the factory modules are loaded by file path so that the repo's `linna` shim is never importable
_pkg = type(sys)("linna_b200")
_pkg.__path__ = [os.path.join(REPO, "linna_b200")]
sys.modules["linna_b200"] = _pkg
arch = _load("linna_b200.arch", os.path.join(REPO, "linna_b200", "arch.py"))
synthetic = _load("linna_b200.synthetic", os.path.join(REPO, "linna_b200", "synthetic.py"))

torch.set_num_threads(1)
FIXTURE = os.path.join(refstubs.REFERENCE_ROOT, "tests", "test_data", "2dgaussian_Fulltconn", "iter_0")


# --------------------------------------------------------------------------------------
def build_reference_objects(p, dtype=torch.float32):
    """Reference Predictor / transforms / Log_prob for a synthetic Problem."""
    cls = getattr(RNN, p.kind)
    model = cls(p.n_in, p.n_out, None)
    sd = {k: torch.from_numpy(v.copy()) for k, v in p.state_dict.items()}
    model.load_state_dict(sd)
    model = model.to(dtype)
    Xt = U.X_transform_class(torch.tensor(p.X_mean, dtype=dtype), torch.tensor(p.X_std, dtype=dtype),
                             "cpu", p.dolog10index)
    yt = U.Y_transform_class(torch.tensor(p.y_mean, dtype=dtype), torch.tensor(p.y_std, dtype=dtype),
                             "cpu", ypositive=p.ypositive)
    pred = PG.Predictor(p.n_in, p.n_out, model=model, X_transform=Xt, y_transform=yt, device="cpu",
                        outdir=None)
    yinv = U.Y_invtransform_data(p.sigma, "cpu")
    yinv.sigma = torch.tensor(p.sigma.astype(np.float32)).to(dtype)   # reference casts sigma to f32 (util.py:453)
    transform = U.Transform(p.priors)
    return pred, yinv, transform


def eval_reference(p, u, dtype=torch.float32, want_m_rows=8, want_hess_rows=2):
    pred, yinv, transform = build_reference_objects(p, dtype)
    data = torch.tensor(p.data.astype(np.float32)).to(dtype)
    invcov = torch.tensor(p.inv_cov.astype(np.float32)).to(dtype)
    if dtype == torch.float64:
        data = torch.tensor(p.data.astype(np.float32).astype(np.float64))
        invcov = torch.tensor(p.inv_cov.astype(np.float32).astype(np.float64))
    lp = U.Log_prob(data, invcov, pred, yinv, transform, p.temperature, U.gaussianlogliklihood, nograd=False)
    lnp, grad = [], []
    for row in u:
        x = torch.tensor(row, dtype=dtype).clone().requires_grad_()
        val = lp(x, returntorch=True, inputnumpy=False)
        g = torch.autograd.grad(val, x)[0]
        lnp.append(val.item())
        grad.append(g.detach().numpy().astype(np.float64))
    # Hessian of lnP at the first rows by double backward: the body of the reference's Ddlnp.__call__
    # (util.py:1043-1051; its constructor is broken at HEAD, SURVEY Q2, so the Log_prob above is used directly)
    hess = []
    for row in u[:want_hess_rows]:
        x = torch.tensor(row, dtype=dtype).clone().requires_grad_()
        val = lp(x, returntorch=True, inputnumpy=False)
        g1 = torch.autograd.grad(val, x, create_graph=True)[0]
        hess.append(torch.stack([torch.autograd.grad(g1[i], x, retain_graph=True)[0] for i in range(len(g1))]).detach().numpy())
    # batched predict through the reference's own Predictor (predictor_gpu.py:461-504)
    ub = torch.tensor(u[:want_m_rows], dtype=dtype)
    theta = transform(ub, inputnumpy=False, returnnumpy=False).reshape(-1, p.n_in)
    with torch.no_grad():
        xhat = pred.X_transform(theta)
        yhat = pred.model(xhat)
        y = pred.predict(theta, no_grad=True)
        m = yinv(y)
    return dict(lnp=np.array(lnp), grad=np.array(grad), theta=theta.detach().numpy(),
                yhat=yhat.numpy(), y=y.numpy(), m=m.detach().numpy(), hess=np.array(hess, dtype=np.float64))


def pack_problem(p):
    out = dict(kind=np.array(p.kind), n_in=p.n_in, n_out=p.n_out,
               X_mean=p.X_mean, X_std=p.X_std, y_mean=p.y_mean, y_std=p.y_std,
               sigma=p.sigma, data=p.data, temperature=p.temperature,
               ypositive=int(p.ypositive),
               dolog10index=np.array(p.dolog10index if p.dolog10index is not None else [], np.int64),
               prior_dist=np.array([pr["dist"] for pr in p.priors]),
               prior_arg1=np.array([pr["arg1"] for pr in p.priors], np.float64),
               prior_arg2=np.array([pr["arg2"] for pr in p.priors], np.float64))
    return out


def synth_case(name, n_in, n_out, nrows, kind="ChtoModelv2", scale=0.3, **kw):
    p = synthetic.make_problem(n_in, n_out, kind=kind, **kw)
    # data vector: m(theta0) + sigma*noise with m from the reference in float64
    pred, yinv, transform = build_reference_objects(p, torch.float64)
    with torch.no_grad():
        m0 = yinv(pred.predict(torch.tensor(p.theta0, dtype=torch.float64), no_grad=True).view(1, -1))
    p.set_data_from_prediction(m0.numpy().ravel())
    u = synthetic.walkers(nrows, n_in, scale=scale, seed=1)
    if nrows >= 4:  # a few prior-wide and far-out rows (large |lnL|, erf tails)
        wide = synthetic.walkers(nrows, n_in, scale=1.0, seed=5)
        u[-3:] = wide[-3:]
        u[-1] *= 3.0
    f32 = eval_reference(p, u, torch.float32)
    f64 = eval_reference(p, u, torch.float64)
    out = pack_problem(p)
    out.update(u=u, make_kwargs=np.array(repr(dict(n_in=n_in, n_out=n_out, kind=kind, **kw))))
    for k, v in f32.items():
        out["f32_" + k] = v.astype(np.float32) if k != "lnp" else v.astype(np.float64)
    for k, v in f64.items():
        out["f64_" + k] = v
    # tiny integrity pins of the regenerated weights (tests regenerate them from the seed)
    out["w_checksum"] = np.array([float(np.sum(v.astype(np.float64))) for v in p.state_dict.values()])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "lnp[:3] f32", f32["lnp"][:3], "f64", f64["lnp"][:3], "max|dlnp|",
          np.max(np.abs(f32["lnp"] - f64["lnp"])), flush=True)
    return p


# --------------------------------------------------------------------------------------
def fixture_case():
    tmp = tempfile.mkdtemp()
    d = os.path.join(tmp, "iter_0")
    shutil.copytree(FIXTURE, d)
    pred, yinv = U.retrieve_model(d, 2, 2)
    with open(os.path.join(d, "model_args.pkl"), "rb") as f:
        import pickle
        args = pickle.load(f)
    cov, inv_cov, sigma, data = args[1], args[2], args[3], args[6]
    priors = [dict(param="x%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)]  # tests/test_main.py
    transform = U.Transform(priors)
    out = {}
    sd = pred.model.state_dict()
    for k, v in sd.items():
        out["sd_" + k] = v.numpy()
    out.update(X_mean=pred.X_transform.X_mean.numpy(), X_std=pred.X_transform.X_std.numpy(),
               y_mean=pred.y_transform.y_mean.numpy(), y_std=pred.y_transform.y_std.numpy(),
               sigma=yinv.sigma.detach().numpy(), cov=np.asarray(cov), inv_cov=np.asarray(inv_cov),
               data=np.asarray(data), sigma64=np.asarray(sigma))
    # predict goldens
    theta = np.array([[0.3, -0.7], [1.0, 1.0], [-1.5, 0.2], [2.0, 2.0], [0.0, 0.0], [-1.9, 1.9]], np.float32)
    with torch.no_grad():
        y = pred.predict(torch.from_numpy(theta))
        y1 = pred.predict(torch.from_numpy(theta[0]))
        m = yinv(y)
    out.update(theta=theta, y=y.numpy(), y_1d=y1.numpy(), m=m.detach().numpy())
    # Log_prob + gradient
    lp = U.Log_prob(np.asarray(data), np.asarray(inv_cov), pred, yinv, transform, 1.0,
                    U.gaussianlogliklihood, nograd=False)
    rng = np.random.default_rng(7)
    u = np.concatenate([np.array([[0.25, -0.4]]), rng.standard_normal((15, 2))]).astype(np.float32)
    u[-1] = [4.0, -6.0]
    lnp, grad = [], []
    for row in u:
        x = torch.from_numpy(row.copy()).requires_grad_()
        v = lp(x, inputnumpy=False)
        grad.append(torch.autograd.grad(v, x)[0].numpy())
        lnp.append(v.item())
    out.update(u=u, lnp=np.array(lnp), grad=np.array(grad))
    # the numpy-input, no-grad entry the samplers use (util.py:990-1021)
    lp2 = U.Log_prob(np.asarray(data), np.asarray(inv_cov), pred, yinv, transform, 4.0,
                     U.gaussianlogliklihood, nograd=True)
    out["lnp_T4"] = np.array([lp2(row.astype(np.float64)).item() for row in u])
    # reference HMC chain on this model (HMCSampler.py:19-68), fixed seeds
    torch.manual_seed(11)
    np.random.seed(11)
    samp = RHMC.HMCSampler(lambda x: lp(x, inputnumpy=False), torch.tensor([0.1, -0.2]), torch.ones(2),
                           transform=lambda x: transform(x, returnnumpy=False, inputnumpy=False))
    import linna.HMCSampler as _h
    _h.tqdm = lambda x: x
    chain = samp.sample(25, 5, 0.1)
    out["hmc_x"] = np.array([c["x"] for c in chain])
    out["hmc_lnp"] = np.array([float(c["lnP"]) for c in chain])
    out["hmc_acc"] = np.array([c["accepted"] for c in chain])
    out["hmc_ratio"] = np.array([float(c["accpet_ratio"]) for c in chain])
    np.savez_compressed(os.path.join(HERE, "fixture.npz"), **out)
    print("fixture lnp[0]", lnp[0], "grad[0]", grad[0], "y[0]", y.numpy()[0], flush=True)
    shutil.rmtree(tmp)


# --------------------------------------------------------------------------------------
def init_case():
    """Pins of the reference weight init (nn.py:34-43, :91-108) under torch.manual_seed."""
    out = {}
    for kind, n_in, n_out in (("ChtoModelv2", 2, 2), ("ChtoModelv2", 5, 40), ("ChtoModelsimple", 3, 4),
                              ("ChtoModelv2_linear", 3, 4)):
        torch.manual_seed(1234)
        m = getattr(RNN, kind)(n_in, n_out, None)
        tag = "%s_%d_%d" % (kind, n_in, n_out)
        sd = m.state_dict()
        out[tag + "_keys"] = np.array(list(sd.keys()))
        out[tag + "_sum"] = np.array([float(v.double().sum()) for v in sd.values()])
        out[tag + "_abssum"] = np.array([float(v.double().abs().sum()) for v in sd.values()])
        out[tag + "_first"] = np.array([float(v.flatten()[0]) for v in sd.values()])
    np.savez_compressed(os.path.join(HERE, "init.npz"), **out)
    print("init pins written", flush=True)


# --------------------------------------------------------------------------------------
def train_case(name, n_in, n_out, batch, nsteps, kind="ChtoModelv2", store_full=True, ypositive=False):
    """Loss_fn (util.py:1055-1116), backward and AdamW (predictor_gpu.py:267-288) goldens."""
    p = synthetic.make_problem(n_in, n_out, kind=kind, ypositive=ypositive, seed=4)
    pred, yinv, transform = build_reference_objects(p, torch.float32)
    with torch.no_grad():
        m0 = yinv(pred.predict(torch.tensor(p.theta0, dtype=torch.float32), no_grad=True).view(1, -1))
    p.set_data_from_prediction(m0.numpy().ravel().astype(np.float64))
    rng = np.random.default_rng(9)
    theta = synthetic.training_set(p, batch * nsteps, seed=3, spread=0.3)
    # targets: a perturbed copy of the emulator itself plus 1 % noise
    p2 = synthetic.make_problem(n_in, n_out, kind=kind, ypositive=ypositive, seed=4)
    for k in p2.state_dict:
        p2.state_dict[k] = (p2.state_dict[k] * (1 + 0.05 * rng.standard_normal(p2.state_dict[k].shape))).astype(np.float32)
    pred2, yinv2, _ = build_reference_objects(p2, torch.float32)
    with torch.no_grad():
        target = yinv2(pred2.predict(torch.tensor(theta, dtype=torch.float32), no_grad=True)).numpy().astype(np.float64)
    target *= (1 + 0.01 * rng.standard_normal(target.shape))
    if not ypositive:
        target[0, 0] = 1e10     # exercises the mask (util.py:1072)
    device = "cpu"
    ytd = U.Y_transform_data(p.sigma, device)
    data_t = torch.from_numpy(p.data.astype(np.float32))
    yinvt = U.Y_invtransform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), data_t, device, ypositive=ypositive)
    loss_fn = U.Loss_fn(data_t, torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinvt, device)
    val_fn = U.Val_metric_fn(data_t, torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinvt, device)
    model = pred.model
    model.train()
    lr = 2e-3
    opt = torch.optim.AdamW(params=model.parameters(), lr=lr, weight_decay=1e-4)
    out = pack_problem(p)
    out.update(cov=p.cov if n_out <= 64 else np.zeros(0), theta=theta, target=target, lr=lr, batch=batch, nsteps=nsteps)
    losses = []
    keys = list(model.state_dict().keys())
    for s in range(nsteps):
        X = torch.tensor(theta[s * batch:(s + 1) * batch], dtype=torch.float32)
        Y = torch.tensor(target[s * batch:(s + 1) * batch], dtype=torch.float32)
        opt.zero_grad()
        y_pred = model(pred.X_transform(X))
        loss = loss_fn(y_pred, Y)
        loss.backward()
        if s == 0:
            grads = {k: prm.grad.detach().numpy().copy() for k, prm in model.named_parameters()}
            with torch.no_grad():
                out["val_metric"] = val_fn(y_pred.detach(), Y).numpy()
                l_, cmd_, cnd_ = loss_fn.auxileryfunction(y_pred.detach(), Y)
                out["loss_rows"] = l_.numpy()
                out["chisqMd"] = cmd_.numpy()
                out["chisqnnd"] = cnd_.numpy()
        opt.step()
        losses.append(loss.item())
    out["losses"] = np.array(losses)
    out["keys"] = np.array(keys)
    # the same steps with the same (float32-valued) constants and inputs, every operation in float64: the error budget
    # the float32 reference itself spends (tests hold the CUDA step to a small multiple of it)
    pred64, _, _ = build_reference_objects(p, torch.float64)
    ytd64 = U.Y_transform_data(p.sigma, device)
    ytd64.sigma = ytd64.sigma.detach().double()
    data64 = data_t.double()
    yinvt64 = U.Y_invtransform_class(torch.tensor(p.y_mean).double(), torch.tensor(p.y_std).double(), data64, device,
                                     ypositive=ypositive)
    loss64 = U.Loss_fn(data64, torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd64, yinvt64, device)
    aux64 = loss64.auxileryfunction
    aux64.inv_transformed_cov = torch.inverse(aux64.transformed_cov).detach()      # the reference casts this one to float32
    aux64.inv_transformed_cov = aux64.inv_transformed_cov.float().double()         # same float32-valued constant, float64 math
    aux64.data_in = aux64.data_in.float().double()
    model64 = pred64.model
    model64.train()
    opt64 = torch.optim.AdamW(params=model64.parameters(), lr=lr, weight_decay=1e-4)
    losses64 = []
    for s in range(nsteps):
        X = torch.tensor(theta[s * batch:(s + 1) * batch], dtype=torch.float32).double()
        Y = torch.tensor(target[s * batch:(s + 1) * batch], dtype=torch.float32).double()
        opt64.zero_grad()
        y_pred = model64(pred64.X_transform(X))
        loss = loss64(y_pred, Y)
        loss.backward()
        if s == 0:
            grads64 = {k: prm.grad.detach().numpy().copy() for k, prm in model64.named_parameters()}
            with torch.no_grad():
                l_, cmd_, cnd_ = aux64(y_pred.detach(), Y)
                out["f64_loss_rows"], out["f64_chisqMd"], out["f64_chisqnnd"] = l_.numpy(), cmd_.numpy(), cnd_.numpy()
        opt64.step()
        losses64.append(loss.item())
    out["f64_losses"] = np.array(losses64)
    sd64 = model64.state_dict()
    # what the float32 reference itself loses against float64 (max-norm relative, per tensor)
    out["f32_grad0_err"] = np.array([float(np.max(np.abs(grads[k] - grads64[k])) / max(np.max(np.abs(grads64[k])), 1e-300)) for k in keys])
    out["f32_final_err"] = np.array([float(np.max(np.abs(model.state_dict()[k].numpy() - sd64[k].numpy()))) for k in keys])
    sd = model.state_dict()
    if store_full:
        for k in keys:
            out["grad0_" + k] = grads[k]
            out["final_" + k] = sd[k].numpy()
            out["f64_grad0_" + k] = grads64[k]
            out["f64_final_" + k] = sd64[k].numpy()
    else:
        out["grad0_sum"] = np.array([float(grads[k].astype(np.float64).sum()) for k in keys])
        out["grad0_norm"] = np.array([float(np.linalg.norm(grads[k].astype(np.float64))) for k in keys])
        out["final_sum"] = np.array([float(sd[k].double().sum()) for k in keys])
        out["final_norm"] = np.array([float(sd[k].double().norm()) for k in keys])
        out["grad0_layer8_row0"] = grads["layer8.weight"][0]
        out["final_layer8_row0"] = sd["layer8.weight"][0].numpy()
        out["grad0_layer1"] = grads["layer1.weight"]
        out["final_layer1"] = sd["layer1.weight"].numpy()
        out["f64_grad0_norm"] = np.array([float(np.linalg.norm(grads64[k])) for k in keys])
        out["f64_final_norm"] = np.array([float(sd64[k].norm()) for k in keys])
        out["f64_grad0_layer8_row0"] = grads64["layer8.weight"][0]
        out["f64_grad0_layer1"] = grads64["layer1.weight"]
        out["f64_final_layer1"] = sd64["layer1.weight"].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "losses", losses, "f64", losses64, "f32 grad err", out["f32_grad0_err"].max(), "final", out["f32_final_err"].max(), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["fixture", "init", "synth", "train"]
    if "fixture" in which:
        fixture_case()
    if "init" in which:
        init_case()
    if "synth" in which:
        synth_case("c1", 33, 33, 8)
        synth_case("c3s", 30, 500, 48)
        synth_case("c3mix", 30, 500, 24, priors="mixed", log10=True, temperature=4.0, cond=1e4, seed=2)
        synth_case("ypos", 5, 40, 16, ypositive=True, priors="gauss", seed=3, scale=0.2)
        synth_case("c4s", 50, 1500, 12, seed=5)
        synth_case("simple", 6, 45, 16, kind="ChtoModelsimple", priors="mixed", seed=6)
        synth_case("v2lin", 6, 8, 16, kind="ChtoModelv2_linear", seed=7)
        synth_case("tiny", 3, 4, 40, seed=8, scale=1.0)
    if "train" in which:
        train_case("train_small", 6, 8, 16, 3)
        train_case("train_c3", 30, 500, 100, 2, store_full=False)
        train_case("train_ypos", 4, 6, 8, 2, ypositive=True)
