"""GPU tests of the reference-shaped autograd entry points that round 1 left as stubs: ``Loss_fn`` / ``Val_metric_fn`` on
free-standing tensors (linna/util.py:1070-1127), ``Predictor.predict(X, no_grad=False)`` (linna/predictor_gpu.py:495-496)
and ``model(x)`` with parameters that require grad (linna/predictor_gpu.py:279-283).  The checker is a float64 torch
restatement of the same formulas, differentiated by torch.autograd."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, synthetic

pytestmark = pytest.mark.gpu


def torch_emulator(sd, kind, n_in, n_out, x):
    """float64 torch restatement of linna/nn.py:45-56, :110-133 on a dict of float64 tensors."""
    s = x
    for op in arch.chto_ops(kind, n_in, n_out):
        nm = op["name"]
        if op["kind"] == "linear":
            s = s @ sd[nm + ".weight"].T + sd[nm + ".bias"]
            if op["act"] == "relu":
                s = torch.relu(s)
        else:
            h = torch.relu(s @ sd[nm + ".layer1.weight"].T + sd[nm + ".layer1.bias"])
            s = torch.relu(op["alpha"] * (h @ sd[nm + ".layer2.weight"].T + sd[nm + ".layer2.bias"]) + s @ sd[nm + ".skip_layer.weight"].T)
    return s


def _loss_objects(p, ypositive=False):
    import linna.util as U
    sig = np.asarray(p.sigma, np.float32)
    ytd = U.Y_transform_data(sig, "cpu")
    yinv = U.Y_invtransform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), torch.tensor(p.data.astype(np.float32)), "cpu",
                                  ypositive=ypositive)
    args = (torch.tensor(p.data.astype(np.float32)), torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinv, "cpu")
    return U.Loss_fn(*args), U.Val_metric_fn(*args)


def test_loss_fn_on_tensors_matches_reference_formulas_and_backpropagates():
    p = synthetic.make_problem(5, 40, seed=9)
    rng = np.random.default_rng(1)
    p.data = rng.standard_normal(40) * p.sigma
    loss_fn, val_fn = _loss_objects(p)
    aux = loss_fn.auxileryfunction
    B = 37
    y_pred = torch.from_numpy(rng.standard_normal((B, 40)).astype(np.float32)).cuda().requires_grad_()
    y_target = (p.data[None, :] + 0.5 * p.sigma * rng.standard_normal((B, 40))).astype(np.float32)
    y_target[3, 7] = 1e10                                      # masked entry (util.py:1072)
    yt = torch.from_numpy(y_target).cuda()
    loss = loss_fn(y_pred, yt)
    loss.backward()
    # float64 restatement of util.py:1070-1088, :1114-1115
    A = aux.inv_transformed_cov.double()
    dhat = aux.data_in.double()
    yp64 = y_pred.detach().cpu().double().requires_grad_()
    t = (torch.from_numpy(y_target).double() / aux.y_transform_data.sigma.detach().double() - aux.y_inv_transform.y_mean.double()) / \
        aux.y_inv_transform.y_std.double()
    mask = (torch.from_numpy(y_target) == 1e10) | (torch.from_numpy(y_target) == 1e-30) | (aux.data_in == 1e-30)
    def chi(d):
        d = torch.where(mask, torch.zeros_like(d), d)
        return ((d @ A) * d).sum(-1)
    md = torch.clamp(chi(t - dhat), min=0.5 * 40)
    ref_rows = chi(t - yp64) / md
    ref = ref_rows.mean()
    ref.backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 2e-5 * abs(float(ref.detach()))
    g, gr = y_pred.grad.cpu().double(), yp64.grad
    assert float((g - gr).abs().max()) <= 2e-5 * float(gr.abs().max())
    assert float(g[3, 7]) == 0.0
    l_rows, cmd, cnd = aux(y_pred.detach(), yt)
    np.testing.assert_allclose(l_rows.cpu().numpy(), ref_rows.detach().numpy(), rtol=3e-5)
    np.testing.assert_allclose(cmd.cpu().numpy(), md.numpy(), rtol=3e-6)
    np.testing.assert_allclose(cnd.cpu().numpy(), chi(yp64.detach() - dhat).numpy(), rtol=3e-5)
    vm = val_fn(y_pred.detach(), yt)
    frac = (chi(yp64.detach() - dhat) / md - 1).abs()
    np.testing.assert_allclose(vm.numpy(), [ref_rows.detach().median(), frac.max(), frac.median()], rtol=1e-4)
    # host tensors go to the GPU and come back
    lh = loss_fn(y_pred.detach().cpu(), yt.cpu())
    assert not lh.is_cuda and abs(float(lh) - float(loss)) < 1e-9 + 1e-6 * abs(float(loss))


@pytest.mark.parametrize("kind,n_in,n_out,log10,ypos", [("ChtoModelv2", 6, 8, False, False), ("ChtoModelsimple", 5, 45, True, False),
                                                          ("ChtoModelv2", 4, 40, False, True)])
def test_predict_with_grad_and_module_autograd(kind, n_in, n_out, log10, ypos):
    import linna.nn as N
    import linna.predictor_gpu as PG
    import linna.util as U
    p = synthetic.make_problem(n_in, n_out, kind=kind, seed=12, log10=log10, ypositive=ypos)
    model = getattr(N, kind)(n_in, n_out, None)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.state_dict.items()})
    xt = U.X_transform_class(torch.tensor(p.X_mean), torch.tensor(p.X_std), "cpu", p.dolog10index)
    yt = U.Y_transform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), "cpu", ypositive=ypos)
    pred = PG.Predictor(n_in, n_out, model=model, X_transform=xt, y_transform=yt, device="cpu")
    rng = np.random.default_rng(3)
    theta = synthetic.training_set(p, 19, seed=5, spread=0.3).astype(np.float32)
    cot = rng.standard_normal((19, n_out)).astype(np.float32)
    sd64 = {k: torch.from_numpy(v.astype(np.float64)) for k, v in p.state_dict.items()}

    def ref_predict(th):
        th2 = th.clone()
        if log10:
            for i in p.dolog10index:
                th2[:, i] = torch.log10(th[:, i])
        xh = (th2 - torch.from_numpy(p.X_mean).double()) / torch.from_numpy(p.X_std).double()
        y = torch_emulator(sd64, kind, n_in, n_out, xh) * torch.from_numpy(p.y_std).double() + torch.from_numpy(p.y_mean).double()
        return torch.exp(y) if ypos else y
    # ---- Predictor.predict(X, no_grad=False): gradient with respect to the input
    X = torch.from_numpy(theta).cuda().requires_grad_()
    y = pred.predict(X, no_grad=False)
    (y * torch.from_numpy(cot).cuda()).sum().backward()
    X64 = torch.from_numpy(theta.astype(np.float64)).requires_grad_()
    y64 = ref_predict(X64)
    (y64 * torch.from_numpy(cot).double()).sum().backward()
    assert float((y.detach().cpu().double() - y64.detach()).abs().max()) <= 1e-5 * float(y64.detach().abs().max())
    assert float((X.grad.cpu().double() - X64.grad).abs().max()) <= 5e-5 * float(X64.grad.abs().max())
    x1 = torch.from_numpy(theta[0]).cuda().requires_grad_()          # 1-D in => 1-D out, as the reference
    y1 = pred.predict(x1, no_grad=False)
    assert y1.shape == (n_out,)
    y1.sum().backward()
    assert x1.grad.shape == (n_in,) and torch.isfinite(x1.grad).all()
    # ---- model(xhat) with parameters that require grad (the reference's training forward under autograd)
    xhat = torch.from_numpy(rng.standard_normal((19, n_in)).astype(np.float32)).cuda().requires_grad_()
    model.zero_grad()
    out = model(xhat)
    (out * torch.from_numpy(cot).cuda()).sum().backward()
    sdg = {k: v.clone().requires_grad_() for k, v in sd64.items()}
    xh64 = xhat.detach().cpu().double().requires_grad_()
    o64 = torch_emulator(sdg, kind, n_in, n_out, xh64)
    (o64 * torch.from_numpy(cot).double()).sum().backward()
    assert float((xhat.grad.cpu().double() - xh64.grad).abs().max()) <= 5e-5 * float(xh64.grad.abs().max())
    for name, prm in model.named_parameters():
        ref = sdg[name].grad
        assert prm.grad is not None, name
        assert float((prm.grad.cpu().double() - ref).abs().max()) <= 5e-5 * float(ref.abs().max()) + 1e-9, name
