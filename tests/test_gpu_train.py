"""GPU parity of the fused training step (forward + loss + backward + AdamW) against goldens made
by the reference's own Loss_fn / autograd / torch.optim.AdamW (tests/golden/make_golden.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, engine, synthetic
from oracle.oracle import flatten_state_dict, normalised_loss_constants, unflatten
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


def _setup(name):
    g = load_golden(name)
    p = synthetic.make_problem(int(g["n_in"]), int(g["n_out"]), kind=str(g["kind"]), ypositive=bool(g["ypositive"]), seed=4)
    p.data = g["data"].astype(np.float64)
    if g["cov"].size:
        p.cov = g["cov"]
    dn, icov = normalised_loss_constants(p.cov, np.asarray(p.sigma, np.float32), p.y_mean, p.y_std, p.data,
                                         ypositive=p.ypositive)
    e = engine.engine_from_problem(p, with_likelihood=False)
    B = int(g["batch"])
    e.train_setup(dn, icov, B)
    shapes = arch.state_dict_shapes(p.kind, p.n_in, p.n_out)
    w = torch.from_numpy(flatten_state_dict(p.state_dict, shapes).astype(np.float32)).cuda()
    assert w.numel() == e.n_params
    return g, p, e, shapes, w, B


@pytest.mark.parametrize("name,full", [("train_small", True), ("train_ypos", True), ("train_c3", False)])
def test_loss_and_gradients(name, full):
    g, p, e, shapes, w, B = _setup(name)
    X = torch.from_numpy(g["theta"][:B].astype(np.float32)).cuda()
    Y = torch.from_numpy(g["target"][:B].astype(np.float32)).cuda()
    cmd_raw = e.train_chisq(X, Y, 1)
    cmd = torch.clamp(cmd_raw, min=0.5 * p.n_out)                      # util.py:1086
    np.testing.assert_allclose(cmd.cpu().numpy(), g["chisqMd"], rtol=2e-4)
    np.testing.assert_allclose(e.train_chisq(X, Y, 2).cpu().numpy(), g["chisqnnd"], rtol=2e-3)
    mnn = e.train_chisq(X, Y, 0)
    np.testing.assert_allclose((mnn / cmd).cpu().numpy(), g["loss_rows"], rtol=2e-3, atol=1e-9)
    grads = torch.zeros_like(w)
    loss, rows = e.train_step(X, Y, cmd, None, None, None, grads, 1, float(g["lr"]), fuse_adam=False)
    torch.cuda.synchronize()
    assert abs(float(loss) - g["losses"][0]) <= 2e-3 * abs(g["losses"][0])
    np.testing.assert_allclose(rows.cpu().numpy(), g["loss_rows"], rtol=2e-3, atol=1e-9)
    gd = unflatten(grads.cpu().numpy(), shapes)
    keys = [str(k) for k in g["keys"]]
    if full:
        for k in keys:
            ref = g["grad0_" + k]
            assert np.max(np.abs(gd[k] - ref)) <= 2e-3 * np.max(np.abs(ref)) + 1e-12, k
    else:
        norms = np.array([np.linalg.norm(gd[k].astype(np.float64)) for k in keys])
        np.testing.assert_allclose(norms, g["grad0_norm"], rtol=2e-3)
        ref = g["grad0_layer1"]
        assert np.max(np.abs(gd["layer1.weight"] - ref)) <= 2e-3 * np.max(np.abs(ref))
        ref = g["grad0_layer8_row0"]
        assert np.max(np.abs(gd["layer8.weight"][0] - ref)) <= 2e-3 * np.max(np.abs(ref))


@pytest.mark.parametrize("name,full", [("train_small", True), ("train_ypos", True), ("train_c3", False)])
@pytest.mark.parametrize("fused", [True, False])
def test_adamw_steps_match_reference(name, full, fused):
    g, p, e, shapes, w, B = _setup(name)
    m, v, grads = torch.zeros_like(w), torch.zeros_like(w), torch.zeros_like(w)
    lr = float(g["lr"])
    for s in range(int(g["nsteps"])):
        X = torch.from_numpy(g["theta"][s * B:(s + 1) * B].astype(np.float32)).cuda()
        Y = torch.from_numpy(g["target"][s * B:(s + 1) * B].astype(np.float32)).cuda()
        cmd = torch.clamp(e.train_chisq(X, Y, 1), min=0.5 * p.n_out)
        if fused:
            loss, _ = e.train_step(X, Y, cmd, w, m, v, None, s + 1, lr, fuse_adam=True)
        else:     # the data-parallel path: gradient out, (all-reduce), stand-alone AdamW
            loss, _ = e.train_step(X, Y, cmd, None, None, None, grads, s + 1, lr, fuse_adam=False)
            e.train_adamw(w, m, v, grads, s + 1, lr)
        assert abs(float(loss) - g["losses"][s]) <= 3e-3 * abs(g["losses"][s]) + 1e-9, (s, float(loss), g["losses"][s])
    wd = unflatten(w.cpu().numpy(), shapes)
    keys = [str(k) for k in g["keys"]]
    if full:
        for k in keys:
            ref = g["final_" + k]
            assert np.max(np.abs(wd[k] - ref)) <= 1e-4 * max(np.max(np.abs(ref)), 1e-3) + 3e-5, k
    else:
        ref = g["final_layer1"]
        assert np.max(np.abs(wd["layer1.weight"] - ref)) <= 1e-4 * np.max(np.abs(ref)) + 3e-5
        np.testing.assert_allclose([float(np.linalg.norm(wd[k].astype(np.float64))) for k in keys], g["final_norm"], rtol=1e-4)
    # the packed weights inside the engine follow the flat vector: predictions use the updated weights
    th = g["theta"][:4].astype(np.float32)
    y_after = e.predict(th, engine.LINNA_OUT_YHAT)
    p2 = synthetic.make_problem(p.n_in, p.n_out, kind=p.kind, ypositive=p.ypositive, seed=4)
    p2.state_dict = {k: np.ascontiguousarray(wd[k]) for k in wd}
    e2 = engine.engine_from_problem(p2, with_likelihood=False)
    np.testing.assert_allclose(y_after, e2.predict(th, engine.LINNA_OUT_YHAT), rtol=1e-6, atol=1e-6)
