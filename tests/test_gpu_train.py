"""GPU parity of the fused training step (forward + loss + backward + AdamW) against goldens made
by the reference's own Loss_fn / autograd / torch.optim.AdamW (tests/golden/make_golden.py).

Bars.  Every golden holds the reference's float32 result AND a float64 run of the same reference modules on the same
float32-valued constants, so the error the reference itself spends is known per quantity (`f32_grad0_err`: 3.5e-6 /
3.1e-6 max-norm relative for the well-conditioned cases, 0.2 for the log-normal `ypositive` case, whose normalised
covariance is ill-conditioned).  The CUDA step has to (a) agree with the reference's float32 numbers to 2e-5 -- a hundred
times tighter than round 1's 2e-3 and ~5x what was measured (profiles/r2_parity_probe.txt: 6.6e-7 .. 3.9e-6) -- and
(b) be no further from float64 than max(4 x the reference's own float32 error, 1e-5)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, engine, synthetic
from oracle.oracle import flatten_state_dict, normalised_loss_constants, unflatten
from tests.helpers import load_golden

pytestmark = pytest.mark.gpu


PATHS = ["ffma", "tc"]     # the FP32 FFMA kernels and the tensor-core (tcgen05, bf16x3) kernels: same bars


def _setup(name, path="auto"):
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
    e.set_train_path(path)
    shapes = arch.state_dict_shapes(p.kind, p.n_in, p.n_out)
    w = torch.from_numpy(flatten_state_dict(p.state_dict, shapes).astype(np.float32)).cuda()
    assert w.numel() == e.n_params
    return g, p, e, shapes, w, B


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("name,full", [("train_small", True), ("train_ypos", True), ("train_c3", False)])
def test_loss_and_gradients(name, full, path):
    g, p, e, shapes, w, B = _setup(name, path)
    X = torch.from_numpy(g["theta"][:B].astype(np.float32)).cuda()
    Y = torch.from_numpy(g["target"][:B].astype(np.float32)).cuda()
    cmd_raw = e.train_chisq(X, Y, 1)
    cmd = torch.clamp(cmd_raw, min=0.5 * p.n_out)                      # util.py:1086
    np.testing.assert_allclose(cmd.cpu().numpy(), g["chisqMd"], rtol=3e-6)
    nnd = e.train_chisq(X, Y, 2)
    np.testing.assert_allclose(nnd.cpu().numpy(), g["chisqnnd"], rtol=3e-6)
    mnn = e.train_chisq(X, Y, 0)
    np.testing.assert_allclose((mnn / cmd).cpu().numpy(), g["loss_rows"], rtol=3e-5, atol=1e-10)
    # Val_metric_fn (linna/util.py:1124-1127) through FusedTrainer.val_metric, against the reference's own output
    import types
    from linna_b200.train import FusedTrainer
    vm = FusedTrainer.val_metric(types.SimpleNamespace(engine=e), X, Y, cmd).cpu().numpy()
    np.testing.assert_allclose(vm[0], g["val_metric"][0], rtol=1e-5)
    np.testing.assert_allclose(vm[1:], g["val_metric"][1:], rtol=3e-4)     # |chi2_nn,d / chi2_M,d - 1|: a difference of near-equal numbers
    grads = torch.zeros_like(w)
    loss, rows = e.train_step(X, Y, cmd, None, None, None, grads, 1, float(g["lr"]), fuse_adam=False)
    torch.cuda.synchronize()
    assert e.last_train_kernel() == path
    assert abs(float(loss) - g["losses"][0]) <= 3e-6 * abs(g["losses"][0])
    np.testing.assert_allclose(rows.cpu().numpy(), g["loss_rows"], rtol=3e-5, atol=1e-10)
    gd = unflatten(grads.cpu().numpy(), shapes)
    keys = [str(k) for k in g["keys"]]
    budget = max(4.0 * float(g["f32_grad0_err"].max()), 1e-5)   # what the reference's own float32 spends, x 4

    def check(got, ref32, ref64, what):
        scale = np.max(np.abs(ref64))
        assert np.max(np.abs(got - ref32)) <= 2e-5 * scale + 1e-12, (what, np.max(np.abs(got - ref32)) / scale)
        assert np.max(np.abs(got - ref64)) <= budget * scale + 1e-12, (what, np.max(np.abs(got - ref64)) / scale)
    if full:
        for k in keys:
            check(gd[k], g["grad0_" + k], g["f64_grad0_" + k], k)
    else:
        norms = np.array([np.linalg.norm(gd[k].astype(np.float64)) for k in keys])
        np.testing.assert_allclose(norms, g["grad0_norm"], rtol=5e-6)
        np.testing.assert_allclose(norms, g["f64_grad0_norm"], rtol=5e-6)
        check(gd["layer1.weight"], g["grad0_layer1"], g["f64_grad0_layer1"], "layer1.weight")
        check(gd["layer8.weight"][0], g["grad0_layer8_row0"], g["f64_grad0_layer8_row0"], "layer8.weight[0]")


@pytest.mark.parametrize("name,full", [("train_small", True), ("train_ypos", True), ("train_c3", False)])
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("path", PATHS)
def test_adamw_steps_match_reference(name, full, fused, path):
    g, p, e, shapes, w, B = _setup(name, path)
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
    # AdamW divides by sqrt(v): where a gradient is within rounding of zero the update is +-lr whatever its size, so two
    # correct float32 evaluations can differ by ~lr in single weights.  The goldens record how far the reference's own
    # float32 result is from its float64 run per tensor (`f32_final_err`: up to 1.4e-3 at the C3 shape, lr = 2e-3); the
    # CUDA result may be as far from the float32 reference as that, and no further.
    # Which weights sit on such a knife edge differs between two float32 evaluations, so the cap for a single weight is
    # the largest spread seen in ANY tensor of the case (at least the tight bar), and only a few weights may use it.
    spread = float(np.max(g["f32_final_err"]))
    if full:
        for k in keys:
            ref = g["final_" + k]
            tight = 1e-4 * max(np.max(np.abs(ref)), 1e-3) + 3e-5
            for which, rr in (("f32", ref), ("f64", g["f64_final_" + k])):
                d = np.abs(wd[k] - rr)
                assert np.max(d) <= max(tight, 1.5 * spread), (k, which)
                if which == "f32":     # against the float32 reference only a few weights may use the wide cap
                    assert np.mean(d > tight) <= 2e-3 + 1.0 / d.size, k
    else:
        ref = g["final_layer1"]
        tight = 1e-4 * np.max(np.abs(ref)) + 3e-5
        for rr in (ref, g["f64_final_layer1"]):
            d = np.abs(wd["layer1.weight"] - rr)
            assert np.max(d) <= max(tight, 1.5 * spread)
            assert np.mean(d > tight) < 2e-3     # and only in a few weights
        np.testing.assert_allclose([float(np.linalg.norm(wd[k].astype(np.float64))) for k in keys], g["final_norm"], rtol=1e-4)
    # the packed weights inside the engine follow the flat vector: predictions use the updated weights
    th = g["theta"][:4].astype(np.float32)
    y_after = e.predict(th, engine.LINNA_OUT_YHAT)
    p2 = synthetic.make_problem(p.n_in, p.n_out, kind=p.kind, ypositive=p.ypositive, seed=4)
    p2.state_dict = {k: np.ascontiguousarray(wd[k]) for k in wd}
    e2 = engine.engine_from_problem(p2, with_likelihood=False)
    np.testing.assert_allclose(y_after, e2.predict(th, engine.LINNA_OUT_YHAT), rtol=1e-6, atol=1e-6)


def test_train_NN_on_reference_fixture(tmp_path):
    """The reference's own training entry point (pickled as linna.util.train_NN with model_args.pkl,
    linna/main.py:189-198, run by linna/train_gpu.py:24-38) on its shipped 20+5-point fixture."""
    import os
    import pickle
    import shutil
    import linna.util as U
    from tests.helpers import GOLDEN
    d = str(tmp_path / "iter_0") + "/"
    shutil.copytree(os.path.join(GOLDEN, "ref_fixture_iter_0"), d)
    for f in ("best.pth.tar", "last.pth.tar", "finish.pkl"):
        os.remove(d + f)
    with open(d + "model_pickle.pkl", "rb") as f:
        fn = pickle.load(f)
    with open(d + "model_args.pkl", "rb") as f:
        args = pickle.load(f)
    assert fn is U.train_NN
    args[4], args[5] = d, [d]
    args[16] = dict(args[16], num_epochs=40)
    fn(*args)
    for f in ("best.pth.tar", "last.pth.tar", "X_transform.pkl", "y_transform.pkl", "y_invtransform_data.pkl",
              "y_transform_data.pkl", "y_invtransform.pkl", "lr.npy"):
        assert os.path.isfile(d + f), f
    ck = torch.load(d + "best.pth.tar", weights_only=False)
    assert set(ck) >= {"epoch", "state_dict", "optim_dict"} and len(ck["state_dict"]) == 23
    assert set(ck["optim_dict"]) == {"state", "param_groups"} and len(ck["optim_dict"]["state"]) == 23
    # the trained emulator reloads through the reference-shaped loader and reproduces theory(x) = x roughly
    pred, yinv = U.retrieve_model(d, 2, 2)
    th = torch.tensor([[0.1, 0.2], [-0.5, 0.4]])
    m = yinv(pred.predict(th)).detach().numpy()
    assert np.all(np.isfinite(m)) and m.shape == (2, 2)


def test_training_converges_and_dp_path_equals_fused():
    """A few hundred fused steps drive the loss down by orders of magnitude; the gradient-out +
    stand-alone AdamW path (what data-parallel ranks run) follows the same trajectory."""
    from linna_b200.train import FusedTrainer
    import linna.nn as N
    import linna.util as U
    p = synthetic.make_problem(6, 8, seed=4)
    p.data = np.zeros(8)
    rng = np.random.default_rng(0)
    theta = synthetic.training_set(p, 512, seed=3, spread=0.5)
    A = rng.standard_normal((6, 8))
    target = np.tanh(theta @ A) * p.sigma + 0.3 * p.sigma          # a smooth "theory"
    p.data = target[0].copy()
    sig = np.asarray(p.sigma, np.float32)
    ytd = U.Y_transform_data(sig, "cpu")
    ymean = torch.tensor(np.median(target / sig, axis=0).astype(np.float32))
    ystd = torch.tensor((np.median(np.abs(target / sig - ymean.numpy()), axis=0)).astype(np.float32))
    yinv = U.Y_invtransform_class(ymean, ystd, torch.tensor(p.data.astype(np.float32)), "cpu")
    loss_fn = U.Loss_fn(torch.tensor(p.data.astype(np.float32)), torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinv, "cpu")
    xt = U.X_transform_class(torch.tensor(theta.mean(0).astype(np.float32)), torch.tensor(theta.std(0).astype(np.float32)), "cpu")
    yt = U.Y_transform_class(ymean, ystd, "cpu")
    X = torch.from_numpy(theta.astype(np.float32)).cuda()
    Y = torch.from_numpy(target.astype(np.float32)).cuda()
    runs = []
    for fused in (True, False):
        torch.manual_seed(3)
        model = N.ChtoModelv2(6, 8, None)
        tr = FusedTrainer(model, xt, yt, loss_fn.auxileryfunction, 128, lr=2e-3)
        cmd = tr.chisq_md(X, Y)
        if not fused:
            tr.world = 2          # forces the gradient-out path; all_reduce is skipped below (the real 2-rank step, with the
            import torch.distributed as dist      # gradient average read from peer memory: tests/test_gpu_multi.py)
            orig = dist.all_reduce
            dist.all_reduce = lambda *a, **k: None
        try:
            hist = []
            g = torch.Generator().manual_seed(1)
            for it in range(300):
                idx = torch.randperm(512, generator=g)[:128].cuda()
                hist.append(tr.step(X[idx], Y[idx], cmd[idx]).clone())
        finally:
            if not fused:
                dist.all_reduce = orig
        hist = torch.cat(hist).cpu().numpy()
        runs.append(hist)
        assert np.isfinite(hist).all()
        assert hist[-20:].mean() < 0.1 * hist[:5].mean(), (hist[:5], hist[-20:])
        vm = tr.val_metric(X, Y, cmd).cpu().numpy()
        assert vm.shape == (3,) and vm[0] < 0.1 * hist[0]
        tr.commit()
    np.testing.assert_allclose(runs[0][:50], runs[1][:50], rtol=2e-3)


@pytest.mark.parametrize("take_log", [False, True])
def test_training_set_statistics_on_device(take_log):
    """y_mean / y_std of Y_transform_class (median and median absolute deviation of y / sigma, linna/util.py:1440-1450,
    :1308-1313) by radix selection on the GPU against the reference's own torch expressions on the CPU: bit-identical
    without the log, to float32 rounding of logf with it."""
    from linna_b200 import engine
    rng = np.random.default_rng(3)
    for n, d in ((10000, 500), (4097, 33), (2, 5), (1, 3)):
        Y = rng.standard_normal((n, d)).astype(np.float32) * 3 + (5 if take_log else 0)
        if take_log:
            Y = np.abs(Y) + 1e-3
        Y[rng.integers(0, n, 7), rng.integers(0, d, 7)] = 1e10 if not take_log else 1e-30     # clipped outliers (util.py:1432)
        Y[:, 0] = 2.5                                                                          # a constant column: MAD = 0
        sigma = (0.5 + rng.random(d)).astype(np.float32)
        v = torch.from_numpy(Y) / torch.from_numpy(sigma)          # Y_transform_data, linna/util.py:432
        if take_log:
            v = torch.log(v)
        want_med = v.median(axis=0).values
        want_mad = torch.abs(v - want_med).median(axis=0).values   # median_absolute_deviation, linna/util.py:1308-1313
        med, mad = engine.column_median_mad(torch.from_numpy(Y).cuda(), sigma, take_log=take_log)
        med, mad = med.cpu(), mad.cpu()
        if take_log:
            assert torch.allclose(med, want_med, rtol=1e-6, atol=1e-6) and torch.allclose(mad, want_mad, rtol=1e-5, atol=2e-6)
        else:
            assert torch.equal(med, want_med) and torch.equal(mad, want_mad), (n, d)
        assert float(mad[0]) == 0.0
