"""CPU-only: pin the C oracle (oracle/) against the golden vectors produced by running the
reference itself (tests/golden/make_golden.py).  These are the tests that make the oracle
trustworthy as the checker for the CUDA path."""
import numpy as np
import pytest

from linna_b200 import arch
from oracle.oracle import NumpyPort, Oracle, flatten_state_dict, normalised_loss_constants, unflatten
from tests.helpers import fixture_problem, lnp_tol, load_golden, problem_from_golden, rel_inf

SYNTH = ["c1", "c3s", "c3mix", "ypos", "c4s", "simple", "v2lin", "tiny"]


def test_fixture_predict_and_lnp():
    g = load_golden("fixture")
    p = fixture_problem(g)
    o = Oracle(p, arch)
    r = o.lnp(g["u"], np.float32, grad=True)
    assert np.all(np.abs(r["lnp"] - g["lnp"]) <= lnp_tol(g["lnp"]))
    assert rel_inf(r["grad"], g["grad"]) < 2e-5
    # SURVEY 8c probe values
    assert abs(r["lnp"][0] - (-3.36370969)) < 1e-5
    np.testing.assert_allclose(r["grad"][0], [-0.53383291, 0.99314642], atol=2e-6)
    # temperature-4 variant through the numpy entry point
    p.temperature = 4.0
    r4 = Oracle(p, arch).lnp(g["u"], np.float32)
    assert np.all(np.abs(r4["lnp"] - g["lnp_T4"]) <= lnp_tol(g["lnp_T4"]))


def test_fixture_model_vectors():
    g = load_golden("fixture")
    p = fixture_problem(g)
    # predict() takes physical parameters: use gauss priors (0,1) so that theta == u
    p.priors = [dict(param="x", dist="gauss", arg1=0.0, arg2=1.0)] * 2
    r = Oracle(p, arch).lnp(g["theta"], np.float32, want=("m", "yhat"))
    y = r["yhat"] * g["y_std"] + g["y_mean"]
    assert rel_inf(y, g["y"]) < 1e-5
    assert rel_inf(r["m"], g["m"]) < 1e-5
    np.testing.assert_allclose(y[0], [0.10346876, -0.31955421], atol=1e-6)
    np.testing.assert_allclose(g["y_1d"], g["y"][0], atol=0)                 # 1-D in => same row


@pytest.mark.parametrize("name", SYNTH)
def test_synthetic_f32(name):
    g = load_golden(name)
    p = problem_from_golden(g)
    o = Oracle(p, arch)
    r = o.lnp(g["u"], np.float32, grad=True, want=("m", "yhat", "theta"))
    k = g["f32_m"].shape[0]
    assert rel_inf(r["theta"][:k], g["f32_theta"]) < 1e-6
    assert rel_inf(r["m"][:k], g["f32_m"]) < 1e-5, "north-star 1e-5 relative on predicted data vectors"
    # lnP: both are float32 evaluations of the same expression -> compare through the float64 run
    err_ref = np.abs(g["f32_lnp"] - g["f64_lnp"])
    err_orc = np.abs(r["lnp"].astype(np.float64) - g["f64_lnp"])
    tol = lnp_tol(g["f64_lnp"])
    assert np.all(err_orc <= np.maximum(4 * tol, 4 * err_ref)), (err_orc.max(), err_ref.max())
    assert rel_inf(r["grad"], g["f32_grad"]) < 5e-4
    # the batched numpy port (CPU baseline of bench.py) computes the same thing
    lp = NumpyPort(o).lnp(g["u"])
    assert np.all(np.abs(lp - g["f64_lnp"]) <= np.maximum(4 * tol, 4 * err_ref))
    lp2, gp = NumpyPort(o).lnp_grad(g["u"])                                  # --mode grad baseline
    assert np.all(np.abs(lp2 - g["f64_lnp"]) <= np.maximum(4 * tol, 4 * err_ref))
    assert rel_inf(gp, g["f64_grad"]) < 5e-4


@pytest.mark.parametrize("name", SYNTH)
def test_synthetic_f64(name):
    g = load_golden(name)
    p = problem_from_golden(g)
    r = Oracle(p, arch).lnp(g["u"], np.float64, grad=True, want=("m",))
    k = g["f64_m"].shape[0]
    assert rel_inf(r["m"][:k], g["f64_m"]) < 1e-12
    np.testing.assert_allclose(r["lnp"], g["f64_lnp"], rtol=1e-11, atol=1e-9)
    assert rel_inf(r["grad"], g["f64_grad"]) < 1e-9


def _train_setup(name):
    g = load_golden(name)
    p = problem_from_golden(dict(g, make_kwargs=np.array(repr(dict(
        n_in=int(g["n_in"]), n_out=int(g["n_out"]), kind=str(g["kind"]), ypositive=bool(g["ypositive"]), seed=4))),
        w_checksum=np.array([float(np.sum(v.astype(np.float64))) for v in __import__("linna_b200.synthetic", fromlist=["x"]).make_problem(
            int(g["n_in"]), int(g["n_out"]), kind=str(g["kind"]), ypositive=bool(g["ypositive"]), seed=4).state_dict.values()])))
    return g, p


@pytest.mark.parametrize("name,full", [("train_small", True), ("train_ypos", True), ("train_c3", False)])
def test_training_step(name, full):
    g, p = _train_setup(name)
    p.cov = g["cov"] if g["cov"].size else p.cov
    o = Oracle(p, arch)
    dn, icov = normalised_loss_constants(p.cov, np.asarray(p.sigma, np.float32), p.y_mean, p.y_std, p.data,
                                         ypositive=p.ypositive)
    w = o.w64.astype(np.float32)
    m, v = np.zeros_like(w), np.zeros_like(w)
    B, nsteps = int(g["batch"]), int(g["nsteps"])
    keys = [str(k) for k in g["keys"]]
    assert keys == [k for k, _ in o.shapes]
    for s in range(nsteps):
        X, Y = g["theta"][s * B:(s + 1) * B], g["target"][s * B:(s + 1) * B]
        r = o.train_step(w, m, v, s + 1, X, Y, dn, icov, float(g["lr"]))
        assert abs(r["loss"] - g["losses"][s]) <= 2e-3 * abs(g["losses"][s]) + 1e-9, (s, r["loss"], g["losses"][s])
        if s == 0:
            np.testing.assert_allclose(r["loss_rows"], g["loss_rows"], rtol=2e-3, atol=1e-9)
            np.testing.assert_allclose(r["chisq_md"], g["chisqMd"], rtol=1e-4)
            np.testing.assert_allclose(r["chisq_nnd"], g["chisqnnd"], rtol=1e-3)
            gd = unflatten(r["grads"], o.shapes)
            if full:
                for k in keys:
                    ref = g["grad0_" + k]
                    assert np.max(np.abs(gd[k] - ref)) <= 2e-3 * np.max(np.abs(ref)) + 1e-12, k
            else:
                norms = np.array([np.linalg.norm(gd[k].astype(np.float64)) for k in keys])
                np.testing.assert_allclose(norms, g["grad0_norm"], rtol=2e-3)
                ref = g["grad0_layer1"]
                assert np.max(np.abs(gd["layer1.weight"] - ref)) <= 2e-3 * np.max(np.abs(ref))
    wd = unflatten(w, o.shapes)
    if full:
        for k in keys:
            ref = g["final_" + k]
            assert np.max(np.abs(wd[k] - ref)) <= 1e-4 * max(np.max(np.abs(ref)), 1e-3) + 2e-5, k
    else:
        ref = g["final_layer1"]
        assert np.max(np.abs(wd["layer1.weight"] - ref)) <= 1e-4 * np.max(np.abs(ref)) + 2e-5
