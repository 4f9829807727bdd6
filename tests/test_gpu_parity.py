"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C ABI,
against the reference-generated golden vectors and the CPU oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, engine, synthetic
from oracle.oracle import Oracle
from tests.helpers import fixture_problem, lnp_tol, load_golden, problem_from_golden, rel_inf

pytestmark = pytest.mark.gpu

SYNTH = ["c1", "c3s", "c3mix", "ypos", "c4s", "simple", "v2lin", "tiny"]


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def _check_case(g, p, quad, rows, path="auto"):
    e = engine.engine_from_problem(p, quad=quad)
    e.set_tile_rows(rows)
    e.set_path(path)
    u = g["u"] if "u" in g else None
    lnp, grad = e.lnp_grad(_dev(u))
    lnp_only = e.lnp(_dev(u))
    torch.cuda.synchronize()
    lnp, grad, lnp_only = lnp.cpu().numpy(), grad.cpu().numpy(), lnp_only.cpu().numpy()
    return e, lnp, grad, lnp_only


@pytest.mark.parametrize("rows", [0, 8, 16, 32])
@pytest.mark.parametrize("quad", ["chol", "dense"])
def test_fixture(quad, rows):
    g = load_golden("fixture")
    p = fixture_problem(g)
    e, lnp, grad, lnp_only = _check_case(g, p, quad, rows)
    assert np.all(np.abs(lnp - g["lnp"]) <= lnp_tol(g["lnp"])), np.abs(lnp - g["lnp"]).max()
    assert np.array_equal(lnp, lnp_only), "forward-only and forward+backward programs must agree bit for bit"
    assert rel_inf(grad, g["grad"]) < 2e-5
    # SURVEY 8c known answers
    assert abs(lnp[0] - (-3.36370969)) < 1e-5
    np.testing.assert_allclose(grad[0], [-0.53383291, 0.99314642], atol=3e-6)
    # Predictor.predict rows (physical parameters in, y out)
    y = e.predict(_dev(g["theta"]), engine.LINNA_OUT_Y).cpu().numpy()
    m = e.predict(_dev(g["theta"]), engine.LINNA_OUT_M).cpu().numpy()
    assert rel_inf(y, g["y"]) < 1e-5 and rel_inf(m, g["m"]) < 1e-5
    np.testing.assert_allclose(y[0], [0.10346876, -0.31955421], atol=1e-6)


@pytest.mark.parametrize("path", ["ffma", "cluster"])
@pytest.mark.parametrize("name", SYNTH)
@pytest.mark.parametrize("quad", ["chol", "dense"])
def test_synthetic_vs_reference_golden(name, quad, path):
    """Both FP32 kernels -- the tiled FFMA kernel and the small-batch cluster kernel (columns of every layer split over
    a thread-block cluster, activations in distributed shared memory) -- against the reference's own outputs."""
    g = load_golden(name)
    p = problem_from_golden(g)
    try:
        e, lnp, grad, lnp_only = _check_case(g, p, quad, 0, path)
    except engine.LinnaError as ex:
        if path == "cluster" and "cluster kernel unavailable" in str(ex) and p.n_out >= 1000:
            pytest.skip("the activation arena of this shape does not fit in shared memory: auto mode serves it with the FFMA kernel")
        raise
    assert e.last_kernel() == path
    k = g["f32_m"].shape[0]
    theta = g["f32_theta"]
    m = e.predict(_dev(theta), engine.LINNA_OUT_M).cpu().numpy()
    y = e.predict(_dev(theta), engine.LINNA_OUT_Y).cpu().numpy()
    yhat = e.predict(_dev(theta), engine.LINNA_OUT_YHAT).cpu().numpy()
    # north star: 1e-5 relative on predicted data vectors (vs the reference's own float32 output)
    assert rel_inf(m, g["f32_m"]) < 1e-5
    assert rel_inf(y, g["f32_y"]) < 1e-5
    assert rel_inf(yhat, g["f32_yhat"]) < 1e-5
    # north star: 1e-4 absolute on lnL, floored at float32 resolution; judged against the float64 run
    err = np.abs(lnp.astype(np.float64) - g["f64_lnp"])
    err_ref = np.abs(g["f32_lnp"] - g["f64_lnp"])
    tol = lnp_tol(g["f64_lnp"])
    assert np.all(err <= np.maximum(tol, 2 * err_ref)), (err.max(), err_ref.max(), tol.max())
    assert np.array_equal(lnp, lnp_only)
    assert rel_inf(grad, g["f64_grad"]) < 2e-4, rel_inf(grad, g["f64_grad"])
    assert rel_inf(grad, g["f64_grad"]) <= max(5 * rel_inf(g["f32_grad"], g["f64_grad"]), 2e-5)


@pytest.mark.parametrize("name", ["c3s", "c4s", "tiny"])
@pytest.mark.parametrize("rows", [8, 16, 32])
def test_tile_variants_agree(name, rows):
    g = load_golden(name)
    p = problem_from_golden(g)
    e, lnp, grad, _ = _check_case(g, p, "chol", rows)
    err = np.abs(lnp.astype(np.float64) - g["f64_lnp"])
    assert np.all(err <= np.maximum(lnp_tol(g["f64_lnp"]), 2 * np.abs(g["f32_lnp"] - g["f64_lnp"])))
    assert rel_inf(grad, g["f64_grad"]) < 2e-4


def test_host_buffer_entry_points_match_device():
    g = load_golden("c3s")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    u = g["u"]
    l_dev, g_dev = e.lnp_grad(_dev(u))
    l_host, g_host = e.lnp_grad(u)
    assert np.array_equal(l_dev.cpu().numpy(), l_host) and np.array_equal(g_dev.cpu().numpy(), g_host)
    assert np.array_equal(e.lnp(u), l_host)
    th = g["f32_theta"]
    assert np.array_equal(e.predict(th), e.predict(_dev(th)).cpu().numpy())


@pytest.mark.parametrize("n", [9000, 23681, 60000])
def test_pipelined_host_buffers_match_device(n):
    """linna_lnp_host / linna_lnp_grad_host on batches large enough to be cut into chunks (one chunk below 1.25 kernel
    rounds, then 1 + 2 + ... rounds; pieces of 512 KB staged by the caller and the helper thread): pageable and pinned
    arrays, every combination, must return exactly what the device-resident call returns -- a walker's value does not
    depend on the chunk it travels in."""
    g = load_golden("c3s")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    u = synthetic.walkers(n, p.n_in, scale=0.3, seed=n)
    l_dev, g_dev = e.lnp_grad(_dev(u))
    l_dev, g_dev = l_dev.cpu().numpy(), g_dev.cpu().numpy()
    for rep in range(2):                      # the second call reuses the staging buffers and the helper thread
        l_host, g_host = e.lnp_grad(u)
        assert np.array_equal(l_dev, l_host) and np.array_equal(g_dev, g_host)
        assert np.array_equal(e.lnp(u), l_dev)
    pin = torch.from_numpy(u).pin_memory().numpy()
    out_l = torch.empty(n, dtype=torch.float32).pin_memory().numpy()
    out_g = torch.empty(n, p.n_in, dtype=torch.float32).pin_memory().numpy()
    e.lnp_grad(pin, out=out_l, out_grad=out_g)                 # pinned in, pinned out: no staging at all
    assert np.array_equal(out_l, l_dev) and np.array_equal(out_g, g_dev)
    out_l[:] = 0
    e.lnp(u, out=out_l)                                        # pageable in, pinned out
    assert np.array_equal(out_l, l_dev)
    assert np.array_equal(e.lnp(pin), l_dev)                   # pinned in, pageable out
    # a smaller batch right after a larger one (staging buffers larger than the call)
    assert np.array_equal(e.lnp(u[:4097]), l_dev[:4097])


@pytest.mark.parametrize("path", ["auto", "ffma", "cluster"])
def test_ragged_empty_and_nan(path):
    g = load_golden("tiny")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    e.set_path(path)
    assert e.lnp(np.zeros((0, p.n_in), np.float32)).shape == (0,)
    o = Oracle(p, arch)
    for n in (1, 7, 8, 9, 33, 257):
        u = synthetic.walkers(n, p.n_in, scale=1.0, seed=n)
        ref = o.lnp(u, np.float64)["lnp"]
        got = e.lnp(u)
        assert got.shape == (n,)
        assert np.all(np.abs(got - ref) <= lnp_tol(ref)), n
    u = synthetic.walkers(5, p.n_in, seed=3)
    u[2, 1] = np.nan
    got = e.lnp(u)
    assert np.isneginf(got[2]) and np.all(np.isfinite(np.delete(got, 2))), "NaN -> -inf (util.py:1015-1016)"


def test_cluster_kernel_is_the_small_batch_default_and_deterministic():
    """Auto mode: batches below one tensor-core walker pair run on the cluster kernel; a row's value does not depend on
    its position in the batch, on the batch size or on the launch (fixed summation order), and the three kernels agree."""
    g = load_golden("c3s")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    u = synthetic.walkers(60, p.n_in, scale=0.5, seed=4)
    a = e.lnp(_dev(u)).cpu().numpy()
    assert e.last_kernel() == "cluster"
    assert np.array_equal(a, e.lnp(_dev(u)).cpu().numpy())
    perm = np.random.default_rng(1).permutation(60)
    assert np.array_equal(e.lnp(_dev(u[perm])).cpu().numpy(), a[perm])
    assert np.array_equal(e.lnp(_dev(u[:5])).cpu().numpy(), a[:5])
    l2, g2 = e.lnp_grad(_dev(u))
    assert e.last_kernel() == "cluster" and np.array_equal(l2.cpu().numpy(), a)
    e.set_path("ffma")
    lf, gf = e.lnp_grad(_dev(u))
    assert e.last_kernel() == "ffma"
    assert np.all(np.abs(lf.cpu().numpy() - a) <= lnp_tol(a))
    assert rel_inf(g2.cpu().numpy(), gf.cpu().numpy()) < 2e-5
    ref = Oracle(p, arch).lnp(u, np.float64, grad=True)
    assert np.all(np.abs(a - ref["lnp"]) <= lnp_tol(ref["lnp"]))
    assert rel_inf(g2.cpu().numpy(), ref["grad"]) < 2e-4
    e.set_path("cluster")   # forced: any batch size walks the tiles with the clusters the device schedules
    ub = synthetic.walkers(3001, p.n_in, scale=0.5, seed=5)
    refb = Oracle(p, arch).lnp(ub[::50], np.float64)
    got = e.lnp(_dev(ub)).cpu().numpy()
    assert e.last_kernel() == "cluster" and np.all(np.abs(got[::50] - refb["lnp"]) <= lnp_tol(refb["lnp"]))
    th = g["f32_theta"]
    assert rel_inf(e.predict(_dev(th), engine.LINNA_OUT_M).cpu().numpy(), g["f32_m"]) < 1e-5 and e.last_kernel() == "cluster"


def test_full_size_c3_properties():
    """BASELINE config C3 at its full batch (1e5 walkers): determinism, batch-composition
    independence and a sampled comparison with the oracle."""
    p = synthetic.make_problem(30, 500, seed=0)
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    e.set_path("ffma")          # this test is about the FP32 kernel's tiling; tests/test_gpu_tc.py has the tensor-core twin
    n = 100000
    u = synthetic.walkers(n, 30, scale=0.3, seed=1)
    ud = _dev(u)
    a = e.lnp(ud).cpu().numpy()
    b = e.lnp(ud).cpu().numpy()
    assert np.array_equal(a, b), "two launches on the same input must agree bit for bit"
    idx = np.sort(np.random.default_rng(0).choice(94720, 96, replace=False))   # rows of the 32-row waves
    # a row's value does not depend on what else is in the batch (same tile height => same bits;
    # other tile heights split the k-sums differently => float32 rounding only)
    e.set_tile_rows(32)
    sub = e.lnp(_dev(u[idx])).cpu().numpy()
    assert np.array_equal(sub, a[idx])
    for rows in (8, 16):
        e.set_tile_rows(rows)
        sub = e.lnp(_dev(u[idx])).cpu().numpy()
        assert np.all(np.abs(sub - a[idx]) <= lnp_tol(a[idx]))
    e.set_tile_rows(0)
    ref = Oracle(p, arch).lnp(u[idx], np.float64, grad=True)
    assert np.all(np.abs(a[idx] - ref["lnp"]) <= lnp_tol(ref["lnp"]))
    l2, g2 = e.lnp_grad(ud)
    assert np.array_equal(l2.cpu().numpy(), a)
    assert rel_inf(g2.cpu().numpy()[idx], ref["grad"]) < 2e-4
    # chi^2 ~ n_out near the fiducial point: sanity on magnitudes
    assert 150 < -np.median(a) < 600


@pytest.mark.parametrize("name", ["c1", "c3s", "c3mix", "c4s", "tiny", "simple"])
def test_folded_tail_matches_unfolded_and_reference(name):
    """lnP folds the last layer + inverse transform + Cholesky product into one GEMM (float64-formed
    operand).  Both orders must satisfy the reference bar; predict() always runs unfolded."""
    g = load_golden(name)
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    out = {}
    for fold in (True, False):
        e.set_fold(fold)
        lnp, grad = e.lnp_grad(_dev(g["u"]))
        out[fold] = (lnp.cpu().numpy(), grad.cpu().numpy(), e.lnp(_dev(g["u"])).cpu().numpy())
        err = np.abs(out[fold][0].astype(np.float64) - g["f64_lnp"])
        tol = lnp_tol(g["f64_lnp"])
        assert np.all(err <= np.maximum(tol, 2 * np.abs(g["f32_lnp"] - g["f64_lnp"]))), (fold, err.max())
        assert np.array_equal(out[fold][0], out[fold][2])
        assert rel_inf(out[fold][1], g["f64_grad"]) < 2e-4
    assert np.all(np.abs(out[True][0] - out[False][0]) <= lnp_tol(out[False][0]))
