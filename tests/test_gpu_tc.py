"""GPU tests of the EXPERIMENTAL tensor-core (tcgen05 / TMEM / TMA, 3xTF32) lnP kernel.

The kernel is opt-in (``Engine.set_path("tc")``): the tensor core truncates its fp32 accumulator on
every tcgen05.mma, which leaves a systematic toward-zero bias proportional to the number of
instructions accumulated in tensor memory (measured here: 6e-6 of chi^2 with plain accumulation,
1.5e-6 with the default two-level scheme, 4e-7 when draining every k-chunk).  That is outside the
1e-4-absolute / 4-ulp bar of the FP32 FFMA kernel for |lnL| >~ 100, so the default path stays FFMA and
these tests hold the tensor-core kernel to the RELATIVE bar  |d lnL| <= 4e-6 |lnL| + 1e-4  instead."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, engine, synthetic
from oracle.oracle import Oracle
from tests.helpers import fixture_problem, lnp_tol, load_golden, problem_from_golden

pytestmark = pytest.mark.gpu


def tc_tol(lnp):
    return 4e-6 * np.abs(lnp) + 1e-4


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def test_tc_needs_cholesky_form():
    g = load_golden("tiny")
    e = engine.engine_from_problem(problem_from_golden(g), quad="dense")
    e.set_path("tc")
    with pytest.raises(engine.LinnaError, match="Cholesky"):
        e.lnp(_dev(g["u"]))
    e.set_path("auto", tc_min_rows=1)          # automatic selection falls back to the FFMA kernel
    assert np.all(np.isfinite(e.lnp(_dev(g["u"])).cpu().numpy()))


@pytest.mark.parametrize("name", ["tiny", "c1", "simple", "ypos", "c3s", "c3mix", "c4s"])
def test_tc_lnp_vs_reference_golden(name):
    g = load_golden(name)
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p, quad="chol")
    e.set_path("ffma")
    ref_ffma = e.lnp(_dev(g["u"])).cpu().numpy()
    e.set_path("tc")
    got = e.lnp(_dev(g["u"])).cpu().numpy()
    err = np.abs(got.astype(np.float64) - g["f64_lnp"])
    tol = tc_tol(g["f64_lnp"])
    assert np.all(err <= tol), (err.max(), tol.max(), got[:4], g["f64_lnp"][:4])
    assert np.all(np.abs(got - ref_ffma) <= tol)


def test_tc_fixture_and_ragged():
    g = load_golden("fixture")
    p = fixture_problem(g)
    e = engine.engine_from_problem(p)
    e.set_path("tc")
    got = e.lnp(_dev(g["u"])).cpu().numpy()
    assert np.all(np.abs(got - g["lnp"]) <= tc_tol(g["lnp"]))
    o = Oracle(p, arch)
    for n in (1, 127, 128, 129, 1000):
        u = synthetic.walkers(n, 2, scale=1.0, seed=n)
        ref = o.lnp(u, np.float64)["lnp"]
        out = e.lnp(_dev(u)).cpu().numpy()
        assert out.shape == (n,) and np.all(np.abs(out - ref) <= tc_tol(ref)), n
    u = synthetic.walkers(300, 2, seed=3)
    u[7, 1] = np.nan
    out = e.lnp(_dev(u)).cpu().numpy()
    assert np.isneginf(out[7]) and np.all(np.isfinite(np.delete(out, 7)))


def test_tc_full_size_c3():
    p = synthetic.make_problem(30, 500, seed=0)
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    n = 100000
    u = synthetic.walkers(n, 30, scale=0.3, seed=1)
    ud = _dev(u)
    e.set_path("tc")
    a = e.lnp(ud).cpu().numpy()
    b = e.lnp(ud).cpu().numpy()
    assert np.array_equal(a, b)
    e.set_path("ffma")
    f = e.lnp(ud).cpu().numpy()
    assert np.all(np.abs(a - f) <= tc_tol(f)), np.abs(a - f).max()
    idx = np.random.default_rng(0).choice(n, 64, replace=False)
    ref = Oracle(p, arch).lnp(u[idx], np.float64)["lnp"]
    assert np.all(np.abs(a[idx] - ref) <= tc_tol(ref)), np.abs(a[idx] - ref).max()
