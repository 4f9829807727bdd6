"""GPU tests of the tensor-core (tcgen05 / TMEM / TMA, split-fp16) lnP and lnP+gradient kernel.

The kernel serves large batches (``Engine.set_path("auto")`` sends batches >= tc_min_rows to it,
``"tc"`` forces it).  Every operand is split into two fp16 halves (22 significant bits) and the
tensor core's truncating fp32 accumulator is drained into round-to-nearest register accumulators
every few k-chunks, so the result tracks the FP32 FFMA kernel to a few ulp; the bars below are
written against a float64 run of the reference modules (tests/golden) / the float64 oracle:

    |d lnL|  <= max(1e-4, 4 ulp_f32(|lnL|))   -- tests.helpers.lnp_tol, the SAME bar the exact-FP32 kernel is held to
                                            (1e-4 absolute is north_star's bar; below |lnL| ~ 400 it is the binding
                                            term, above it float32 resolution of lnL itself is: the reference returns
                                            float32 and is itself 0.6 - 0.75 of this bar away from float64 on these
                                            goldens, profiles/r2_parity_probe.txt)
    gradient: relative infinity-norm error < 2e-4 (same bar as the FFMA kernel)

Rows in which an activation leaves the fp16 range come out NaN on the tensor-core path (the relu clamp propagates
NaN) and are recomputed by the FP32 kernel in a fix-up launch: `test_tc_fp16_overflow_rows_are_fixed_up`.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from linna_b200 import arch, engine, synthetic
from oracle.oracle import Oracle
from tests.helpers import fixture_problem, lnp_tol, load_golden, problem_from_golden, rel_inf

pytestmark = pytest.mark.gpu


def tc_tol(lnp):
    return lnp_tol(lnp)


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


@pytest.mark.parametrize("name", ["tiny", "c1", "c3s", "c3mix", "c4s"])
def test_tc_grad_vs_reference_golden(name):
    g = load_golden(name)
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p, quad="chol")
    e.set_path("tc")
    lnp, grad = e.lnp_grad(_dev(g["u"]))
    lnp, grad = lnp.cpu().numpy(), grad.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(lnp - g["f64_lnp"]) <= tc_tol(g["f64_lnp"]))
    ref = g["f64_grad"]
    rel = np.max(np.abs(grad - ref), axis=1) / np.max(np.abs(ref), axis=1)
    assert rel.max() < 2e-4, rel.max()


def test_tc_fixture_and_ragged():
    g = load_golden("fixture")
    p = fixture_problem(g)
    e = engine.engine_from_problem(p)
    e.set_path("tc")
    got = e.lnp(_dev(g["u"])).cpu().numpy()
    assert np.all(np.abs(got - g["lnp"]) <= tc_tol(g["lnp"]))
    o = Oracle(p, arch)
    for n in (1, 127, 128, 129, 255, 256, 257, 1000):
        u = synthetic.walkers(n, 2, scale=1.0, seed=n)
        ref = o.lnp(u, np.float64, grad=True)
        out = e.lnp(_dev(u)).cpu().numpy()
        assert out.shape == (n,) and np.all(np.abs(out - ref["lnp"]) <= tc_tol(ref["lnp"])), n
        l2, gr = e.lnp_grad(_dev(u))
        assert np.array_equal(l2.cpu().numpy(), out), n
        gr = gr.cpu().numpy()
        assert gr.shape == (n, 2)
        assert np.max(np.abs(gr - ref["grad"])) <= 2e-4 * np.max(np.abs(ref["grad"])), n
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
    la, ga = e.lnp_grad(ud)
    assert np.array_equal(la.cpu().numpy(), a)
    # a walker's value does not depend on what else is in the batch, nor on where in the batch it sits
    idx = np.sort(np.random.default_rng(1).choice(n, 3000, replace=False))
    assert np.array_equal(e.lnp(_dev(u[idx])).cpu().numpy(), a[idx])
    e.set_path("ffma")
    f = e.lnp(ud).cpu().numpy()
    # two float32 results that are each within the bar of the float64 value can be two bars apart
    assert np.all(np.abs(a - f) <= 2 * tc_tol(f)), np.abs(a - f).max()
    lf, gf = e.lnp_grad(ud)
    ga, gf = ga.cpu().numpy(), gf.cpu().numpy()
    # The gradient is discontinuous where a relu pre-activation crosses zero: a unit whose pre-activation is
    # within float32 rounding of 0 can come out on either side in two correct float32 evaluations (measured
    # against the float64 oracle, it is the FFMA kernel as often as this one).  ~4e8 relu units per batch make
    # that a few dozen rows; everything else has to agree to the usual bar.
    rel = np.max(np.abs(ga - gf), axis=1) / np.max(np.abs(gf), axis=1)
    assert np.median(rel) < 1e-5 and np.mean(rel > 2e-4) < 1e-3, (np.median(rel), np.mean(rel > 2e-4))
    idx = np.random.default_rng(0).choice(n, 2048, replace=False)
    ref = Oracle(p, arch).lnp(u[idx], np.float64)["lnp"]
    assert np.all(np.abs(a[idx] - ref) <= tc_tol(ref)), np.abs(a[idx] - ref).max()
    assert np.all(np.abs(f[idx] - ref) <= tc_tol(ref)), np.abs(f[idx] - ref).max()


def test_tc_grad_far_from_the_peak():
    """A data vector tens of sigma away from everything the emulator predicts (chi^2 ~ 1e6, |r| ~ 50 per element):
    the backward pass runs at unit scale per walker (a power of two taken from its own chi^2) so that no gradient
    leaves the fp16 range, and the result still tracks the float64 oracle."""
    p = synthetic.make_problem(30, 500, seed=0)
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    rng = np.random.default_rng(4)
    p.data = m0.astype(np.float64) + 60.0 * np.asarray(p.sigma, np.float64) * rng.choice([-1.0, 1.0], size=500)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    u = np.concatenate([synthetic.walkers(1200, 30, scale=s, seed=7 + i) for i, s in enumerate((0.3, 1.0, 2.0))])
    e.set_path("tc")
    lnp, grad = e.lnp_grad(_dev(u))
    lnp, grad = lnp.cpu().numpy(), grad.cpu().numpy().astype(np.float64)
    assert np.all(np.isfinite(lnp)) and np.all(np.isfinite(grad))
    assert -lnp.max() > 1e5                                   # far from the peak indeed
    idx = np.arange(0, 3600, 60)
    ref = Oracle(p, arch).lnp(u[idx], np.float64, grad=True)
    assert np.all(np.abs(lnp[idx] - ref["lnp"]) <= tc_tol(ref["lnp"]))
    rel = np.max(np.abs(grad[idx] - ref["grad"]), axis=1) / np.max(np.abs(ref["grad"]), axis=1)
    assert np.median(rel) < 1e-5 and np.mean(rel > 2e-4) <= 0.05, (np.median(rel), rel.max())
    e.set_path("ffma")
    lf, gf = e.lnp_grad(_dev(u))
    relf = np.max(np.abs(grad - gf.cpu().numpy()), axis=1) / np.max(np.abs(gf.cpu().numpy()), axis=1)
    assert np.median(relf) < 1e-5 and np.mean(relf > 2e-4) < 0.02


def test_tc_fp16_overflow_rows_are_fixed_up():
    """An input normalisation with a tiny X_std puts xhat (and the first activations) of some walkers beyond the fp16
    range (65 504): the reference, in float32, returns a finite lnP for them.  On the tensor-core path such a row turns
    into NaN, is flagged, and the FP32 kernel recomputes exactly those rows in the fix-up launch -- value and gradient
    agree with the float64 oracle like any other row, and the neighbours are untouched."""
    p = synthetic.make_problem(12, 40, seed=21)
    p.X_std = p.X_std.copy()
    p.X_std[3] = 2e-5
    # a unit Gaussian prior centred on 0 keeps theta_3 = u_3 exact: xhat_3 = u_3 / 2e-5 is well conditioned in float32
    # (a flat prior would compute theta = 10 Phi(u) - 5 with float32 cancellation and make the TEST ill-conditioned)
    p.priors[3] = dict(param="p3", dist="gauss", arg1=0.0, arg2=1.0)
    p.theta0 = p.theta0.copy()
    p.theta0[3] = 0.0
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    n = 3000
    u = synthetic.walkers(n, 12, scale=0.3, seed=2)
    u[:, 3] *= 1e-4                                     # most walkers: |xhat_3| <~ 5, inside the fp16 range
    hot = np.arange(5, n, 97)
    u[hot, 3] = np.linspace(1.6, 4.0, hot.size) * np.where(np.arange(hot.size) % 2, 1, -1)   # |xhat_3| = 0.8e5 .. 2e5
    o = Oracle(p, arch)
    ref = o.lnp(u, np.float64, grad=True)
    assert np.all(np.isfinite(ref["lnp"]))
    e.set_path("tc")
    l0 = engine.launch_count()
    lnp, grad = e.lnp_grad(_dev(u))
    lnp, grad = lnp.cpu().numpy(), grad.cpu().numpy().astype(np.float64)
    assert e.last_kernel() == "tc" and engine.launch_count() - l0 == 2     # tensor-core launch + fix-up launch
    assert np.all(np.isfinite(lnp)), np.where(~np.isfinite(lnp))[0][:10]
    err = np.abs(lnp - ref["lnp"]) / tc_tol(ref["lnp"])
    cold = np.setdiff1d(np.arange(n), hot)
    assert err[cold].max() <= 1.0, (err[cold].max(), cold[np.argmax(err[cold])])
    # the fixed-up rows sit at chi^2 ~ 1e8 .. 1e10: the FP32 kernel's own accuracy there (a few 1e-6 relative)
    assert np.all(np.abs(lnp[hot] - ref["lnp"][hot]) <= 5e-6 * np.abs(ref["lnp"][hot])), np.abs(lnp[hot] / ref["lnp"][hot] - 1).max()
    e.set_path("ffma")
    ffl, ffg = e.lnp_grad(_dev(u))                                # they ARE FP32-kernel rows (8-row tiles in the fix-up)
    ff, ffg = ffl.cpu().numpy()[hot].astype(np.float64), ffg.cpu().numpy()[hot].astype(np.float64)
    assert np.all(np.abs(ff - lnp[hot]) <= 2e-6 * np.abs(ff))
    assert np.max(np.max(np.abs(ffg - grad[hot]), axis=1) / np.max(np.abs(ffg), axis=1)) < 1e-4
    e.set_path("tc")
    rel = np.max(np.abs(grad - ref["grad"]), axis=1) / np.max(np.abs(ref["grad"]), axis=1)
    # (the fixed-up rows carry float32's own error at |xhat| ~ 1e5: a few per cent of a gradient dominated by 1/X_std)
    # 1 / X_std = 5e4 amplifies the relu kinks of the gradient along u_3 (two correct float32 evaluations can put a unit
    # whose pre-activation is within rounding of 0 on either side, cf. test_tc_full_size_c3): a few rows may differ more
    assert np.median(rel) < 1e-5 and np.percentile(rel[cold], 99) < 2e-4 and rel[hot].max() < 5e-2, (np.median(rel), rel[cold].max(), rel[hot].max())
    assert np.array_equal(e.lnp(_dev(u)).cpu().numpy(), lnp.astype(np.float32))
    # a NaN input is still -inf (util.py:1015-1016) after the fix-up
    u2 = u.copy()
    u2[11, 0] = np.nan
    out = e.lnp(_dev(u2)).cpu().numpy()
    assert np.isneginf(out[11]) and np.all(np.isfinite(np.delete(out, 11)))


@pytest.mark.parametrize("n_in,n_out", [(3, 37), (64, 40), (17, 129), (8, 257)])
def test_tc_odd_shapes(n_in, n_out):
    """Widths that are not multiples of anything: K / N padding, a 2-chunk last layer (257), the widest input the
    tensor-core prologue takes (64)."""
    p = synthetic.make_problem(n_in, n_out, seed=11, priors="mixed")
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, p.temperature)
    u = synthetic.walkers(300, n_in, scale=0.5, seed=5)
    ref = Oracle(p, arch).lnp(u, np.float64, grad=True)
    e.set_path("tc")
    lnp, grad = e.lnp_grad(_dev(u))
    assert e.last_kernel() == "tc"
    lnp, grad = lnp.cpu().numpy(), grad.cpu().numpy().astype(np.float64)
    assert np.all(np.abs(lnp - ref["lnp"]) <= tc_tol(ref["lnp"])), np.abs(lnp - ref["lnp"]).max()
    rel = np.max(np.abs(grad - ref["grad"]), axis=1) / np.max(np.abs(ref["grad"]), axis=1)
    assert np.median(rel) < 1e-5 and rel.max() < 2e-4, rel.max()
    assert np.array_equal(e.lnp(_dev(u)).cpu().numpy(), lnp)


def test_tc_falls_back_when_the_shape_is_not_supported():
    p = synthetic.make_problem(70, 20, seed=3)       # 70 input parameters: wider than the tensor-core prologue
    e = engine.engine_from_problem(p, with_likelihood=False)
    m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    p.set_data_from_prediction(m0)
    e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
    u = synthetic.walkers(2048, 70, scale=0.3, seed=1)
    ref = Oracle(p, arch).lnp(u[:32], np.float64)["lnp"]
    out = e.lnp(_dev(u)).cpu().numpy()                # automatic selection: FP32 kernel
    assert e.last_kernel() == "ffma" and np.all(np.abs(out[:32] - ref) <= tc_tol(ref))
    e.set_path("tc")
    with pytest.raises(engine.LinnaError, match="64 input"):
        e.lnp(_dev(u))


@pytest.mark.parametrize("seg_kc,slots", [(2, 2), (3, 2), (4, 2), (5, 1), (8, 2)])
def test_tc_segment_lengths_and_slots(monkeypatch, seg_kc, slots):
    """The producer and the MMA issuer decide alike how many k-chunks a stage of a narrow step takes (two, unless
    that would cross an accumulation segment) and all roles walk the four-entry tensor-memory ring alike: any
    segment length (odd ones included) and either number of walker pairs per cluster give the same answer.  A
    protocol mismatch would trap (bounded mbarrier waits) rather than hang."""
    monkeypatch.setenv("LINNA_TC_SEG_KC", str(seg_kc))
    monkeypatch.setenv("LINNA_TC_SLOTS", str(slots))
    g = load_golden("c3s")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p, quad="chol")
    e.set_path("tc")
    reps = 1600                                        # 76 800 rows = 300 walker pairs: several visits of every cluster
    u = np.tile(g["u"], (reps, 1))
    lnp, grad = e.lnp_grad(_dev(u))
    assert e.last_kernel() == "tc"
    lnp, grad = lnp.cpu().numpy(), grad.cpu().numpy().astype(np.float64)
    n = g["u"].shape[0]
    tol = tc_tol(g["f64_lnp"]) * (2.0 if seg_kc != 6 else 1.0)   # 6 is what ships; other lengths are protocol tests
    for r in (0, reps // 2, reps - 1):
        sl = slice(r * n, (r + 1) * n)
        assert np.all(np.abs(lnp[sl] - g["f64_lnp"]) <= tol), np.abs(lnp[sl] - g["f64_lnp"]).max()
        rel = np.max(np.abs(grad[sl] - g["f64_grad"]), axis=1) / np.max(np.abs(g["f64_grad"]), axis=1)
        assert rel.max() < 2e-4, rel.max()
    assert np.array_equal(lnp[:n], lnp[-n:])           # batch position does not matter


@pytest.mark.parametrize("name", ["c3s", "c4s", "ypos", "c1", "simple"])
def test_tc_predict_data_vectors(name):
    """Predictor.predict on the tensor-core kernel (PREDICT program: the unfolded last layer, the inverse output transform
    in its epilogue): north star's 1e-5 relative bar on the predicted data vectors, against the reference's own float32
    outputs -- the bar the tensor-core path could not be held to while it only produced lnP."""
    g = load_golden(name)
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    theta = g["f32_theta"]
    rep = int(np.ceil(600 / len(theta)))
    big = np.ascontiguousarray(np.tile(theta, (rep, 1)), np.float32)        # >= 256 rows: the tensor-core path
    for kind, want in ((engine.LINNA_OUT_M, g["f32_m"]), (engine.LINNA_OUT_Y, g["f32_y"]), (engine.LINNA_OUT_YHAT, g["f32_yhat"])):
        e.set_path("tc")
        got = e.predict(torch.from_numpy(big).cuda(), kind).cpu().numpy()
        assert e.last_kernel() == "tc"
        assert got.shape == (len(big), p.n_out)
        assert np.array_equal(got[:len(theta)], got[len(theta):2 * len(theta)])      # a row does not depend on its position
        assert rel_inf(got[:len(theta)], want) < 1e-5, (name, kind, rel_inf(got[:len(theta)], want))
        e.set_path("ffma")
        ref = e.predict(torch.from_numpy(big).cuda(), kind).cpu().numpy()
        assert rel_inf(got, ref) < 1e-5
    # auto mode: large predict batches take the tensor-core kernel, and a model without likelihood constants can use it too
    e.set_path("auto")
    e.predict(torch.from_numpy(big).cuda(), engine.LINNA_OUT_M)
    assert e.last_kernel() == "tc"
    e2 = engine.engine_from_problem(p, with_likelihood=False)
    got2 = e2.predict(torch.from_numpy(big).cuda(), engine.LINNA_OUT_M).cpu().numpy()
    assert e2.last_kernel() == "tc" and rel_inf(got2[:len(theta)], g["f32_m"]) < 1e-5


def test_tc_predict_overflow_rows_are_fixed_up():
    """A row whose activations leave the fp16 range comes out of the tensor-core predict program as NaN and is recomputed
    by the FP32 kernel in the fix-up launch: the caller sees the reference's finite numbers."""
    g = load_golden("c3s")
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p)
    theta = np.tile(g["f32_theta"], (40, 1)).astype(np.float32)[:300]
    theta[7] *= 4e4          # |xhat| ~ 1e5: the first layer overflows fp16
    theta[123] *= -3e4
    e.set_path("ffma")
    ref = e.predict(torch.from_numpy(theta).cuda(), engine.LINNA_OUT_M).cpu().numpy()
    e.set_path("tc")
    got = e.predict(torch.from_numpy(theta).cuda(), engine.LINNA_OUT_M).cpu().numpy()
    assert e.last_kernel() == "tc"
    assert np.all(np.isfinite(ref[7])) and np.array_equal(got[7], ref[7]) and np.array_equal(got[123], ref[123])
    keep = np.ones(300, bool)
    keep[[7, 123]] = False
    assert rel_inf(got[keep], ref[keep]) < 1e-5
