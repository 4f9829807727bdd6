"""Shared helpers for the parity tests: golden loading and tolerances."""
import ast
import os

import numpy as np

from linna_b200 import arch, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def problem_from_golden(g):
    """Re-create the synthetic Problem of a golden file (weights regenerated from the seed and
    pinned by the stored checksums; data vector taken from the file)."""
    kw = ast.literal_eval(str(g["make_kwargs"]))
    p = synthetic.make_problem(**kw)
    chk = np.array([float(np.sum(v.astype(np.float64))) for v in p.state_dict.values()])
    np.testing.assert_allclose(chk, g["w_checksum"], rtol=0, atol=1e-9)
    p.data = g["data"].astype(np.float64)
    assert np.array_equal(p.X_mean, g["X_mean"]) and np.array_equal(p.y_std, g["y_std"])
    return p


def fixture_problem(g):
    """The reference's shipped 2-D fixture (tests/test_data/2dgaussian_Fulltconn/iter_0) as a Problem."""
    p = synthetic.Problem()
    p.kind, p.n_in, p.n_out = "ChtoModelv2", 2, 2
    p.state_dict = {k[3:]: g[k] for k in g if k.startswith("sd_")}
    p.priors = [dict(param="x%d" % i, dist="flat", arg1=-2.0, arg2=2.0) for i in range(2)]
    p.X_mean, p.X_std, p.y_mean, p.y_std = g["X_mean"], g["X_std"], g["y_mean"], g["y_std"]
    p.sigma = g["sigma"].astype(np.float64)
    p.cov, p.inv_cov, p.data = g["cov"], g["inv_cov"], g["data"].astype(np.float64)
    p.temperature = 1.0
    return p


def ulp32(x):
    return np.spacing(np.abs(np.asarray(x, np.float32))).astype(np.float64)


def lnp_tol(lnp_ref):
    """|d lnL| <= max(1e-4, 4 ulp_f32(|lnL|)): 1e-4 absolute (BASELINE.json north_star) is below
    float32 resolution once |lnL| >~ 1000 and the reference itself returns float32."""
    return np.maximum(1e-4, 4 * ulp32(lnp_ref))


def rel_inf(a, b):
    """max over rows of ||a-b||_inf / ||b||_inf."""
    a, b = np.atleast_2d(a).astype(np.float64), np.atleast_2d(b).astype(np.float64)
    return float(np.max(np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)))
