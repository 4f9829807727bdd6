"""CPU-only checks of the drop-in boundary: reference pickles resolve through the `linna` shim,
host-side transforms match the reference goldens, the C-ABI library loads and exports every symbol
the header declares, and the product fails loudly without a GPU."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, load_golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    from linna_b200 import engine
    hdr = open(os.path.join(REPO, "include", "linna_b200.h")).read()
    names = set(re.findall(r"\b(linna_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 14
    lib = ctypes.CDLL(engine.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert lib.linna_abi_version() == 1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from linna_b200 import engine, synthetic
    p = synthetic.make_problem(3, 4, seed=8)
    with pytest.raises(engine.LinnaError, match="no CUDA device"):
        engine.engine_from_problem(p, with_likelihood=False)
    import linna.nn as N
    m = N.ChtoModelv2(3, 4, None)
    with pytest.raises(RuntimeError, match="GPU only"):
        m(torch.zeros(2, 3))


def test_shim_module_paths_and_pickles(tmp_path):
    import linna
    import linna.nn
    import linna.util as U
    assert U.X_transform_class.__module__ == "linna.util"
    xt = U.X_transform_class(torch.zeros(3), torch.ones(3), "cpu", [0])
    xt.pickle(str(tmp_path / "X_transform.pkl"))
    raw = open(tmp_path / "X_transform.pkl", "rb").read()
    assert b"linna.util" in raw and b"linna_b200" not in raw, "pickles must name the reference module path"
    back = U._unpickle(str(tmp_path / "X_transform.pkl"))
    assert isinstance(back, U.X_transform_class) and back.dolog10index == [0] and back.dev == "cpu"
    m = linna.nn.ChtoModelv2(2, 2, None)
    assert pickle.loads(pickle.dumps(type(m))) is linna.nn.ChtoModelv2
    assert pickle.loads(pickle.dumps(U.train_NN)) is U.train_NN


def test_transforms_match_reference_goldens():
    import linna.util as U
    g = load_golden("c3mix")
    pri = [dict(dist=str(d), arg1=float(a), arg2=float(b)) for d, a, b in zip(g["prior_dist"], g["prior_arg1"], g["prior_arg2"])]
    k = g["f32_theta"].shape[0]
    th = U.Transform(pri)(g["u"][:k])
    np.testing.assert_allclose(th, g["f32_theta"], rtol=2e-6, atol=1e-6)
    back = U.invTransform(pri)(th.astype(np.float64))
    np.testing.assert_allclose(back, g["u"][:k], rtol=2e-3, atol=2e-4)
    # diagonal transforms: y -> m
    yt = U.Y_transform_class(torch.tensor(g["y_mean"]), torch.tensor(g["y_std"]), "cpu")
    yi = U.Y_invtransform_data(g["sigma"], "cpu")
    m = yi(yt(torch.tensor(g["f32_yhat"]))).detach().numpy()
    np.testing.assert_allclose(m, g["f32_m"], rtol=1e-6, atol=1e-6)
    xt = U.X_transform_class(torch.tensor(g["X_mean"]), torch.tensor(g["X_std"]), "cpu", list(g["dolog10index"]))
    xh = xt(torch.tensor(g["f32_theta"])).numpy()
    assert np.all(np.isfinite(xh))


def test_fixture_state_dict_loads_into_model():
    import linna.nn as N
    g = load_golden("fixture")
    m = N.ChtoModelv2(2, 2, None)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith("sd_")}
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)


def test_weight_init_matches_reference():
    import linna.nn as N
    g = dict(np.load(os.path.join(GOLDEN, "init.npz")))
    for kind, n_in, n_out in (("ChtoModelv2", 2, 2), ("ChtoModelv2", 5, 40), ("ChtoModelsimple", 3, 4),
                              ("ChtoModelv2_linear", 3, 4)):
        torch.manual_seed(1234)
        m = getattr(N, kind)(n_in, n_out, None)
        tag = "%s_%d_%d" % (kind, n_in, n_out)
        sd = m.state_dict()
        assert list(sd.keys()) == [str(k) for k in g[tag + "_keys"]]
        np.testing.assert_allclose([float(v.double().sum()) for v in sd.values()], g[tag + "_sum"], atol=1e-9)
        np.testing.assert_allclose([float(v.double().abs().sum()) for v in sd.values()], g[tag + "_abssum"], atol=1e-9)


def test_early_stopping_codes():
    from linna.predictor_gpu import EarlyStopping
    es = EarlyStopping(patience=10)
    assert es.step(1.0, 1.0) == 0
    assert es.step(0.5, 0.5) == 0           # improvement
    codes = [es.step(0.6, 0.4) for _ in range(12)]
    assert 1 in codes                        # asks for a cooler learning rate near the patience limit
    es2 = EarlyStopping(patience=3)
    es2.step(1.0, 1.0)
    assert [es2.step(2.0, 1.0) for _ in range(3)][-1] == 2


def test_loss_constants_match_oracle_setup():
    """Auxilleryfunc.__init__ (host, float64) against the oracle's restatement, incl. ypositive."""
    import linna.util as U
    from linna_b200 import synthetic
    from oracle.oracle import normalised_loss_constants
    for ypos in (False, True):
        p = synthetic.make_problem(4, 6, seed=4, ypositive=ypos)
        p.data = np.abs(np.random.default_rng(1).standard_normal(6)) + 0.5
        sig = np.asarray(p.sigma, np.float32)
        ytd = U.Y_transform_data(sig, "cpu")
        dt = torch.tensor(p.data.astype(np.float32))
        yinv = U.Y_invtransform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), dt, "cpu", ypositive=ypos)
        aux = U.Loss_fn(dt, torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinv, "cpu").auxileryfunction
        dh, ic, sg, ym, ys, yp = aux.constants()
        dn, icov = normalised_loss_constants(p.cov, sig, p.y_mean, p.y_std, p.data, ypositive=ypos)
        assert dh.shape == (6,) and ic.shape == (6, 6) and sg.shape == (6,) and yp == ypos
        np.testing.assert_allclose(dh, dn, rtol=1e-6)
        np.testing.assert_allclose(ic, icov, rtol=1e-5, atol=1e-6)
