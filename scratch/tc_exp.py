import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine
from tests.helpers import load_golden, problem_from_golden, lnp_tol
for name in ("c3s", "c4s", "c3mix"):
    g = load_golden(name); p = problem_from_golden(g)
    e = engine.engine_from_problem(p); e.set_path("tc")
    got = e.lnp(torch.from_numpy(g["u"]).cuda()).cpu().numpy().astype(np.float64)
    err = got - g["f64_lnp"]
    print(name, "SEG", os.environ.get("LINNA_TC_SEG_KC"), "mean err %.3e max|err| %.3e  ref f32 max|err| %.3e tol %.3e" % (err.mean(), np.abs(err).max(), np.abs(g["f32_lnp"]-g["f64_lnp"]).max(), lnp_tol(g["f64_lnp"]).max()))
