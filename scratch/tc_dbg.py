import sys, os, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from linna_b200 import engine, synthetic
from tests.helpers import load_golden, problem_from_golden, lnp_tol
names = sys.argv[1:] or ["tiny", "c1", "c3s", "c3mix", "c4s", "ypos", "simple"]
for name in names:
    g = load_golden(name); p = problem_from_golden(g)
    e = engine.engine_from_problem(p); e.set_path("tc")
    ud = torch.from_numpy(g["u"]).cuda()
    try:
        got = e.lnp(ud).cpu().numpy().astype(np.float64)
    except Exception as ex:
        print(name, "LNP FAILED", ex); continue
    err = got - g["f64_lnp"]
    print(name, "SEG", os.environ.get("LINNA_TC_SEG_KC"), "lnp: mean rel err %.3e max|err| %.3e max rel %.3e | ref f32 max|err| %.3e tol %.3e" % (
        (err / np.abs(g["f64_lnp"])).mean(), np.abs(err).max(), np.abs(err / g["f64_lnp"]).max(),
        np.abs(g["f32_lnp"] - g["f64_lnp"]).max(), lnp_tol(g["f64_lnp"]).max()), flush=True)
    try:
        l2, gr = e.lnp_grad(ud)
        gr = gr.cpu().numpy().astype(np.float64); ref = g["f64_grad"]
        rel = np.max(np.abs(gr - ref), axis=1) / np.max(np.abs(ref), axis=1)
        relf = np.max(np.abs(g["f32_grad"] - ref), axis=1) / np.max(np.abs(ref), axis=1)
        print(name, "grad rel err %.3e (ref f32 %.3e) lnp equal %s" % (rel.max(), relf.max(), np.array_equal(l2.cpu().numpy(), got.astype(np.float32))), flush=True)
    except Exception as ex:
        print(name, "GRAD FAILED", ex)
