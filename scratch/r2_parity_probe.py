"""Round-2 parity probe (GPU box): prints the measured errors the test bars are set from.

  * tensor-core lnP error / helpers.lnp_tol for every golden, and over 10^5 C3 walkers against the float64 oracle
    (a histogram of |d lnL| / tol goes to gpurun_out/r2_tc_error_hist.json)
  * training step: CUDA vs the reference's float32 goldens and vs the float64 run of the same reference modules
  * Ddlnp vs the reference's double-backward Hessian goldens
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linna_b200 import arch, engine, synthetic  # noqa: E402
from oracle.oracle import Oracle, flatten_state_dict, normalised_loss_constants, unflatten  # noqa: E402
from tests.helpers import load_golden, lnp_tol, problem_from_golden  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


out = {}
for name in ["tiny", "c1", "simple", "ypos", "c3s", "c3mix", "c4s"]:
    g = load_golden(name)
    p = problem_from_golden(g)
    e = engine.engine_from_problem(p, quad="chol")
    e.set_path("tc")
    got = e.lnp(dev(g["u"])).cpu().numpy().astype(np.float64)
    e.set_path("ffma")
    ff = e.lnp(dev(g["u"])).cpu().numpy().astype(np.float64)
    tol = lnp_tol(g["f64_lnp"])
    r_tc = np.abs(got - g["f64_lnp"]) / tol
    r_ff = np.abs(ff - g["f64_lnp"]) / tol
    r_ref = np.abs(g["f32_lnp"] - g["f64_lnp"]) / tol
    print("%-6s |lnL| %.0f  err/tol: tc %.3f  ffma %.3f  reference-f32 %.3f" % (name, np.abs(g["f64_lnp"]).max(), r_tc.max(), r_ff.max(), r_ref.max()))
    out[name] = dict(tc=float(r_tc.max()), ffma=float(r_ff.max()), ref32=float(r_ref.max()))
    # Hessian
    if "f64_hess" in g and p.n_in <= 64:
        import linna.util as U
        import linna.nn as N
        import linna.predictor_gpu as PG
        for row in range(g["f64_hess"].shape[0]):
            Href = g["f64_hess"][row]
            try:
                H = e.hessian(g["u"][row]) if hasattr(e, "hessian") else None
            except Exception as ex:
                H = None
                print("   hessian failed:", ex)
            if H is not None:
                print("   hessian row %d: rel err (max-norm) %.3e ; reference-f32 vs f64 %.3e" % (
                    row, np.max(np.abs(H - Href)) / np.max(np.abs(Href)),
                    np.max(np.abs(g["f32_hess"][row] - Href)) / np.max(np.abs(Href))))
    e.close()

# ---- 10^5 C3 walkers vs float64 oracle (sample of 4096 rows)
p = synthetic.make_problem(30, 500, seed=0)
e = engine.engine_from_problem(p, with_likelihood=False)
m0 = e.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
p.set_data_from_prediction(m0)
e.set_likelihood(p.priors, np.asarray(p.data, np.float32), p.inv_cov, 1.0)
n = 100000
u = synthetic.walkers(n, 30, scale=0.3, seed=1)
e.set_path("tc")
a = e.lnp(dev(u)).cpu().numpy().astype(np.float64)
e.set_path("ffma")
f = e.lnp(dev(u)).cpu().numpy().astype(np.float64)
idx = np.random.default_rng(0).choice(n, 4096, replace=False)
t0 = time.time()
ref = Oracle(p, arch).lnp(u[idx], np.float64)["lnp"]
print("oracle f64 on 4096 rows: %.1f s" % (time.time() - t0))
tol = lnp_tol(ref)
rt, rf = np.abs(a[idx] - ref) / tol, np.abs(f[idx] - ref) / tol
bins = [0, 0.1, 0.25, 0.5, 0.75, 1.0, 1.5, 2.0, 1e9]
hist = {"bins": bins, "tc": np.histogram(rt, bins)[0].tolist(), "ffma": np.histogram(rf, bins)[0].tolist(),
        "tc_max": float(rt.max()), "ffma_max": float(rf.max()), "abs_lnl_median": float(np.median(np.abs(ref))),
        "tc_abs_err_max": float(np.abs(a[idx] - ref).max()), "ffma_abs_err_max": float(np.abs(f[idx] - ref).max()),
        "tc_vs_ffma_abs_max_all_rows": float(np.abs(a - f).max()),
        "what": "C3, 4096 of 10^5 walkers (u ~ N(0, 0.3^2)), |lnP_kernel - lnP_float64_oracle| / max(1e-4, 4 ulp32)"}
print(json.dumps(hist))
out["c3_hist"] = hist
e.close()

# ---- training
for name in ["train_small", "train_ypos", "train_c3"]:
    g = load_golden(name)
    p = synthetic.make_problem(int(g["n_in"]), int(g["n_out"]), kind=str(g["kind"]), ypositive=bool(g["ypositive"]), seed=4)
    p.data = g["data"].astype(np.float64)
    if g["cov"].size:
        p.cov = g["cov"]
    dn, icov = normalised_loss_constants(p.cov, np.asarray(p.sigma, np.float32), p.y_mean, p.y_std, p.data, ypositive=p.ypositive)
    e = engine.engine_from_problem(p, with_likelihood=False)
    B = int(g["batch"])
    e.train_setup(dn, icov, B)
    shapes = arch.state_dict_shapes(p.kind, p.n_in, p.n_out)
    w = torch.from_numpy(flatten_state_dict(p.state_dict, shapes).astype(np.float32)).cuda()
    X, Y = dev(g["theta"][:B]), dev(g["target"][:B])
    cmd = torch.clamp(e.train_chisq(X, Y, 1), min=0.5 * p.n_out)
    mnn = e.train_chisq(X, Y, 0)
    nnd = e.train_chisq(X, Y, 2)

    def rel(a, b):
        return float(np.max(np.abs(np.asarray(a, np.float64) - b) / np.maximum(np.abs(b), 1e-300)))
    print(name, "chisqMd  vs f32 %.2e vs f64 %.2e (ref32 vs f64 %.2e)" % (rel(cmd.cpu().numpy(), g["chisqMd"]), rel(cmd.cpu().numpy(), g["f64_chisqMd"]), rel(g["chisqMd"], g["f64_chisqMd"])))
    print(name, "chisqnnd vs f32 %.2e vs f64 %.2e (ref32 vs f64 %.2e)" % (rel(nnd.cpu().numpy(), g["chisqnnd"]), rel(nnd.cpu().numpy(), g["f64_chisqnnd"]), rel(g["chisqnnd"], g["f64_chisqnnd"])))
    lr_ = (mnn / cmd).cpu().numpy()
    print(name, "loss_rows vs f32 %.2e vs f64 %.2e (ref32 vs f64 %.2e)" % (rel(lr_, g["loss_rows"]), rel(lr_, g["f64_loss_rows"]), rel(g["loss_rows"], g["f64_loss_rows"])))
    grads = torch.zeros_like(w)
    loss, rows = e.train_step(X, Y, cmd, None, None, None, grads, 1, float(g["lr"]), fuse_adam=False)
    print(name, "loss %.9e f32 %.9e f64 %.9e" % (float(loss), g["losses"][0], g["f64_losses"][0]))
    gd = unflatten(grads.cpu().numpy(), shapes)
    keys = [str(k) for k in g["keys"]]
    if "grad0_" + keys[0] in g:
        e32 = max(np.max(np.abs(gd[k] - g["grad0_" + k])) / np.max(np.abs(g["f64_grad0_" + k])) for k in keys)
        e64 = max(np.max(np.abs(gd[k] - g["f64_grad0_" + k])) / np.max(np.abs(g["f64_grad0_" + k])) for k in keys)
        print(name, "grad max-norm rel err: vs f32 %.2e vs f64 %.2e ; reference-f32 vs f64 %.2e" % (e32, e64, g["f32_grad0_err"].max()))
    else:
        for k, kk in (("layer1.weight", "grad0_layer1"),):
            print(name, k, "vs f32 %.2e vs f64 %.2e ; ref32 vs f64 %.2e" % (
                np.max(np.abs(gd[k] - g[kk])) / np.max(np.abs(g["f64_" + kk])), np.max(np.abs(gd[k] - g["f64_" + kk])) / np.max(np.abs(g["f64_" + kk])),
                np.max(np.abs(g[kk] - g["f64_" + kk])) / np.max(np.abs(g["f64_" + kk]))))
        norms = np.array([np.linalg.norm(gd[k].astype(np.float64)) for k in keys])
        print(name, "grad norms rel vs f64 %.2e (ref32 %.2e)" % (np.max(np.abs(norms / g["f64_grad0_norm"] - 1)), np.max(np.abs(g["grad0_norm"] / g["f64_grad0_norm"] - 1))))
    # val metric
    lossr = mnn / cmd
    frac = torch.abs(nnd / cmd - 1)
    vm = torch.stack([torch.median(lossr), torch.max(frac), torch.median(frac)]).cpu().numpy()
    print(name, "val_metric", vm, "golden", g["val_metric"])
    e.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r2_parity_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
