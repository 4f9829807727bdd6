#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE ONLY -- times the UNMODIFIED reference (the copy under oracle/_ref, see
oracle/make_ref.py) on this host and prints one JSON object.  Called as a subprocess by bench.py (the reference arm
and the `cpu_baseline` leg); never imported by the product.  It must be its own process: the reference package is
called `linna`, the same name as this repo's pickle shim.

    python oracle/ref_bench.py --task lnp|grad --workload c3|c4|c1 --procs P --seconds S [--mkldnn]
        R1 of SURVEY 8d: the reference's own Log_prob.__call__ (linna/util.py:990-1021), ONE CALL PER WALKER as
        emcee / zeus make them (linna/sampler.py:495, :728: no `vectorize`), `torch.set_num_threads(1)` per
        process, P worker processes each looping over its own walkers -- the reference's parallel model (one MPI
        rank per walker task, linna/util.py:159-231; mpi4py itself is not installed).  --task grad is
        Log_prob(nograd=False) + torch.autograd.grad as linna/HMCSampler.py:29-32.
    python oracle/ref_bench.py --task train --device cuda|cpu --steps K --warmup W [--batch 500]
        R3 of SURVEY 8d: the reference's training inner loop (linna/predictor_gpu.py:273-288) in stock eager PyTorch:
        DataLoader(pin_memory, shuffle, drop_last) -> .to(device) -> zero_grad -> model(X_transform(X)) -> Loss_fn ->
        backward -> AdamW.step -> loss.item(), with the reference's own ChtoModelv2 / Loss_fn / X_transform_class.
"""
import argparse
import importlib.util
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path = [p for p in sys.path if os.path.realpath(p or ".") != os.path.realpath(REPO)]
os.environ["LINNA_REFERENCE_ROOT"] = os.path.join(HERE, "_ref")

WORKLOADS = {"c3": (30, 500, 0), "c4": (50, 1500, 0), "c1": (33, 33, 0)}


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def _imports():
    import torch
    refstubs = _load("refstubs", os.path.join(HERE, "refstubs.py"))
    refstubs.import_reference()
    import linna.util as U
    import linna.predictor_gpu as PG
    import linna.nn as RNN
    pkg = type(sys)("linna_b200")
    pkg.__path__ = [os.path.join(REPO, "linna_b200")]
    sys.modules["linna_b200"] = pkg
    _load("linna_b200.arch", os.path.join(REPO, "linna_b200", "arch.py"))
    synthetic = _load("linna_b200.synthetic", os.path.join(REPO, "linna_b200", "synthetic.py"))
    return torch, U, PG, RNN, synthetic


def reference_objects(torch, U, PG, RNN, p, device="cpu"):
    """The reference's Predictor / transforms for a synthetic Problem (as tests/golden/make_golden.py builds them)."""
    model = getattr(RNN, p.kind)(p.n_in, p.n_out, None)
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in p.state_dict.items()})
    Xt = U.X_transform_class(torch.tensor(p.X_mean), torch.tensor(p.X_std), device, p.dolog10index)
    yt = U.Y_transform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), device, ypositive=p.ypositive)
    pred = PG.Predictor(p.n_in, p.n_out, model=model, X_transform=Xt, y_transform=yt, device=device, outdir=None)
    yinv = U.Y_invtransform_data(np.asarray(p.sigma), device)
    return pred, yinv, U.Transform(p.priors)


def _worker(args):
    """One process of the walker farm: per-walker Log_prob calls for `seconds`; returns (calls, seconds)."""
    lp, u, grad, seconds, warm, calls = args
    import torch
    torch.set_num_threads(1)

    def one(row):
        if grad:
            x = torch.tensor(row, dtype=torch.float32).clone().requires_grad_()
            val = lp(x, returntorch=True, inputnumpy=False)
            torch.autograd.grad(val, x)
        else:
            lp(row)          # numpy row in, 0-dim tensor out: what emcee's map_fn does per walker
    for i in range(warm):
        one(u[i % len(u)])
    n, t0 = 0, time.perf_counter()
    while True:
        one(u[n % len(u)])
        n += 1
        if calls:
            if n >= calls:
                break
        elif (n & 15) == 0 and time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def run_lnp(a):
    torch, U, PG, RNN, synthetic = _imports()
    torch.set_num_threads(1)
    n_in, n_out, _ = WORKLOADS[a.workload]
    p = synthetic.make_problem(n_in, n_out, seed=0)
    pred, yinv, transform = reference_objects(torch, U, PG, RNN, p)
    with torch.no_grad():
        m0 = yinv(pred.predict(torch.tensor(np.asarray(p.theta0, np.float32)))).detach().numpy()
    p.set_data_from_prediction(m0)
    if a.mkldnn:                                                   # the setting linna/main.py:266-268 picks on CPU
        pred.model = pred.model.to(memory_format=torch.channels_last)
        pred.MKLDNN = True
    data = torch.from_numpy(p.data.astype(np.float32)).clone().requires_grad_()
    invcov = torch.from_numpy(p.inv_cov.astype(np.float32)).clone().requires_grad_()
    grad = a.task == "grad"
    lp = U.Log_prob(data, invcov, pred, yinv, transform, p.temperature, U.gaussianlogliklihood, nograd=not grad)
    u = synthetic.walkers(4096, n_in, scale=0.3, seed=1)
    procs = max(1, a.procs)
    per = (a.calls + procs - 1) // procs if a.calls else 0
    jobs = [(lp, u[i::procs].astype(np.float64 if not grad else np.float32), grad, a.seconds, 20, per) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_worker, jobs)
    wall = time.perf_counter() - t0
    # --seconds: every process runs its own timed loop; --calls: a fixed number of walkers, rate = walkers / slowest process
    rate = sum(n for n, _ in res) / max(dt for _, dt in res) if a.calls else sum(n / dt for n, dt in res)
    one = _worker((lp, u.astype(np.float64 if not grad else np.float32), grad, min(a.seconds, 2.0), 20, 0)) if procs > 1 else res[0]
    return {"task": a.task, "workload": a.workload, "procs": procs, "calls": int(sum(n for n, _ in res)),
            "evals_per_s": rate, "evals_per_s_one_process": one[0] / one[1], "seconds": a.seconds, "wall_s": wall,
            "mkldnn": bool(a.mkldnn), "torch": torch.__version__,
            "what": "reference Log_prob.__call__ per walker (linna/util.py:990-1021)%s, torch.set_num_threads(1) per process, "
                    "%d processes" % (" + torch.autograd.grad (linna/HMCSampler.py:29-32)" if grad else "", procs)}


def run_train(a):
    torch, U, PG, RNN, synthetic = _imports()
    from torch.utils.data import DataLoader
    dev = a.device
    if dev == "cuda" and not torch.cuda.is_available():
        return {"task": "train", "unavailable": "no CUDA device"}
    B = a.batch
    p = synthetic.make_problem(30, 500, seed=4)
    rng = np.random.default_rng(9)
    theta = synthetic.training_set(p, 10000, seed=3, spread=0.3)
    pred, yinv, _ = reference_objects(torch, U, PG, RNN, p, device="cpu")
    with torch.no_grad():
        m = yinv(pred.predict(torch.tensor(theta.astype(np.float32)))).detach().numpy().astype(np.float64)
    p.data = m[0].copy()
    target = m * (1 + 0.01 * rng.standard_normal(m.shape))
    sig = np.asarray(p.sigma, np.float32)
    ytd = U.Y_transform_data(sig, dev)
    yinvc = U.Y_invtransform_class(torch.tensor(p.y_mean).to(dev), torch.tensor(p.y_std).to(dev),
                                   torch.tensor(p.data.astype(np.float32)).to(dev), dev)
    loss_fn = U.Loss_fn(torch.tensor(p.data.astype(np.float32)).to(dev), torch.tensor(p.cov).to(dev),
                        torch.tensor(p.inv_cov).to(dev), ytd, yinvc, dev)
    Xt = U.X_transform_class(torch.tensor(p.X_mean).to(dev), torch.tensor(p.X_std).to(dev), dev)
    torch.manual_seed(1234)                                                  # predictor_gpu.py:221
    model = RNN.ChtoModelv2(30, 500, None).to(dev)
    optim = torch.optim.AdamW(params=model.parameters(), lr=1e-3, weight_decay=1e-4)   # :267
    loader = DataLoader(U.ArrayDataset(theta, target), batch_size=B, shuffle=True, drop_last=True, num_workers=0,
                        pin_memory=(dev == "cuda"))                          # util.py:1285
    if dev == "cpu":
        torch.set_num_threads(a.threads or os.cpu_count() or 1)
    losses = []
    done, t0 = 0, None
    total = a.steps + a.warmup
    model.train()
    while done < total:
        for X, y_target in loader:                                          # predictor_gpu.py:273-288
            if done == a.warmup:
                if dev == "cuda":
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
            X = X.to(dev)
            y_target = y_target.to(dev)
            optim.zero_grad()
            y_pred = model(Xt(X))
            loss = loss_fn(y_pred, y_target)
            loss.backward()
            optim.step()
            losses.append(loss.item())
            done += 1
            if done >= total:
                break
    if dev == "cuda":
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"task": "train", "device": dev, "batch": B, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps,
            "rows_per_s": B * a.steps / dt, "loss_first": losses[0], "loss_last": losses[-1], "torch": torch.__version__,
            "threads": torch.get_num_threads(),
            "what": "reference training inner loop (linna/predictor_gpu.py:273-288) in stock eager PyTorch on %s: reference "
                    "ChtoModelv2 + Loss_fn + X_transform_class, torch.optim.AdamW(wd=1e-4), DataLoader(pin_memory, shuffle, "
                    "drop_last), loss.item() every step" % dev}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="lnp", choices=["lnp", "grad", "train"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--seconds", type=float, default=5.0)
    ap.add_argument("--calls", type=int, default=0, help="total per-walker calls over all processes (instead of --seconds)")
    ap.add_argument("--mkldnn", action="store_true")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    if not os.path.isfile(os.path.join(HERE, "_ref", "linna", "util.py")):
        print(json.dumps({"unavailable": "oracle/_ref/linna missing (run oracle/make_ref.py where /root/reference exists)"}))
        return 0
    out = run_train(a) if a.task == "train" else run_lnp(a)
    print("REF_BENCH " + json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
