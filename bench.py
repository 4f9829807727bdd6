#!/usr/bin/env python
"""Headline benchmark: emulator log-likelihood evaluations per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c3|c4|c1|c5] [--mode lnp|grad] [--no-extra]
    python bench.py --impl reference ...   # the UNMODIFIED reference (oracle/_ref) on the box's host cores

A *step* is one pass of the hot path over one batch of synthetic walkers: u[n, n_in] -> lnP[n]
(--mode grad: also d lnP/du).  At N=1 the workload is BASELINE config C3 (DES-Y3-3x2pt-shaped
emulator, n_in=30, n_out=500, 1e5 walkers); with N>1 every rank evaluates its own 1e5 walkers
(independent walkers shard with no collective on the data path -> weak scaling).

One JSON line is printed by rank 0.  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` is the same metric through the reference-facing call with PAGEABLE numpy buffers in and
out (Engine.lnp -> linna_lnp_host), host<->device copies inside the timed region (`e2e.pinned` is the
same call with caller-pinned buffers).  The default run (C3 lnP) also carries, under `extra`, the rest of
the metric: C3 lnP+grad, C4 lnP+grad (weak, and BASELINE's "1e4 chains over N GPUs" split), C5 training (weak
B=500 per GPU and strong B=500/N, NCCL all-reduce inside the step for N>1, next to the reference's own
training loop in eager PyTorch on the same B200), C1 latency, and a `sustained` figure (seconds-long loop, clocks
sampled, fraction against the sustained bf16 peak).  `cpu_baseline` is the reference's own per-walker
Log_prob.__call__ farmed over the host cores (oracle/ref_bench.py, kind "reference").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from linna_b200 import arch, synthetic  # noqa: E402

WORKLOADS = {
    # name: (n_in, n_out, walkers per GPU, description)
    "c3": (30, 500, 100000, "C3 DES-Y3-3x2pt-shaped ChtoModelv2 30->500, 1e5 walkers/GPU, flat priors, T=1"),
    "c4": (50, 1500, 10000, "C4 LSST-Y10-6x2pt+N-shaped ChtoModelv2 50->1500, 1e4 chains/GPU"),
    "c1": (33, 33, 4, "C1 README 33-dim Gaussian, 4 walkers"),
    "c5": (30, 500, 500, "C5 emulator training at the C3 shape, batch 500"),
}
METRIC = "emulator log-likelihood evals/sec"
METRIC_GRAD = "emulator log-likelihood+grad evals/sec"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        # one ~1 ms kernel launched back to back for a few tens of ms: the burst figure is the honest denominator;
        # the seconds-long `sustained` loop is held against the sustained figure
        return dict(bf16_tflops=d.get("bf16_tflops", d.get("bf16_tflops_sustained")),
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d.get("bf16_tflops")), hbm_gbs=d.get("hbm_gbs"),
                    source="MEASURED_PEAKS.json (bf16 burst, of measured)")
    return dict(bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, hbm_gbs=6650.0,
                source="fallback (B200_PROFILING.md), of fallback")


def load_traffic(kernel, workload, mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of the same workload (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("%s:%s:%s" % (kernel, workload, mode))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=100):
        self.index, self.rows, self.proc, self.period = index, [], None, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is supposed to use every host thread it can."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=host_threads())
    except Exception:
        import contextlib
        return contextlib.nullcontext()


# ----------------------------------------------------------------------------------------------------------
# CPU / reference baselines (the only place bench.py executes anything under oracle/)
def ref_bench(*argv, timeout=900):
    """Run oracle/ref_bench.py (the UNMODIFIED reference from oracle/_ref) in its own process; returns its dict or
    {'unavailable': why}."""
    if not os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "linna", "util.py")):
        return {"unavailable": "oracle/_ref/linna missing"}
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py")] + [str(a) for a in argv],
                           cwd="/tmp", env=env, capture_output=True, text=True, timeout=timeout)
    except Exception as e:
        return {"unavailable": "ref_bench failed: %r" % (e,)}
    for line in r.stdout.splitlines():
        if line.startswith("REF_BENCH "):
            return json.loads(line[len("REF_BENCH "):])
    return {"unavailable": "ref_bench printed no result (rc=%d): %s" % (r.returncode, (r.stderr or r.stdout)[-300:])}


def cpu_port_rate(p, data, n_sample, repeats, grad=False):
    """Time the oracle's batched numpy/BLAS port of the reference arithmetic on the host cores (R2 of SURVEY 8d: the
    charitable, vectorised CPU number)."""
    from oracle.oracle import NumpyPort, Oracle
    p.data = data
    o = Oracle(p, arch)
    port = NumpyPort(o)
    call = port.lnp_grad if grad else port.lnp
    u = synthetic.walkers(n_sample, p.n_in, scale=0.3, seed=11)
    with all_host_threads():
        call(u[:256])
        t0 = time.perf_counter()
        for _ in range(repeats):
            call(u)
        dt = time.perf_counter() - t0
    return n_sample * repeats / dt, dt


def reference_cpu_baseline(workload, mode, seconds, p=None, data=None):
    """cpu_baseline object: R1 (the reference's own per-walker Log_prob.__call__, one process per host thread) when
    oracle/_ref travelled, else the numpy port."""
    cores = host_threads()
    r = ref_bench("--task", "grad" if mode == "grad" else "lnp", "--workload", workload, "--procs", cores, "--seconds", seconds)
    if "unavailable" not in r:
        out = {"value": r["evals_per_s"], "unit": "evals/s", "cores": cores, "kind": "reference",
               "sample": "%d per-walker calls of the reference's own Log_prob.__call__ (linna/util.py:990-1021%s) from oracle/_ref, "
                         "%d processes x 1 thread for %.0f s each (the reference's walker farm, linna/util.py:159-231)"
                         % (r["calls"], " + torch.autograd.grad" if mode == "grad" else "", cores, seconds),
               "per_process": r["evals_per_s_one_process"]}
    else:
        out = {"value": None, "unit": "evals/s", "cores": cores, "kind": "port", "sample": "reference unavailable: " + r["unavailable"]}
    if p is not None:
        n_sample = 16384
        rate, dt_cpu = cpu_port_rate(p, data, n_sample, 4 if mode == "lnp" else 2, grad=mode == "grad")
        out["vectorised_port"] = {"value": rate, "cores": cores,
                                  "sample": "oracle.NumpyPort (batched numpy/BLAS port of the same arithmetic, R2), %d walkers x %d, %.1f s"
                                            % (n_sample, 4 if mode == "lnp" else 2, dt_cpu)}
        if out["value"] is None:
            out["value"] = rate
    return out


def run_reference(args):
    """--impl reference: the UNMODIFIED reference's own implementation of the path on the box's host cores.  A step is
    a bounded sample of `--ref-sample` walkers, each evaluated by its own Log_prob.__call__ (linna/util.py:990-1021)
    exactly as emcee / zeus call it, farmed over one process per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload == "c5":
        return run_reference_train(args)
    n_in, n_out, n, desc = WORKLOADS[args.workload]
    cores = host_threads()
    sample = max(cores, min(args.ref_sample, 2048))
    calls = sample * (args.steps + max(args.warmup, 0))
    r = ref_bench("--task", "grad" if args.mode == "grad" else "lnp", "--workload", args.workload, "--procs", cores,
                  "--calls", calls)
    metric = METRIC if args.mode == "lnp" else METRIC_GRAD
    if "unavailable" in r:   # the oracle always exists: fall back to its port, and say so
        p = synthetic.make_problem(n_in, n_out, seed=0)
        from oracle.oracle import Oracle
        m0 = Oracle(p, arch).lnp(np.zeros((1, p.n_in), np.float32), want=("m",))
        data = p.set_data_from_prediction(m0["m"][0])
        val, dt = cpu_port_rate(p, data, 16384, max(args.steps, 1), grad=args.mode == "grad")
        kind, smp = "port", "oracle.NumpyPort, 16384 walkers per step (%s)" % r["unavailable"]
        ms = 1e3 * dt / max(args.steps, 1)
    else:
        val = r["evals_per_s"]
        ms = 1e3 * sample / val
        kind = "reference"
        smp = ("%d walkers per step, each one call of the reference's own Log_prob.__call__ (oracle/_ref, linna/util.py:990-1021%s), "
               "%d processes x 1 thread" % (sample, " + torch.autograd.grad, linna/HMCSampler.py:29-32" if args.mode == "grad" else "", cores))
    line = {"impl": "reference", "metric": metric, "value": val, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "mode": args.mode},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": kind, "sample": smp},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_reference_train(args):
    """--impl reference --workload c5: the reference's own training inner loop (linna/predictor_gpu.py:273-288) in stock
    eager PyTorch -- on the B200 when one is visible (docuda=True, linna/util.py:1320-1321: where the reference trains),
    else on the host cores."""
    import torch
    B = args.walkers or 500
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    steps = max(1, args.steps if dev == "cuda" else min(args.steps, 10))
    r = ref_bench("--task", "train", "--device", dev, "--steps", steps, "--warmup", max(args.warmup, 1), "--batch", B)
    if "unavailable" in r:
        print(json.dumps({"impl": "reference", "unavailable": r["unavailable"]}), flush=True)
        return 0
    val = r["rows_per_s"]
    line = {"impl": "reference", "metric": "emulator training rows/sec", "value": val, "unit": "rows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS["c5"][3], "mode": "train", "device": dev},
            "cpu_baseline": {"value": val, "unit": "rows/s", "cores": host_threads(), "kind": "reference", "sample": r["what"]},
            "e2e": {"value": val, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------
class Dist:
    """One process per GPU; world 1 without torchrun."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- linna_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([float(v)], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_engine(D, workload, path="auto"):
    from linna_b200 import engine
    n_in, n_out, n, desc = WORKLOADS[workload]
    p = synthetic.make_problem(n_in, n_out, seed=0)
    eng = engine.engine_from_problem(p, device=D.local, with_likelihood=False)
    m0 = eng.predict(np.asarray(p.theta0, np.float32)[None, :], engine.LINNA_OUT_M)[0]
    data = p.set_data_from_prediction(m0)
    eng.set_likelihood(p.priors, np.asarray(data, np.float32), p.inv_cov, p.temperature)
    eng.set_path(path)
    return p, eng, data


def measure_lnp(D, p, eng, n, mode, steps, warmup, want_e2e=True, sample_clocks=True):
    """Device-resident and host-buffer throughput of one lnP / lnP+grad pass over n walkers per rank."""
    torch = D.torch
    from linna_b200 import engine
    nbuf = 16 if n * p.n_in * 4 * 16 <= (1 << 31) else 4
    nbuf = min(nbuf, max(2, (1 << 28) // max(n * p.n_in * 4, 1)))
    u_host = [synthetic.walkers(n, p.n_in, scale=0.3, seed=100 + 17 * D.rank + b) for b in range(nbuf)]
    u_dev = [torch.from_numpy(u).cuda() for u in u_host]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    call = eng.lnp_grad if mode == "grad" else eng.lnp
    for w in range(warmup):
        call(u_dev[w % nbuf])
    flush.zero_()
    D.barrier()
    sampler = ClockSampler(D.local).start() if (D.rank == 0 and sample_clocks) else None
    l0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(steps):
        call(u_dev[s % nbuf])
    ev1.record()
    D.barrier()
    launches = engine.launch_count() - l0
    ms = D.max(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    out = {"ms_per_step": ms / steps, "launches": int(launches), "clocks": clocks, "nbuf": nbuf, "kernel_path": eng.last_kernel()}
    if not want_e2e:
        return out
    h2d = n * p.n_in * 4
    d2h = n * 4 * (1 + (p.n_in if mode == "grad" else 0))

    def timed(fn):
        # host-side timing of a few-ms region is at the mercy of one scheduling hiccup: at least 50 calls, reported as
        # the time of `steps` calls
        reps = max(steps, 50)
        for w in range(3):
            fn(w)
        D.barrier()
        t0 = time.perf_counter()
        for s in range(reps):
            fn(s)
        torch.cuda.synchronize()
        return D.max(time.perf_counter() - t0) * steps / reps

    # (a) the reference-facing case: an emcee/zeus caller hands over PAGEABLE numpy arrays and receives fresh ones
    page = u_host[:4]

    def call_pageable(s):
        return eng.lnp_grad(page[s % len(page)]) if mode == "grad" else eng.lnp(page[s % len(page)])
    dt_page = timed(call_pageable)
    # (b) best case: caller-pinned input and output buffers (two result sets, alternating)
    pin = [torch.from_numpy(u).pin_memory().numpy() for u in u_host[:4]]
    res_l = [torch.empty(n, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
    res_g = [torch.empty(n, p.n_in, dtype=torch.float32).pin_memory().numpy() for _ in range(2)] if mode == "grad" else None

    def call_pinned(s):
        if mode == "grad":
            return eng.lnp_grad(pin[s % 4], out=res_l[s % 2], out_grad=res_g[s % 2])
        return eng.lnp(pin[s % 4], out=res_l[s % 2])
    dt_pin = timed(call_pinned)
    out["e2e_s"], out["e2e_pinned_s"], out["h2d"], out["d2h"] = dt_page, dt_pin, h2d, d2h
    return out


def measure_sustained(D, p, eng, n, mode, seconds=2.5):
    """The same launch back to back for `seconds` (the regime a sampler lives in): CUDA-event time, clocks sampled."""
    torch = D.torch
    u_dev = [torch.from_numpy(synthetic.walkers(n, p.n_in, scale=0.3, seed=300 + D.rank + b)).cuda() for b in range(4)]
    call = eng.lnp_grad if mode == "grad" else eng.lnp
    for w in range(5):
        call(u_dev[w % 4])
    D.barrier()
    # size the loop from a short probe so that every rank issues the same number of launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(20):
        call(u_dev[s % 4])
    ev1.record()
    torch.cuda.synchronize()
    per = D.max(ev0.elapsed_time(ev1) / 20.0)
    iters = max(50, int(seconds * 1e3 / per))
    D.barrier()
    sampler = ClockSampler(D.local, 200).start() if D.rank == 0 else None
    ev0.record()
    for s in range(iters):
        call(u_dev[s % 4])
    ev1.record()
    D.barrier()
    ms = D.max(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    return {"ms_per_step": ms / iters, "steps": iters, "seconds": ms * 1e-3, "clocks": clocks}


def train_problem(seed=4):
    """C5: emulator training at the C3 shape, B=500 (yamlfile/training_3x2pt_gpu.yaml:36-40)."""
    p = synthetic.make_problem(30, 500, seed=seed)
    rng = np.random.default_rng(9)
    theta = synthetic.training_set(p, 10000, seed=3, spread=0.3)
    return p, theta, rng


_TRAIN_CACHE = {}


def train_setup(D):
    """Synthetic training set resident on the GPU + the loss constants (shared by the weak and strong legs)."""
    if "set" in _TRAIN_CACHE:
        return _TRAIN_CACHE["set"]
    import torch
    from linna_b200 import engine
    import linna.util as U
    p, theta, rng = train_problem()
    eng = engine.engine_from_problem(p, device=D.local, with_likelihood=False)
    th32 = np.ascontiguousarray(theta, np.float32)
    m = eng.predict(th32, engine.LINNA_OUT_M).astype(np.float64)
    p.data = m[0].copy()
    target = m * (1 + 0.01 * rng.standard_normal(m.shape))          # model(theta) + 1 % noise
    eng.close()
    sig = np.asarray(p.sigma, np.float32)
    ytd = U.Y_transform_data(sig, "cpu")
    yinv = U.Y_invtransform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), torch.tensor(p.data.astype(np.float32)), "cpu")
    loss_fn = U.Loss_fn(torch.tensor(p.data.astype(np.float32)), torch.tensor(p.cov), torch.tensor(p.inv_cov), ytd, yinv, "cpu")
    xt = U.X_transform_class(torch.tensor(p.X_mean), torch.tensor(p.X_std), "cpu")
    yt = U.Y_transform_class(torch.tensor(p.y_mean), torch.tensor(p.y_std), "cpu")
    X = torch.from_numpy(th32).cuda()
    Y = torch.from_numpy(target.astype(np.float32)).cuda()
    _TRAIN_CACHE["set"] = (p, xt, yt, loss_fn, X, Y)
    return _TRAIN_CACHE["set"]


def measure_train(D, B_rank, steps, warmup, want_e2e=True):
    """One step = one AdamW optimiser step on B_rank rows per rank (forward, loss, backward, weight gradients, NCCL
    all-reduce of the flat gradient when world > 1, update)."""
    import torch
    from linna_b200 import engine
    from linna_b200.train import FusedTrainer
    import linna.nn as N
    p, xt, yt, loss_fn, X, Y = train_setup(D)
    torch.manual_seed(1234)
    model = N.ChtoModelv2(30, 500, None)
    tr = FusedTrainer(model, xt, yt, loss_fn.auxileryfunction, B_rank, device_index=D.local, lr=1e-3, world_size=D.world)
    cmd = tr.chisq_md(X, Y)
    n = X.shape[0]
    B = B_rank
    nb = n // B
    gen = torch.Generator().manual_seed(D.rank)
    perm = torch.randperm(n, generator=gen).cuda()
    nb = min(nb, 40)
    batches = [perm[b * B:(b + 1) * B] for b in range(nb)]
    xb = [X[i].contiguous() for i in batches]
    yb = [Y[i].contiguous() for i in batches]
    cb = [cmd[i].contiguous() for i in batches]
    losses = torch.zeros(steps + warmup, device="cuda")
    for w in range(warmup):
        tr.step(xb[w % nb], yb[w % nb], cb[w % nb], loss_out=losses[w:w + 1])
    D.barrier()
    sampler = ClockSampler(D.local).start() if D.rank == 0 else None
    l0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for s in range(steps):
        j = (s + warmup) % nb
        tr.step(xb[j], yb[j], cb[j], loss_out=losses[warmup + s:warmup + s + 1])
    ev1.record()
    D.barrier()
    launches = engine.launch_count() - l0
    ms = D.max(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    out = {"ms_per_step": ms / steps, "launches": int(launches), "clocks": clocks, "kernel_path": tr.kernel_path(),
           "loss_first": float(losses[0].item()), "loss_last": float(losses[-1].item())}
    if want_e2e:
        # batch arrives in pinned host memory, loss is read back every step (the reference's loss.item())
        hx = [t_.cpu().pin_memory() for t_ in xb[:8]]
        hy = [t_.cpu().pin_memory() for t_ in yb[:8]]
        hc = [t_.cpu().pin_memory() for t_ in cb[:8]]
        def e2e_step(s):
            j = s % len(hx)
            l_ = tr.step(hx[j].cuda(non_blocking=True), hy[j].cuda(non_blocking=True), hc[j].cuda(non_blocking=True))
            return float(l_.item())
        reps = max(steps, 100)                     # (a 20-step region is 8 ms of host time: one hiccup moves it by 30 %)
        for s in range(3):
            e2e_step(s)
        D.barrier()
        t0 = time.perf_counter()
        for s in range(reps):
            e2e_step(s)
        torch.cuda.synchronize()
        out["e2e_s"] = D.max(time.perf_counter() - t0) * steps / reps
    tr.engine.close()
    return out


def train_flops_row():
    return 6 * arch.macs_forward("ChtoModelv2", 30, 500) + 500 * 501


def train_line(D, args, main=True):
    """The C5 record (weak: B=500 per GPU; plus, for N > 1, strong: B=500 over all GPUs)."""
    B = args.walkers or 500
    steps, warmup = max(args.steps, 20), max(args.warmup, 3)
    r = measure_train(D, B, steps, warmup)
    world = D.world
    value = B * world * steps / (r["ms_per_step"] * steps * 1e-3)
    peaks = load_peaks()
    achieved = train_flops_row() * B / (r["ms_per_step"] * 1e-3) / 1e12
    rec = {"name": "c5_train_weak", "metric": "emulator training rows/sec", "value": value, "unit": "rows/s", "n_gpus": world,
           "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "scaling": "weak",
           "config": {"workload": "C5 emulator training, ChtoModelv2 30->500, batch %d per GPU, AdamW wd=1e-4, 10^4-row synthetic set%s"
                                  % (B, ", NCCL all-reduce of the flat gradient inside every step" if world > 1 else ""),
                      "steps_per_sec": 1e3 / r["ms_per_step"], "flops_per_row": train_flops_row(), "kernel_path": r["kernel_path"],
                      "loss_first": r["loss_first"], "loss_last": r["loss_last"],
                      "l2": "batches rotate over the 10^4-row set; working set (weights+moments 21 MB) is L2 resident"},
           "clocks": r["clocks"], "gpu_launches": r["launches"],
           "e2e": {"value": B * world * steps / r["e2e_s"], "unit": "rows/s", "h2d_bytes_per_step": B * (30 + 500 + 1) * 4,
                   "d2h_bytes_per_step": 4},
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["bf16_tflops"], "traffic": None, "peak_source": peaks["source"],
                        "note": "useful flops = B x (6 MACs_fwd + n_out(n_out+1)) = %.2f GFLOP per step (SURVEY 8d)" % (train_flops_row() * B / 1e9)}}
    if world > 1 and B % world == 0:
        rs = measure_train(D, B // world, steps, warmup, want_e2e=False)
        rec["strong"] = {"name": "c5_train_strong", "value": B * steps / (rs["ms_per_step"] * steps * 1e-3), "unit": "rows/s",
                         "ms_per_step": rs["ms_per_step"], "scaling": "strong", "rows_per_gpu": B // world,
                         "gpu_launches": rs["launches"]}
    if D.rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the real competitor: the reference's own training loop in stock eager PyTorch on this B200
        D.torch.cuda.synchronize()
        eager = ref_bench("--task", "train", "--device", "cuda", "--steps", 60, "--warmup", 10, "--batch", B)
        cpu = ref_bench("--task", "train", "--device", "cpu", "--steps", 6, "--warmup", 2, "--batch", B)
        if "unavailable" not in eager:
            rec["eager_torch_b200"] = {"value": eager["rows_per_s"], "unit": "rows/s", "ms_per_step": eager["ms_per_step"],
                                       "kind": "reference", "sample": eager["what"]}
            rec["vs_eager_torch_b200"] = rec["e2e"]["value"] / eager["rows_per_s"]
            rec["vs_eager_torch_b200_note"] = "e2e (pinned host batch in, loss.item() every step) / eager rows/s, same protocol"
        else:
            rec["eager_torch_b200"] = eager
        if "unavailable" not in cpu:
            rec["cpu_baseline"] = {"value": cpu["rows_per_s"], "unit": "rows/s", "cores": cpu["threads"], "kind": "reference",
                                   "sample": cpu["what"]}
        else:
            rec["cpu_baseline"] = {"value": None, "kind": "reference", "sample": cpu["unavailable"]}
    return rec


def lnp_record(D, args, workload, mode, n, steps, warmup, name, p=None, eng=None, scaling="weak", want_e2e=True):
    close = eng is None
    if eng is None:
        p, eng, _ = make_engine(D, workload, args.path)
    r = measure_lnp(D, p, eng, n, mode, steps, warmup, want_e2e=want_e2e)
    total = n * D.world * steps
    value = total / (r["ms_per_step"] * steps * 1e-3)
    peaks = load_peaks()
    fe = arch.flops_lnl_grad(p.kind, p.n_in, p.n_out) if mode == "grad" else arch.flops_lnl(p.kind, p.n_in, p.n_out)
    achieved = fe * n / (r["ms_per_step"] * 1e-3) / 1e12
    rec = {"name": name, "metric": METRIC if mode == "lnp" else METRIC_GRAD, "value": value, "unit": "evals/s", "n_gpus": D.world,
           "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "scaling": scaling,
           "config": {"workload": WORKLOADS[workload][3], "mode": mode, "walkers_per_gpu": n, "flops_per_eval": fe,
                      "kernel_path": r["kernel_path"]},
           "clocks": r["clocks"], "gpu_launches": r["launches"],
           "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["bf16_tflops"],
                        "traffic": load_traffic("linna::tc_f16_kernel", workload, mode) if r["kernel_path"] == "tc" else None}}
    if want_e2e:
        rec["e2e"] = {"value": total / r["e2e_s"], "unit": "evals/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                      "buffers": "pageable numpy in / fresh numpy out", "pinned": total / r["e2e_pinned_s"]}
    if close:
        eng.close()
    return rec


def c1_record(D, args):
    """C1 (README 33-dim Gaussian, 4 walkers): latency regime -- microseconds per ensemble call."""
    torch = D.torch
    p, eng, data = make_engine(D, "c1", "auto")
    n = 4
    u = synthetic.walkers(n, p.n_in, scale=0.3, seed=5)
    ud = torch.from_numpy(u).cuda()
    for _ in range(20):
        eng.lnp(ud)
    D.barrier()
    iters = 300
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(iters):
        eng.lnp(ud)
    ev1.record()
    torch.cuda.synchronize()
    us_dev = D.max(ev0.elapsed_time(ev1) / iters * 1e3)
    for _ in range(10):
        eng.lnp(u)
    t0 = time.perf_counter()
    for _ in range(iters):
        eng.lnp(u)
    us_host = D.max((time.perf_counter() - t0) / iters * 1e6)
    rec = {"name": "c1_latency", "metric": "microseconds per 4-walker ensemble call", "value": us_dev, "unit": "us", "higher_is_better": False,
           "n_gpus": D.world, "steps": iters, "config": {"workload": WORKLOADS["c1"][3], "kernel_path": eng.last_kernel()},
           "evals_per_s": n * 1e6 / us_dev,
           "e2e": {"value": us_host, "unit": "us", "note": "numpy in / numpy out through linna_lnp_host, one call = 4 walkers",
                   "evals_per_s": n * 1e6 / us_host}}
    if D.rank == 0 and D.world == 1 and not args.no_cpu_baseline:
        r = ref_bench("--task", "lnp", "--workload", "c1", "--procs", 1, "--seconds", 2)
        if "unavailable" not in r:
            rec["cpu_baseline"] = {"value": r["evals_per_s"], "unit": "evals/s", "cores": 1, "kind": "reference",
                                   "sample": "reference Log_prob.__call__ per walker, one process (4 walkers do not fill a farm), 2 s"}
    eng.close()
    return rec


def sampler_record(D, args, p, eng, iters=10, warmup=3):
    """f1: the ensemble sampler's own loop on the device -- per iteration one permutation, and per half-ensemble one
    stretch-move proposal kernel, one likelihood launch over 10^5 proposals and one accept kernel
    (linna_b200/sampler.py: EnsembleSampler.sample; the reference drives emcee / zeus with one Log_prob call per walker,
    linna/sampler.py:495)."""
    torch = D.torch
    from linna_b200 import engine
    from linna_b200.sampler import EnsembleSampler
    W = 2 * WORKLOADS["c3"][2]
    smp = EnsembleSampler(W, p.n_in, eng.lnp, seed=11 + D.rank, device="cuda:%d" % D.local)
    x = torch.from_numpy(synthetic.walkers(W, p.n_in, scale=0.3, seed=900 + D.rank)).cuda()
    x, lnp = smp.run_mcmc(x, warmup, store=False)
    D.barrier()
    l0 = engine.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for x, lnp in smp.sample(x, iters, store=False, lnp0=lnp):
        pass
    ev1.record()
    D.barrier()
    ms = D.max(ev0.elapsed_time(ev1))
    acc = float(smp.acceptance_fraction.mean())
    return {"name": "c3_ensemble_sampler", "metric": "ensemble-sampler walker updates/sec", "value": W * D.world * iters / (ms * 1e-3),
            "unit": "walker updates/s", "n_gpus": D.world, "steps": iters, "warmup": warmup, "ms_per_step": ms / iters, "scaling": "weak",
            "config": {"workload": "C3, Goodman-Weare stretch move, %d walkers per GPU (two half-ensemble updates of 1e5 proposals per "
                                   "iteration), every array device-resident" % W,
                       "kernel_path": eng.last_kernel()},
            "gpu_launches": int(engine.launch_count() - l0), "acceptance_fraction": acc,
            "lnp_share": "one iteration = 2 likelihood launches of 1e5 walkers + 2 proposal + 2 accept kernels + torch.randperm"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="lnp", choices=["lnp", "grad"])
    ap.add_argument("--walkers", type=int, default=0, help="override walkers per GPU")
    ap.add_argument("--ref-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="only the headline record (no C3 grad / C4 / C5 / C1 / sustained)")
    ap.add_argument("--path", default="auto", choices=["auto", "ffma", "tc"], help="kernel serving lnP")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    D = Dist()
    world, rank = D.world, D.rank
    if args.workload == "c5":
        rec = train_line(D, args)
        if rank == 0:
            line = {"metric": rec["metric"], "value": rec["value"], "unit": rec["unit"], "n_gpus": world, "steps": rec["steps"],
                    "warmup": rec["warmup"], "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
            line.update({k: v for k, v in rec.items() if k not in line and k != "name"})
            line.setdefault("cpu_baseline", None)
            print(json.dumps(line), flush=True)
        D.close()
        return 0

    n_in, n_out, n, desc = WORKLOADS[args.workload]
    if args.walkers:
        n = args.walkers
    p, eng, data = make_engine(D, args.workload, args.path)
    r = measure_lnp(D, p, eng, n, args.mode, args.steps, args.warmup)
    total_evals = n * world * args.steps
    value = total_evals / (r["ms_per_step"] * args.steps * 1e-3)
    clocks = r["clocks"]

    extra, sustained = [], None
    do_extra = not args.no_extra and args.workload == "c3" and args.mode == "lnp" and not args.walkers
    if do_extra:
        s = measure_sustained(D, p, eng, n, "lnp")
        peaks = load_peaks()
        fe = arch.flops_lnl(p.kind, p.n_in, p.n_out)
        ach = fe * n / (s["ms_per_step"] * 1e-3) / 1e12
        sustained = {"value": n * world / (s["ms_per_step"] * 1e-3), "unit": "evals/s", "ms_per_step": s["ms_per_step"], "steps": s["steps"],
                     "seconds": s["seconds"], "clocks": s["clocks"], "achieved_tflops": ach,
                     "peak": peaks["bf16_tflops_sustained"], "frac": ach / peaks["bf16_tflops_sustained"],
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"}
        extra.append(lnp_record(D, args, "c3", "grad", n, max(args.steps, 10), args.warmup, "c3_lnp_grad", p=p, eng=eng))
        sg = measure_sustained(D, p, eng, n, "grad", seconds=2.0)
        extra[-1]["sustained"] = {"value": n * world / (sg["ms_per_step"] * 1e-3), "ms_per_step": sg["ms_per_step"], "seconds": sg["seconds"],
                                  "clocks": sg["clocks"]}
        extra.append(sampler_record(D, args, p, eng))

    if rank == 0:
        # ---- roofline of the dominant (only) kernel: useful flops / measured kernel time
        peaks = load_peaks()
        flops_eval = arch.flops_lnl_grad(p.kind, p.n_in, p.n_out) if args.mode == "grad" else arch.flops_lnl(p.kind, p.n_in, p.n_out)
        per_step_s = r["ms_per_step"] * 1e-3
        achieved = flops_eval * n / per_step_s / 1e12
        sms = eng.info()["num_sms"]
        sm_mhz = (clocks or {}).get("sm_mhz") or 0.0
        if r["kernel_path"] == "tc":
            kernel = "linna::tc_f16_kernel"
            note = ("tcgen05 kind::f16 MMAs (cta_group::2, TMEM accumulators, TMA operands); every fp32 operand is split "
                    "into two fp16 halves and multiplied in 3 passes, so the tensor pipe ISSUES 3x the useful flops: "
                    "%.0f TFLOP/s issued = %.3f of the measured bf16 peak" % (3 * achieved, 3 * achieved / peaks["bf16_tflops"]))
        else:
            kernel = "linna::fused_ffma_kernel<4>"
            ffma_peak = sms * 128 * 2 * sm_mhz * 1e6 / 1e12
            note = ("kernel computes in FP32 FFMA (exact-fp32 path); FP32 CUDA-core peak at the sampled clock = "
                    "%.1f TFLOP/s -> %.3f of that" % (ffma_peak, achieved / max(ffma_peak, 1e-9)))
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / peaks["bf16_tflops"], "traffic": load_traffic(kernel, args.workload, args.mode),
                    "kernel": kernel, "peak_source": peaks["source"], "note": note}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # reported at N=1 only (the reference arm covers the other N)
        cpu = reference_cpu_baseline(args.workload, args.mode, 6.0, p, data)
        if do_extra:
            extra[0]["cpu_baseline"] = reference_cpu_baseline("c3", "grad", 4.0)
    eng.close()

    if do_extra:
        # C4: LSST-shaped HMC gradient evaluations -- weak (1e4 chains per GPU) and BASELINE's split (1e4 chains over N GPUs)
        n4 = WORKLOADS["c4"][2]
        p4, eng4, data4 = make_engine(D, "c4", args.path)
        rec4 = lnp_record(D, args, "c4", "grad", n4, 10, 3, "c4_lnp_grad_weak", p=p4, eng=eng4)
        if world > 1:
            ns = (n4 + world - 1) // world
            rs = measure_lnp(D, p4, eng4, ns, "grad", 10, 3, want_e2e=False, sample_clocks=False)
            rec4["split"] = {"name": "c4_lnp_grad_1e4_chains_over_N_gpus", "value": ns * world * 10 / (rs["ms_per_step"] * 10 * 1e-3),
                             "unit": "evals/s", "ms_per_step": rs["ms_per_step"], "chains_per_gpu": ns, "scaling": "strong"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            rec4["cpu_baseline"] = reference_cpu_baseline("c4", "grad", 4.0)
        eng4.close()
        extra.append(rec4)
        extra.append(train_line(D, args))
        extra.append(c1_record(D, args))

    if rank == 0:
        line = {"metric": METRIC if args.mode == "lnp" else METRIC_GRAD,
                "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "mode": args.mode, "walkers_per_gpu": n, "n_in": p.n_in, "n_out": p.n_out,
                           "flops_per_eval": flops_eval, "kernel_path": r["kernel_path"],
                           "l2": "inputs rotate over %d distinct buffers; weights (%.1f MB) are L2-resident by design" % (
                               r["nbuf"], arch.n_params(p.kind, p.n_in, p.n_out) * 4 / 1e6)},
                "clocks": clocks,
                "e2e": {"value": total_evals / r["e2e_s"], "unit": "evals/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                        "buffers": "pageable numpy in / fresh numpy out (what an emcee/zeus caller holds)",
                        "pinned": total_evals / r["e2e_pinned_s"]},
                "gpu_launches": r["launches"], "roofline": roofline, "cpu_baseline": cpu}
        if sustained is not None:
            line["sustained"] = sustained
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    D.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
