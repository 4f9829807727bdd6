"""Where the time of one convergence check goes (linna_b200/sampler.py: flush_and_check) at the C3 scale:
10^5 walkers x 100 new steps x 30 parameters.  Is the cuFFT part of the autocorrelation time worth a hand-written kernel?"""
import os, sys, time, tempfile
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linna_b200 import sampler as S
import linna.util as U

nw, nd, nstep = int(sys.argv[1]) if len(sys.argv) > 1 else 100000, 30, 100
priors = [dict(param="p%d" % i, dist="flat", arg1=-5.0, arg2=5.0) for i in range(nd)]
tr = U.Transform(priors)
tmp = tempfile.mkdtemp()
store = S.ChainStore(os.path.join(tmp, "chemcee_256.h5"), tr, fresh=True)
g = torch.Generator(device="cuda").manual_seed(0)
def sync(): torch.cuda.synchronize()
T = {}
for rep in range(3):          # three checks: the chain grows to 300 steps
    xb = torch.randn(nstep, nw, nd, device="cuda", generator=g)
    # AR(1) along steps so that tau is not degenerate
    for t in range(1, nstep): xb[t] = 0.9 * xb[t - 1] + 0.436 * xb[t]
    lb = torch.randn(nstep, nw, device="cuda", generator=g)
    sync(); t0 = time.perf_counter()
    tb = S.transform_block_device(tr, xb)
    xh, lh, th = S.to_host_pinned("x", xb), S.to_host_pinned("lnp", lb), S.to_host_pinned("theta", tb)
    t1 = time.perf_counter()
    store.extend(xh, lh, transformed=th); store.save()
    t2 = time.perf_counter()
    thin = S.thin_for_tau(store.chain)
    t3 = time.perf_counter()
    # the FFT part alone, on the device (CUDA events)
    xd = torch.from_numpy(np.ascontiguousarray(thin)).cuda()
    sync(); t4 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = S._next_pow_two(xd.shape[0])
    x = xd.to(torch.float64); x = x - x.mean(dim=0, keepdim=True)
    for _rep in range(2):        # the first call at a new length creates the cuFFT plans; the second is the FFT itself
        e0.record()
        f = torch.fft.rfft(x, n=2 * n, dim=0)
        acf = torch.fft.irfft(f * f.conj(), n=2 * n, dim=0)[:xd.shape[0]]
        e1.record(); sync()
        if _rep == 0: fft_first_ms = e0.elapsed_time(e1)
    fft_ms = e0.elapsed_time(e1)
    t5 = time.perf_counter()
    tau = S.integrated_time(thin)
    t6 = time.perf_counter()
    keep = max(int(2 * np.mean(tau)), 2)
    ok = S.checkmeanstd(np.asarray(store.chain[-keep:]), 0.1, 0.1)
    t7 = time.perf_counter()
    T = dict(fft_first_call_ms=fft_first_ms, d2h=t1 - t0, store_extend=t2 - t1, thin=t3 - t2, h2d_thin=t4 - t3, fft_device_ms=fft_ms, integrated_time_total=t6 - t5,
             checkmeanstd=t7 - t6, total=(t3 - t0) + (t7 - t5), steps=store.iteration, tau=float(np.mean(tau)), thin_shape=thin.shape)
    print(rep, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in T.items()}, flush=True)
print("FFT share of the check: %.3f %%" % (100 * T["fft_device_ms"] * 1e-3 / T["total"]))
