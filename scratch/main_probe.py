import sys, os, numpy as np, torch, tempfile, io, contextlib
sys.path.insert(0, '/root/repo')
from copy import deepcopy
from linna.main import ml_sampler_core
from linna.nn import ChtoModelv2
def theory(x, outdirs): return deepcopy(x[1])
np.random.seed(1); torch.manual_seed(1)
ndim = 3
means = np.array([0.3, -0.5, 0.8]); cov = np.diag([0.04, 0.09, 0.0225])
pri = [{"param": "t%d" % i, "dist": "flat", "arg1": -3.0, "arg2": 3.0} for i in range(ndim)]
ep = int(sys.argv[1]); nt = int(sys.argv[2]); temps = [float(t) for t in sys.argv[3].split(",")]
k = len(temps)
params = {"trainingoption": 1, "num_epochs": ep, "batch_size": 200}
out = tempfile.mkdtemp() + "/"
buf = io.StringIO()
ebuf = io.StringIO()
with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(ebuf):
    chain, lp = ml_sampler_core([nt]*k, [200]*k, [4]*(k-1)+[8], [10]*(k-1)+[25], [0.1]*(k-1)+[0.05], [0.3]*(k-1)+[0.2], [0.3]*(k-1)+[0.2], out, theory, pri, means, cov, means + 0.05, None, 32, "cuda", None, False, temps, params=params, method="emcee")
sd = np.sqrt(np.diag(cov))
print("len", len(chain), "dmean/sd", (chain.mean(axis=0) - means) / sd, "std ratio", chain.std(axis=0) / sd)
for i in range(k):
    z = np.load(os.path.join(out, "iter_%d" % i, "chemcee_256.npz"))
    c = z["chain_transformed"]; c = c[len(c)//2:].reshape(-1, ndim)
    print("iter", i, "steps", len(z["chain"]), "dmean/sd", (c.mean(axis=0) - means) / sd, "std/sd", c.std(axis=0) / sd)

import re
txt = ebuf.getvalue()
runs = txt.split("  0%|")
for r in runs[1:]:
    ls = re.findall(r"Train/val Loss: ([0-9.e+-]+), ([0-9.e+-]+)", r)
    if len(ls) > 5: print("training run:", len(ls), " ".join("%s/%s" % (a[:7], b[:7]) for a, b in ls[::20]))
print([l for l in buf.getvalue().splitlines() if "learning" in l.lower() or "bad" in l.lower() or "early" in l.lower() or "decay" in l.lower()][:20])
for i in range(k):
    d = os.path.join(out, "iter_%d" % i)
    print(i, "lr.npy", np.load(os.path.join(d, "lr.npy")) if os.path.isfile(os.path.join(d, "lr.npy")) else None)
for i in range(k):
    d = os.path.join(out, "iter_%d" % i)
    x = np.loadtxt(os.path.join(d, "train_samples_x.txt")); y = np.load(os.path.join(d, "train_samples_y.npy"))
    print(i, "train x mean", x.mean(axis=0).round(3), "std", x.std(axis=0).round(3), "n", len(x), "unique", len(np.unique(x, axis=0)), "y==x", np.allclose(x, y))
    for nm in ("X_transform.pkl", "y_transform.pkl"):
        import linna.util as U
        with open(os.path.join(d, nm), "rb") as f:
            o = U.CPU_Unpickler(f).load()
        print("   ", nm, {kk: (np.asarray(v.detach()).round(4).tolist() if hasattr(v, "detach") else v) for kk, v in vars(o).items() if kk not in ("dev", "device")})

import shutil
dst = "/root/repo/gpurun_out/mp_out"
shutil.rmtree(dst, ignore_errors=True)
os.makedirs(dst)
for i in range(k):
    os.makedirs(os.path.join(dst, "iter_%d" % i))
    for f in ("train_samples_x.txt", "train_samples_y.npy", "val_samples_x.txt", "val_samples_y.npy"):
        shutil.copy(os.path.join(out, "iter_%d" % i, f), os.path.join(dst, "iter_%d" % i, f))
