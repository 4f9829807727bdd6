import os, sys, io, re, contextlib, shutil, tempfile, pickle
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import linna.util as U, linna.nn as RNN
src = os.environ.get("PROBE_SRC", "/root/repo/scratch/probe_data")
out = tempfile.mkdtemp() + "/"
shutil.copytree(src, out, dirs_exist_ok=True)
means = np.array([0.3, -0.5, 0.8]); cov = np.diag([0.04, 0.09, 0.0225]); icov = np.linalg.inv(cov)
lst = [out + "iter_0/", out + "iter_1/"]
if not os.environ.get("FIND_LR"):
    np.save(lst[1] + "lr.npy", float(sys.argv[2]) if len(sys.argv) > 2 else 7.2e-4)
ns = U.NN_samplerv1(lst[1], [[-3, 3]] * 3)
if os.environ.get("SEED"): torch.manual_seed(int(os.environ["SEED"]))
params = {"trainingoption": 1, "num_epochs": int(sys.argv[1]), "batch_size": 200}
ebuf = io.StringIO()
import tqdm.auto as TA
DESCS = []
class _Bar:
    def __init__(self, it): self.it = it
    def __iter__(self): return iter(self.it)
    def set_description(self, s): DESCS.append(s)
TA.tqdm = _Bar
if os.environ.get("PRETRAIN"):
    ns0 = U.NN_samplerv1(lst[0], [[-3, 3]] * 3)
    U.train_NN(ns0, cov, icov, np.sqrt(np.diag(cov)), lst[0], lst[:1], means, None, False, False, 2, 1.0, False, None, 1, RNN.ChtoModelv2, params, False)
    print("pretrain epochs", len(DESCS), DESCS[-1]); DESCS.clear()
    stage = int(os.environ.get("STAGE", "0"))
    if stage >= 1:
        model, yinv = U.retrieve_model(lst[0], 3, 3, RNN.ChtoModelv2)
    if stage >= 2:
        pri = [{"param": "t%d" % i, "dist": "flat", "arg1": -3.0, "arg2": 3.0} for i in range(3)]
        tf = U.Transform(pri)
        lp = U.Log_prob(torch.from_numpy(means.astype(np.float32)), torch.from_numpy(icov.astype(np.float32)), model, yinv, tf, 1.0, nograd=True, loglikelihoodfunc=U.gaussianlogliklihood)
        u0 = np.asarray(U.invTransform(pri)(means))
        print("lnp", lp(u0))
    if stage >= 3:
        from linna_b200 import sampler as S
        es = S.EnsembleSampler(32, 3, lp)
        es.run_mcmc(u0 + 0.1 * np.random.randn(32, 3), 50)
        print("sampled", es.get_chain().shape)
    if stage >= 4:
        it = es.sample(es.get_chain()[-1], 1000)
        for n_, _ in enumerate(it):
            if n_ == 20: break
        print("broke out of generator; grad enabled:", torch.is_grad_enabled())
U.train_NN(ns, cov, icov, np.sqrt(np.diag(cov)), lst[1], lst, means, None, False, False, 2, 1.0, False, None, 1, RNN.ChtoModelv2, params, False)
ls = re.findall(r"Train/val Loss: ([0-9.e+-]+), ([0-9.e+-]+)", "\n".join(DESCS))
print("ours epochs", len(ls))
print(" ".join("%.3g/%.3g" % (float(a), float(b)) for a, b in ls[::10]))
print("final val", ls[-1])
print("lr.npy", np.load(lst[1] + "lr.npy"))
for nm in ():
    with open(lst[1] + nm, "rb") as f:
        o = U.CPU_Unpickler(f).load()
    print(nm, {k: (v.numpy().round(5).tolist() if hasattr(v, "numpy") else v) for k, v in vars(o).items() if k not in ("dev", "device")})
