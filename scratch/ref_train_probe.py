"""Container-only: run the UNMODIFIED reference train_NN on a two-iteration toy data set (CPU)."""
import importlib.util, os, sys, io, re, contextlib, shutil, tempfile
import numpy as np, torch
REPO = "/root/repo"
sys.path = [p for p in sys.path if os.path.realpath(p or ".") != os.path.realpath(REPO)]
def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path); m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m; spec.loader.exec_module(m); return m
refstubs = _load("refstubs", os.path.join(REPO, "oracle", "refstubs.py"))
refstubs.import_reference()
import linna.util as U, linna.nn as RNN, linna.predictor_gpu as PG
DESCS = []
class _Bar:
    def __init__(self, it): self.it = it
    def __iter__(self): return iter(self.it)
    def set_description(self, s): DESCS.append(s)
PG.tqdm = _Bar
class _A:
    def __getattr__(self, n): return _A()
    def __call__(self, *a, **k): return _A()
    def __getitem__(self, i): return _A()
class _Plt(_A):
    def subplots(self, *a, **k): return _A(), _A()
PG.plt = _Plt(); U.plt = _Plt()
src = os.environ.get("PROBE_SRC", os.path.join(REPO, "scratch", "probe_data"))
out = tempfile.mkdtemp() + "/"
shutil.copytree(src, out, dirs_exist_ok=True)
means = np.array([0.3, -0.5, 0.8]); cov = np.diag([0.04, 0.09, 0.0225]); icov = np.linalg.inv(cov)
lst = [out + "iter_0/", out + "iter_1/"]
np.save(lst[1] + "lr.npy", float(sys.argv[2]) if len(sys.argv) > 2 else 7.2e-4)
ns = U.NN_samplerv1(lst[1], [[-3, 3]] * 3)
if os.environ.get("SEED"): torch.manual_seed(int(os.environ["SEED"]))
params = {"trainingoption": 1, "num_epochs": int(sys.argv[1]), "batch_size": 200}
ebuf = io.StringIO()
torch.set_num_threads(8)
with contextlib.redirect_stderr(ebuf):
  try:
    U.train_NN(ns, cov, icov, np.sqrt(np.diag(cov)), lst[1], lst, means, None, False, False, 2, 1.0, False, None, 1, RNN.ChtoModelv2, params, False)
  except TypeError as e:
    print("(plot stub)", e)
ls = re.findall(r"Train/val Loss: ([0-9.e+-]+), ([0-9.e+-]+)", "\n".join(DESCS))
seen = ls
print("reference epochs", len(seen))
print(" ".join("%.3g/%.3g" % (float(a), float(b)) for a, b in seen[::10]))
print("final val", seen[-1])
import pickle
for nm in ("X_transform.pkl", "y_transform.pkl", "y_invtransform_data.pkl"):
    with open(lst[1] + nm, "rb") as f:
        o = U.CPU_Unpickler(f).load()
    print(nm, {k: (v.numpy().round(5).tolist() if hasattr(v, "numpy") else v) for k, v in vars(o).items() if k not in ("dev", "device")})
