"""Deterministic synthetic emulator problems (SURVEY 8d "Synthetic inputs").

Everything is drawn from ``numpy.random.default_rng`` (PCG64, stable across
machines and numpy versions) so that the golden-vector generator (which runs the
reference in the build container) and the GPU tests / bench (which run on a box
without the reference) see bit-identical weights and inputs.

No model arithmetic lives here: the data vector of a problem is either supplied
by the caller (golden files) or derived from a prediction made by the CUDA path
(``Problem.set_data_from_prediction``).
"""
import numpy as np

from . import arch


class Problem:
    """A bag of host arrays describing one emulator likelihood."""

    def __init__(self):
        self.kind = None
        self.n_in = self.n_out = 0
        self.state_dict = {}      # key -> float32 ndarray, nn.Linear layout [out, in]
        self.priors = []          # list of {'param','dist','arg1','arg2'}
        self.dolog10index = None
        self.ypositive = False
        self.X_mean = self.X_std = None
        self.y_mean = self.y_std = None
        self.cov = None           # float64 [n_out, n_out]
        self.sigma = None         # float64 sqrt(diag cov)
        self.inv_cov = None       # float64 np.linalg.inv(cov) (reference main.py:120)
        self.data = None          # float64 [n_out]
        self.temperature = 1.0
        self.theta0 = None

    def set_data_from_prediction(self, m0, noise_seed=2, noise=1.0):
        """data = m(theta0) + noise * sigma * N(0,1)   (SURVEY 8d)."""
        rng = np.random.default_rng(noise_seed)
        self.data = np.asarray(m0, np.float64) + noise * self.sigma * rng.standard_normal(self.n_out)
        return self.data


def xavier_uniform(rng, shape):
    fan_out, fan_in = shape
    bound = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def make_state_dict(kind, n_in, n_out, seed=0, skip_xavier=True, bias=1e-2, bias_jitter=0.0):
    """Weights shaped like the reference init (``linna/nn.py:34-43``, ``:91-108``):
    Xavier-uniform weights, bias 1e-2.  The reference zero-inits the skip
    weights; SURVEY 8d asks for Xavier there too so that the skip GEMMs do not
    multiply zeros (``skip_xavier``)."""
    rng = np.random.default_rng(seed)
    sd = {}
    for key, shp in arch.state_dict_shapes(kind, n_in, n_out):
        if key.endswith(".weight"):
            if key.endswith("skip_layer.weight") and not skip_xavier:
                sd[key] = np.zeros(shp, np.float32)
            elif key == "linearlayer.weight":
                sd[key] = np.full(shp, 1e-5, np.float32)
            else:
                sd[key] = xavier_uniform(rng, shp)
        else:
            if key == "linearlayer.bias":
                sd[key] = np.zeros(shp, np.float32)
            else:
                b = np.full(shp, bias, np.float32)
                if bias_jitter:
                    b = b + (bias_jitter * rng.standard_normal(shp)).astype(np.float32)
                sd[key] = b
    return sd


def make_problem(n_in, n_out, kind="ChtoModelv2", seed=0, priors="flat", log10=False,
                 ypositive=False, temperature=1.0, cond=None, bias_jitter=0.05):
    """Build one synthetic problem.

    priors: "flat" (all flat on [-5,5], README.rst:73-80), "mixed" (60 % gauss) or
            "gauss".
    log10:  mark parameters 0 and 1 as ``dolog10index`` (cosmolike_run.py:320) and
            give them positive flat priors.
    cond:   if given, stretch the covariance spectrum to roughly this condition
            number (SURVEY 8d "realistic 1e4-1e6 conditioning").
    """
    p = Problem()
    p.kind, p.n_in, p.n_out = kind, n_in, n_out
    p.state_dict = make_state_dict(kind, n_in, n_out, seed=seed, bias_jitter=bias_jitter)
    rng = np.random.default_rng(seed + 1000)

    # priors
    p.priors = []
    for i in range(n_in):
        if priors == "flat":
            pr = dict(param="p%d" % i, dist="flat", arg1=-5.0, arg2=5.0)
        elif priors == "gauss":
            pr = dict(param="p%d" % i, dist="gauss", arg1=float(rng.normal()), arg2=float(rng.uniform(0.5, 2.0)))
        else:
            if rng.uniform() < 0.6:
                pr = dict(param="p%d" % i, dist="gauss", arg1=float(rng.normal()),
                          arg2=float(rng.uniform(0.5, 2.0)))
            else:
                pr = dict(param="p%d" % i, dist="flat", arg1=float(rng.uniform(-5, -1)),
                          arg2=float(rng.uniform(1, 5)))
        p.priors.append(pr)
    if log10:
        p.dolog10index = [0, 1]
        for i in p.dolog10index:
            p.priors[i] = dict(param="p%d" % i, dist="flat", arg1=0.5, arg2=5.0)
    p.ypositive = bool(ypositive)

    # input normalisation: identity unless log10, where we centre a bit
    p.X_mean = np.zeros(n_in, np.float32)
    p.X_std = np.ones(n_in, np.float32)
    if log10:
        p.X_mean[:2] = 0.3
        p.X_std[:2] = 0.25

    # output normalisation
    if ypositive:
        p.y_mean = (0.1 * rng.standard_normal(n_out)).astype(np.float32)
        p.y_std = rng.uniform(0.05, 0.2, n_out).astype(np.float32)
    else:
        p.y_mean = rng.standard_normal(n_out).astype(np.float32)
        p.y_std = rng.uniform(0.5, 2.0, n_out).astype(np.float32)

    # covariance  C = A A^T / n_out + I   (rng seeded 0 in SURVEY; here seed-derived)
    crng = np.random.default_rng(seed)
    A = crng.standard_normal((n_out, n_out))
    C = A @ A.T / n_out + np.eye(n_out)
    if cond is not None:
        w, V = np.linalg.eigh(C)
        t = (w - w.min()) / max(w.max() - w.min(), 1e-300)
        w2 = np.exp(np.log(1.0) + t * np.log(cond))
        C = (V * w2) @ V.T
        C = 0.5 * (C + C.T)
    p.cov = C
    p.sigma = np.sqrt(np.diag(C))
    p.inv_cov = np.linalg.inv(C)
    p.temperature = float(temperature)

    # fiducial point in physical space: centre of each prior
    th = []
    for pr in p.priors:
        if pr["dist"] == "flat":
            th.append(0.5 * (pr["arg1"] + pr["arg2"]))
        else:
            th.append(pr["arg1"])
    p.theta0 = np.asarray(th, np.float64)
    return p


def walkers(n, n_in, scale=0.3, seed=1):
    """Latent-space walker positions u ~ N(0, scale^2 I)."""
    rng = np.random.default_rng(seed)
    return (scale * rng.standard_normal((n, n_in))).astype(np.float32)


def training_set(problem, n, seed=3, spread=0.5):
    """Synthetic (theta, target) inputs for the training configs: theta drawn
    around theta0 inside the flat priors.  Targets are filled by the caller from
    a prediction (they need the model)."""
    rng = np.random.default_rng(seed)
    th = np.empty((n, problem.n_in), np.float64)
    for i, pr in enumerate(problem.priors):
        if pr["dist"] == "flat":
            lo, hi = pr["arg1"], pr["arg2"]
            c, w = 0.5 * (lo + hi), 0.5 * (hi - lo)
            th[:, i] = c + spread * w * rng.uniform(-1, 1, n)
        else:
            th[:, i] = pr["arg1"] + spread * pr["arg2"] * rng.standard_normal(n)
    return th
