"""TEST INFRASTRUCTURE ONLY.  Import shims that let the *unmodified* reference
(``/root/reference/linna``) be imported in the build container, where seven of its
third-party dependencies are absent.  None of the stubbed modules is
on the arithmetic path of the emulator likelihood -- that lives in PyTorch,
which is installed.  Used by ``tests/golden/make_golden.py`` and by the
container-only live-reference tests; never by the product or on the GPU box.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LINNA_REFERENCE_ROOT", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Register empty stand-ins for the missing modules (idempotent)."""
    if "linna_refstubs_installed" in sys.modules:
        return
    _mod("linna_refstubs_installed")

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    def _subplots(*a, **k):
        return _Anything(), _Anything()

    plt = _mod("matplotlib.pyplot", subplots=_subplots, figure=_Anything, plot=_Anything(),
               savefig=_Anything(), close=_Anything(), legend=_Anything(), xlabel=_Anything(),
               ylabel=_Anything(), switch_backend=_Anything())
    _mod("matplotlib", pyplot=plt, use=lambda *a, **k: None)
    _mod("torch_lr_finder", LRFinder=_Anything)

    class Move:  # base class of the reference's HMC/NUTS moves (sampler.py:101, :186)
        pass

    class HDFBackend:  # base class of the reference's Transformbackend (sampler.py:322)
        def __init__(self, *a, **k):
            pass

    class SaveProgressCallback:  # base of ZeusTransformCallback (sampler.py:556)
        def __init__(self, *a, **k):
            pass

    state = _mod("emcee.state", State=_Anything)
    moves = _mod("emcee.moves", Move=Move)
    backends = _mod("emcee.backends", HDFBackend=HDFBackend)
    _mod("emcee", state=state, moves=moves, backends=backends, EnsembleSampler=_Anything)
    autocorr = _mod("zeus.autocorr", AutoCorrTime=_Anything)
    callbacks = _mod("zeus.callbacks", SaveProgressCallback=SaveProgressCallback)
    _mod("zeus", autocorr=autocorr, callbacks=callbacks, EnsembleSampler=_Anything)
    _mod("h5py", File=_Anything)
    _mod("pyDOE2", lhs=_Anything())
    _mod("sample_generator")
    _mod("numdifftools", Hessian=_Anything)
    try:
        import scipy.misc  # noqa: F401  (nnutils.py:6)
    except Exception:
        import scipy
        scipy.misc = _mod("scipy.misc")


def import_reference():
    """Return the reference's ``linna`` package (util, predictor_gpu, nn, HMCSampler).

    Must run in a process where the repo's own ``linna`` shim is NOT importable
    first: the reference root is pushed to the front of ``sys.path`` and any
    already-imported ``linna*`` modules are dropped.
    """
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "linna")):
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    install_stubs()
    for k in [k for k in sys.modules if k == "linna" or k.startswith("linna.")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import linna.util as util            # noqa
        import linna.predictor_gpu as pg     # noqa
        import linna.nn as rnn               # noqa
        import linna.HMCSampler as hmc       # noqa
        import linna
    finally:
        sys.path.remove(REFERENCE_ROOT)
    assert os.path.realpath(linna.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    return linna
