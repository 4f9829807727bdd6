"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference runnable on the GPU box.

The reference (chto/linna) is a pure-Python package: there is nothing to compile.  What `oracle/_ref/` holds is
therefore a verbatim copy of its package directory, made by this script from where the sources lie
(`/root/reference/linna/*.py`) into `oracle/_ref/linna/`.  `oracle/_ref/` is git-ignored (reference sources never
enter the history) but NOT gpurun-ignored, so the copy travels to the GPU box like a built `.so`, where
`bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time the reference's own `Log_prob.__call__`
(linna/util.py:990-1021) per walker and its training inner loop (linna/predictor_gpu.py:273-288) through
`oracle/ref_bench.py`.  The seven third-party modules the package imports but the path never executes (emcee, zeus,
h5py, ...) are stubbed by `oracle/refstubs.py`, exactly as for the golden-vector generator.

    python oracle/make_ref.py            # (re)create oracle/_ref/linna from /root/reference

Nothing under `linna_b200/` reads `oracle/_ref/`.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SRC = os.environ.get("LINNA_REFERENCE_ROOT", "/root/reference")


def have_ref():
    return os.path.isfile(os.path.join(REF_DIR, "linna", "util.py"))


def ensure_ref(force=False):
    """Copy <reference>/linna/*.py to oracle/_ref/linna/ when the reference checkout is present (the build
    container); on the GPU box the prebuilt copy is used as it is.  Returns True when oracle/_ref is usable."""
    src = os.path.join(SRC, "linna")
    if os.path.isdir(src) and (force or not have_ref()):
        dst = os.path.join(REF_DIR, "linna")
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(src)):
            if f.endswith(".py"):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
        with open(os.path.join(REF_DIR, "README"), "w") as fh:
            fh.write("verbatim copy of %s/*.py made by oracle/make_ref.py; git-ignored, not product source\n" % src)
    return have_ref()


if __name__ == "__main__":
    print("oracle/_ref ready" if ensure_ref(force=True) else "reference checkout not found at %s" % SRC)
