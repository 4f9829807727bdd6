"""Training hand-off of one ``ml_sampler`` iteration (linna/train_gpu.py): load ``model_pickle.pkl`` (the
``train_NN`` function, pickled by path) and ``model_args.pkl`` from the iteration directory, train, write
``finish.pkl``.  The reference starts this file as a separate process (``os.system`` / ``srun``,
linna/main.py:199-257); here ``main(outdir)`` is called in-process -- same files in, same files out."""
import os
import pickle
import sys


def main(outdir, device="cuda"):
    import torch
    if device == "cuda" and not torch.cuda.is_available():
        raise RuntimeError("train_gpu: no CUDA device -- linna_b200 has no CPU training path")
    with open(os.path.join(outdir, "model_pickle.pkl"), "rb") as f:
        model = pickle.load(f)
    with open(os.path.join(outdir, "model_args.pkl"), "rb") as f:
        args = pickle.load(f)
    model(*args)
    with open(os.path.join(outdir, "finish.pkl"), "wb") as f:
        pickle.dump([True], f)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "cuda")
