"""Checkpoint I/O with the reference's on-disk layout (``linna/nnutils.py:109-151``):
``<dir>/last.pth.tar`` = ``torch.save({'epoch', 'state_dict', 'optim_dict'[, 'mpi_state_dict']})``,
copied to ``best.pth.tar`` when the validation loss improved."""
import os
import shutil

import torch


def save_checkpoint(state, is_best, checkpoint):
    """Write ``last.pth.tar`` (and ``best.pth.tar`` if ``is_best``) under ``checkpoint``."""
    if not os.path.exists(checkpoint):
        print("Checkpoint Directory does not exist! Making directory {}".format(checkpoint))
        os.mkdir(checkpoint)
    last = os.path.join(checkpoint, "last.pth.tar")
    torch.save(state, last)
    if is_best:
        shutil.copyfile(last, os.path.join(checkpoint, "best.pth.tar"))


def load_checkpoint(checkpoint, model, optimizer=None, device=None, ismpi=False):
    """Load ``state_dict`` (and ``optim_dict`` when an optimizer object is given) from a checkpoint file."""
    if not os.path.exists(checkpoint):
        raise FileNotFoundError("File doesn't exist {}".format(checkpoint))
    ckpt = torch.load(checkpoint, map_location=device if device is not None else None, weights_only=False)
    model.load_state_dict(ckpt["mpi_state_dict" if ismpi else "state_dict"])
    if optimizer:
        optimizer.load_state_dict(ckpt["optim_dict"])
    return ckpt


class RunningAverage:
    """Streaming mean (linna/nnutils.py:48-68; unused by the hot path)."""

    def __init__(self):
        self.steps, self.total = 0, 0.0

    def update(self, val):
        self.total += val
        self.steps += 1

    def __call__(self):
        return self.total / float(self.steps)


for _f in (save_checkpoint, load_checkpoint, RunningAverage):
    _f.__module__ = "linna.nnutils"
del _f
