"""Host side of emulator training (placeholder until the fused training kernels land)."""


def run_training(*a, **k):
    raise NotImplementedError("fused training step not built yet")


def train_nn(*a, **k):
    raise NotImplementedError("fused training step not built yet")


def train_NN(*a, **k):
    raise NotImplementedError("fused training step not built yet")


train_nn.__module__ = "linna.util"
train_NN.__module__ = "linna.util"
