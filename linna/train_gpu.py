"""Re-export of ``linna_b200.train_gpu`` under the reference module path ``linna.train_gpu``."""
from linna_b200.train_gpu import *  # noqa: F401,F403
from linna_b200 import train_gpu as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
