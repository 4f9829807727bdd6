"""Re-export of ``linna_b200.predictor_gpu`` under the reference module path ``linna.predictor_gpu``."""
from linna_b200.predictor_gpu import *  # noqa: F401,F403
from linna_b200 import predictor_gpu as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
