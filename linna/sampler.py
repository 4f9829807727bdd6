"""Re-export of ``linna_b200.sampler`` under the reference module path ``linna.sampler``."""
from linna_b200.sampler import *  # noqa: F401,F403
from linna_b200 import sampler as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
