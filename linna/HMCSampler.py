"""Re-export of ``linna_b200.HMCSampler`` under the reference module path ``linna.HMCSampler``."""
from linna_b200.HMCSampler import *  # noqa: F401,F403
from linna_b200 import HMCSampler as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
