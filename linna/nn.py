"""Re-export of ``linna_b200.nn`` under the reference module path ``linna.nn``."""
from linna_b200.nn import *  # noqa: F401,F403
from linna_b200 import nn as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
