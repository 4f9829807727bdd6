"""Re-export of ``linna_b200.util`` under the reference module path ``linna.util``."""
from linna_b200.util import *  # noqa: F401,F403
from linna_b200 import util as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
