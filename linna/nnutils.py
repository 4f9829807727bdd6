"""Re-export of ``linna_b200.nnutils`` under the reference module path ``linna.nnutils``."""
from linna_b200.nnutils import *  # noqa: F401,F403
from linna_b200 import nnutils as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
