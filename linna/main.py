"""Re-export of ``linna_b200.main`` under the reference module path ``linna.main``."""
from linna_b200.main import *  # noqa: F401,F403
from linna_b200 import main as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
