"""Reference import path `largesteps.geometry` -> B200 implementation (largesteps_b200.geometry)."""
from largesteps_b200.geometry import *  # noqa: F401,F403
from largesteps_b200 import geometry as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
