"""Reference import path `largesteps.optimize` -> B200 implementation (largesteps_b200.optimize)."""
from largesteps_b200.optimize import *  # noqa: F401,F403
from largesteps_b200 import optimize as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
