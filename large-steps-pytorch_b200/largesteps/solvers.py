"""Reference import path `largesteps.solvers` -> B200 implementation (largesteps_b200.solvers)."""
from largesteps_b200.solvers import *  # noqa: F401,F403
from largesteps_b200 import solvers as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
