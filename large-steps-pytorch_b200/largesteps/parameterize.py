"""Reference import path `largesteps.parameterize` -> B200 implementation (largesteps_b200.parameterize)."""
from largesteps_b200.parameterize import *  # noqa: F401,F403
from largesteps_b200 import parameterize as _impl
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
