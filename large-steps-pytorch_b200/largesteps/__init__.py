"""Drop-in shim: `import largesteps` resolves to the B200 implementation (largesteps_b200).

Put `large-steps-pytorch_b200/` on PYTHONPATH *instead of* installing the reference package and the reference's
own scripts (scripts/main.py:8-10, Tutorial.ipynb) import these modules unchanged.
"""
from largesteps_b200 import __version__, reference_version  # noqa: F401
