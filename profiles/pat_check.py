#!/usr/bin/env python
"""LS_PCG_PATTERN=1 python profiles/pat_check.py : the experimental pattern-only path's parity cases in one process
(the same script tests/test_gpu_zz_pattern_experimental.py runs in a subprocess)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LS_PCG_PATTERN", "1")
src = open(os.path.join(ROOT, "tests", "test_gpu_zz_pattern_experimental.py")).read()
case = src.split("CASE = r'''", 1)[1].split("'''", 1)[0]
exec(compile(case, "pattern_case", "exec"), {"ROOT": ROOT, "__name__": "__main__"})
