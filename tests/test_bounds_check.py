"""compute-sanitizer is closed on this pool, so the hand-computed shared-memory offsets of the warp-synchronous kernels are
validated by a debug build of the library (-DP64B_BOUNDS_CHECK: every shared-memory access of me_search_kernel, mb_encode_kernel
and mb_decode_kernel checked against the region its thread may touch) run over edge / corner macroblocks, every search mode and
range, ragged stream and pair counts, rate control and the decoder (tests/bounds_check_run.py)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_product_build_reports_no_bounds_checking():
    import ctypes as C
    from p64_b200 import _lib
    assert "p64b_debug_oob" in _lib.SIGNATURES


@pytest.mark.gpu
def test_no_shared_memory_access_leaves_its_region():
    from p64_b200 import build
    lib = build.build_bounds_lib()
    env = dict(os.environ, P64B_LIB=lib)
    r = subprocess.run([sys.executable, os.path.join(HERE, "bounds_check_run.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["violations"] == 0 and res["cases"] >= 24, res
