"""Runs the native C++ test of the drop-in surface (tests/cpp/test_cpp_api.cu): ZstdBatchManager,
NvcompV5BatchManager, single-buffer and inference API, metadata helpers, status maps."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_surface():
    exe = os.path.join(ROOT, "tests", "cpp", "test_cpp_api")
    if not os.path.exists(exe):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__ as ge
        ge.build_cpp_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL PASSED" in r.stdout
