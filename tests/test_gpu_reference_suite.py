"""The reference's own batch-path test programs, compiled unchanged against this repo's include/ and library
(oracle/build_ref_tests.sh; sources stay under /root/reference, binaries under oracle/_ref/tests/), run on the GPU.

They exercise exactly the drop-in boundary of SURVEY.md 8(b): the C ABI (tests/test_c_api.cpp), NvcompV5BatchManager
(tests/test_nvcomp_batch.cu, tests/test_nvcomp_interface.cu), the inference API (tests/test_inference_api.cu) and the
manager-level round trips.  Nothing here reads /root/reference at run time."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "tests")
PROGRAMS = ["test_c_api", "test_c_api_edge_cases", "test_compressible_data", "test_concurrency_repro", "test_correctness",
            "test_extended_validation", "test_gpu_bitstream", "test_inference_api", "test_lz77_comprehensive",
            "test_metadata_roundtrip", "test_nvcomp_batch", "test_nvcomp_interface", "test_parallel_compression",
            "test_rfc8878_compliance", "test_roundtrip", "test_scale_repro", "test_two_phase_unit", "test_pipeline_integration"]
OUT_OF_SCOPE_SUBTESTS = {}          # (round 1 listed the dictionary round trip of test_c_api_edge_cases here)


@pytest.mark.gpu
@pytest.mark.parametrize("name", PROGRAMS)
def test_reference_program(name):
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("reference test programs not built (oracle/build_ref_tests.sh needs the reference tree)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "custom-nvcomp-with-zstd_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    # tests/test_pipeline_integration.cu:34-37 verifies its output with `zstd -q -t`; the image has libzstd but not the
    # command line tool, so tests/bin/zstd (stock libzstd behind the same command line) stands in for it
    env["PATH"] = os.path.join(ROOT, "tests", "bin") + ":" + env.get("PATH", "")
    r = subprocess.run([exe], cwd=BIN, env=env, capture_output=True, text=True, errors="replace", timeout=600)
    tail = (r.stdout[-3000:] + "\n--- stderr ---\n" + r.stderr[-2000:])
    if name in OUT_OF_SCOPE_SUBTESTS and r.returncode != 0:
        # the program also covers a feature SURVEY.md 8(f) ranks after this round (dictionaries: the symbols exist and
        # answer ERROR_NOT_IMPLEMENTED); every OTHER subtest of the program must pass
        failed = [ln.strip() for ln in r.stdout.splitlines() if ln.strip().endswith(": FAIL")]
        assert failed == [s + ": FAIL" for s in OUT_OF_SCOPE_SUBTESTS[name]], f"{name}: unexpected failures {failed}\n{tail}"
        return
    assert r.returncode != 77, f"{name} skipped itself\n{tail}"
    assert r.returncode == 0, f"{name} exited {r.returncode}\n{tail}"


@pytest.mark.gpu
def test_reference_python_suite():
    """The reference's own Python package: its pybind11 module (python/src/binding.cpp) compiled UNCHANGED against this
    repo's include/ + library (oracle/build_ref_python.sh) and its own python/tests/test_basic.py run against it.
    Pinned outcome: everything passes except TestHybridEngine::test_query_routing, which asserts that small host inputs
    are routed to the CPU codec -- this build has no CPU route by design -- and the two CuPy tests that
    skip themselves (CuPy is not in the image)."""
    import sys
    pydir = os.path.join(ROOT, "oracle", "_ref", "python")
    if not os.path.exists(os.path.join(pydir, "tests", "test_basic.py")):
        pytest.skip("reference Python module not built (oracle/build_ref_python.sh needs the reference tree)")
    env = dict(os.environ)
    env["PYTHONPATH"] = pydir + ":" + env.get("PYTHONPATH", "")
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "custom-nvcomp-with-zstd_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_basic.py", "-q", "-p", "no:cacheprovider", "-x", "--deselect",
                        "tests/test_basic.py::TestHybridEngine::test_query_routing"],
                       cwd=pydir, env=env, capture_output=True, text=True, errors="replace", timeout=900)
    tail = r.stdout[-3000:] + "\n--- stderr ---\n" + r.stderr[-1500:]
    assert r.returncode == 0, tail
    last = [ln for ln in r.stdout.splitlines() if " passed" in ln][-1]
    assert "73 passed" in last and "failed" not in last, tail
    # the deselected test fails exactly where expected: GPU_KERNELS instead of CPU_LIBZSTD
    r2 = subprocess.run([sys.executable, "-m", "pytest", "tests/test_basic.py::TestHybridEngine::test_query_routing", "-q", "-p",
                         "no:cacheprovider"], cwd=pydir, env=env, capture_output=True, text=True, errors="replace", timeout=300)
    assert r2.returncode != 0 and "GPU_KERNELS" in r2.stdout and "CPU_LIBZSTD" in r2.stdout, r2.stdout[-2000:]
