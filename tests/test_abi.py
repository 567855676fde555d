"""The C-ABI library loads, exports every symbol include/*.h declares, and answers the size queries
and status maps without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("cuda_zstd_batch_c.h", "cuda_zstd_manager.h", "cuda_zstd_nvcomp.h", "pipeline_manager.hpp", "cuda_zstd_hybrid.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        for block in re.findall(r'extern "C" \{(.*?)\n\}', text, flags=re.S):
            names.update(re.findall(r"\b((?:cuda_zstd|nvcomp_zstd)_[a-z0-9_]+)\s*\(", block))
    return sorted(names)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    syms = declared_symbols()
    assert len(syms) >= 41
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(pkg.binding.EXPORTS) == syms


def test_size_queries_and_status_map_without_gpu(pkg, oracle):
    lib = pkg.load_library()
    h = lib.cuda_zstd_batch_create(3, 0)
    assert h
    assert lib.cuda_zstd_batch_get_max_compressed_size(h, 65536) == 66308 == oracle.max_compressed_size(65536)
    assert lib.cuda_zstd_batch_get_max_compressed_size(h, 131072) == 132101
    sizes = np.full(16384, 65536, dtype=np.uint64)
    ct = lib.cuda_zstd_batch_get_compress_temp_size(h, sizes.ctypes.data, 16384)
    dt = lib.cuda_zstd_batch_get_decompress_temp_size(h, sizes.ctypes.data, 16384)
    assert 0 < dt and 0 < ct                           # (a compress workspace always suffices for decompress: GPU tests)
    assert ct < 8 << 30 and dt < 6 << 30               # the reference needs ~213 GB / 61 GB for this batch
    real = np.full(16384, 12157, dtype=np.uint64)      # libzstd L3 frames of the P=0.50 class
    assert lib.cuda_zstd_batch_get_decompress_temp_size(h, real.ctypes.data, 16384) < 2200 << 20
    assert lib.cuda_zstd_batch_get_compress_temp_size(h, sizes.ctypes.data, 0) == 0
    assert lib.cuda_zstd_batch_get_max_compressed_size(None, 65536) == 0        # null handle -> 0 (nvcomp.cpp:815-829)
    lib.cuda_zstd_batch_destroy(h)
    # reference's lossy status -> int map (tests/test_nvcomp_interface.cu:637-668)
    for s in range(0, 29):
        assert pkg.status_to_nvcomp_error(s) == (s if s in (0, 2, 3, 4, 6, 7, 10, 12) else 1)
    assert lib.cuda_zstd_get_error_string(7) == b"ERROR_BUFFER_TOO_SMALL"
    assert lib.cuda_zstd_is_error(0) == 0 and lib.cuda_zstd_is_error(6) == 1
    # empty batch is a success; null tables are invalid parameters (nvcomp.cpp:305-310)
    h = lib.cuda_zstd_batch_create(3, 0)
    assert lib.cuda_zstd_batch_compress(h, None, None, 0, None, None, None, 0, None) == 0
    assert lib.cuda_zstd_batch_compress(h, None, None, 4, None, None, None, 0, None) == 2
    assert lib.cuda_zstd_batch_decompress(h, None, None, 4, None, None, None, 0, None) == 2
    lib.cuda_zstd_batch_destroy(h)
    m = lib.cuda_zstd_create_manager(99)               # out-of-range level is replaced, creation succeeds
    assert m
    assert lib.cuda_zstd_get_compress_workspace_size(m, 65536) > 0
    lib.cuda_zstd_destroy_manager(m)
    assert lib.cuda_zstd_set_dictionary(None, None) != 0


def test_product_does_not_touch_oracle_or_libzstd():
    # the shipped library must not link libzstd nor anything under oracle/
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "custom-nvcomp-with-zstd_b200", "libcuda_zstd_b200.so")],
                         capture_output=True, text=True).stdout
    assert "libzstd" not in out and "oracle" not in out
    for dirpath, _, files in os.walk(os.path.join(ROOT, "custom-nvcomp-with-zstd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "ORACLEPATH") or f in ("zstd_encode_core.cuh",), f
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "libzstd.so" not in text, f
