import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def libzstd():
    from oracle.oracle import LibZstd
    return LibZstd()


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge
    return ge.import_package()


class Model:
    """Host-side model of the GPU compressor (tests/model/enc_model.cpp)."""

    def __init__(self):
        import __graft_entry__ as ge
        self.lib = C.CDLL(ge.build_model())
        self.lib.model_compress.restype = C.c_size_t
        self.lib.model_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]

    def compress(self, data: np.ndarray, level: int, checksum: bool = False) -> np.ndarray:
        data = np.ascontiguousarray(data)
        cap = data.size + data.size // 255 + 3 * (data.size // (128 * 1024) + 1) + 512
        dst = np.empty(cap, np.uint8)
        n = self.lib.model_compress(data.ctypes.data, data.size, dst.ctypes.data, cap, level, int(checksum))
        return dst[:n].copy()


@pytest.fixture(scope="session")
def model():
    return Model()


@pytest.fixture(scope="session")
def gpu_codec_factory(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")

    def make(level=3, checksum=False):
        return pkg.ZstdBatchCodec(level=level, checksum=checksum)
    return make
