"""HybridEngine surface (include/cuda_zstd_hybrid.h; reference include/cuda_zstd_hybrid.h:88-343), GPU-only in this build,
through its C ABI: host and device buffers, frames checked against stock libzstd and the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

HOST, DEVICE, UNKNOWN = 0, 1, 3
GPU_KERNELS = 1


@pytest.fixture(scope="module")
def hybrid(pkg):
    h = pkg.ZstdHybrid(level=3)
    yield h
    h.close()


def _inputs(oracle):
    rng = np.random.default_rng(11)
    return {
        "tiny": np.frombuffer(b"The quick brown fox jumps over the lazy dog", np.uint8).copy(),
        "one": np.array([7], np.uint8),
        "text64k": oracle.gen_textlike(65536),
        "p50_128k": oracle.gen_batch(131072, 1, 0, 32768),
        "random100k": rng.integers(0, 256, 100_000, dtype=np.uint8),
        "zeros1m": np.zeros(1 << 20, np.uint8),
        "p50_3m+5": oracle.gen_batch(65536, 49, 0, 32768)[: 3 * (1 << 20) + 5].copy(),     # multi-block frame
    }


@pytest.mark.gpu
def test_host_to_host_round_trip_and_libzstd_cross_decode(hybrid, oracle, libzstd):
    for name, data in _inputs(oracle).items():
        n = data.size
        cap = hybrid.max_compressed_size(n)
        comp = np.empty(cap, np.uint8)
        rc, csize, res = hybrid.compress(data, n, comp, cap, HOST, HOST)
        assert rc == 0, (name, rc)
        assert 0 < csize <= cap and res.backend_used == GPU_KERNELS and res.input_bytes == n and res.output_bytes == csize
        frame = comp[:csize].copy()
        assert np.array_equal(libzstd.decompress(frame, n), data), name           # stock libzstd reads our frame
        orc_rc, orc_out = oracle.decompress(frame, n)
        assert orc_rc == 0 and np.array_equal(orc_out, data), name                # so does the RFC 8878 restatement
        out = np.empty(n, np.uint8)
        rc, dsize, _ = hybrid.decompress(frame, csize, out, n, HOST, HOST)
        assert rc == 0 and dsize == n and np.array_equal(out, data), name
        # libzstd's frame of the same data through the engine
        ref = libzstd.compress(data, 3)
        out[:] = 0
        rc, dsize, _ = hybrid.decompress(ref, ref.size, out, n, UNKNOWN, UNKNOWN)
        assert rc == 0 and dsize == n and np.array_equal(out, data), name


@pytest.mark.gpu
def test_device_and_mixed_locations(hybrid, oracle, libzstd):
    data = oracle.gen_batch(65536, 4, 0, 40000)
    n = data.size
    d_in = torch.from_numpy(data).cuda()
    cap = hybrid.max_compressed_size(n)
    d_comp = torch.empty(cap, dtype=torch.uint8, device="cuda")
    rc, csize, res = hybrid.compress(d_in, n, d_comp, cap, DEVICE, DEVICE)
    assert rc == 0 and res.input_location == DEVICE and res.output_location == DEVICE
    frame = d_comp[:csize].cpu().numpy()
    assert np.array_equal(libzstd.decompress(frame, n), data)
    # device frame -> host output, locations detected
    out = np.empty(n, np.uint8)
    rc, dsize, res = hybrid.decompress(d_comp, csize, out, n, UNKNOWN, UNKNOWN)
    assert rc == 0 and dsize == n and np.array_equal(out, data)
    assert res.input_location == DEVICE and res.output_location == HOST
    # host frame -> device output
    d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rc, dsize, _ = hybrid.decompress(frame, csize, d_out, n, HOST, DEVICE)
    assert rc == 0 and dsize == n and np.array_equal(d_out.cpu().numpy(), data)


@pytest.mark.gpu
def test_errors_and_routing(hybrid, pkg, oracle):
    S = pkg.Status
    data = oracle.gen_textlike(4096)
    comp = np.empty(hybrid.max_compressed_size(4096), np.uint8)
    assert hybrid.compress(None, 4096, comp, comp.size, HOST, HOST)[0] == S.ERROR_INVALID_PARAMETER      # hybrid.cu:783
    assert hybrid.compress(data, 0, comp, comp.size, HOST, HOST)[0] == S.ERROR_INVALID_PARAMETER
    rc, csize, _ = hybrid.compress(data, 4096, comp, comp.size, HOST, HOST)
    assert rc == 0
    small = np.empty(100, np.uint8)
    assert hybrid.decompress(comp, csize, small, small.size, HOST, HOST)[0] == S.ERROR_BUFFER_TOO_SMALL   # the binding's retry loop keys on this (binding.cpp:496)
    junk = np.arange(64, dtype=np.uint8)
    out = np.empty(4096, np.uint8)
    assert hybrid.decompress(junk, 64, out, out.size, HOST, HOST)[0] != 0
    # one route: every mode, size and location answers a GPU backend
    for mode in range(6):
        h = pkg.ZstdHybrid(level=1, mode=mode)
        assert h.query_routing(100) == GPU_KERNELS and h.query_routing(1 << 30, DEVICE, DEVICE, False) == GPU_KERNELS
        rc, cs, res = h.compress(data, 4096, comp, comp.size, HOST, HOST)
        assert rc == 0 and res.backend_used == GPU_KERNELS
        h.close()


@pytest.mark.gpu
def test_dictionary_calls_behave_like_the_reference_at_batch_sizes(oracle, libzstd, pkg):
    """cuda_zstd_train_dictionary / set_dictionary (src/cuda_zstd_c_api.cpp:128-195): the dictionary is accepted and kept;
    like the reference's host route at these sizes (src/cuda_zstd_manager.cu:1604-1668: plain ZSTD_compress) the frames
    carry no Dictionary_ID and decode anywhere -- with stock libzstd, without the dictionary."""
    import ctypes as C
    import torch
    lib = pkg.load_library()
    lib.cuda_zstd_train_dictionary.restype = C.c_void_p
    lib.cuda_zstd_train_dictionary.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    lib.cuda_zstd_destroy_dictionary.argtypes = [C.c_void_p]
    lib.cuda_zstd_set_dictionary.restype = C.c_int
    lib.cuda_zstd_set_dictionary.argtypes = [C.c_void_p, C.c_void_p]
    s1, s2 = oracle.gen_batch(16384, 1, 0, 30000), oracle.gen_batch(16384, 1, 0, 30000, first_idx=9)
    ptrs = (C.c_void_p * 2)(s1.ctypes.data, s2.ctypes.data)
    sizes = (C.c_size_t * 2)(s1.size, s2.size)
    assert not lib.cuda_zstd_train_dictionary(None, None, 0, 1024)
    d = lib.cuda_zstd_train_dictionary(ptrs, sizes, 2, 20000)
    assert d
    s = pkg.ZstdSingle(3)
    assert lib.cuda_zstd_set_dictionary(s.h, d) == 0
    assert lib.cuda_zstd_set_dictionary(s.h, None) != 0
    n = 65536
    x = oracle.gen_batch(n, 1, 0, 30000, first_idx=4)
    xd = torch.from_numpy(x).cuda()
    comp = torch.zeros(n * 2, dtype=torch.uint8, device="cuda")
    w = torch.empty(s.compress_workspace(n), dtype=torch.uint8, device="cuda")
    rc, csz = s.compress(xd, n, comp, comp.numel(), w, w.numel())
    assert rc == 0 and 0 < csz < n
    frame = comp.cpu().numpy()[:csz]
    assert frame[4] & 3 == 0                                           # no Dictionary_ID field
    assert np.array_equal(libzstd.decompress(frame, n), x)
    back = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rc, dsz = s.decompress(comp, csz, back, n, w, w.numel())
    assert rc == 0 and dsz == n and torch.equal(back, xd)
    s.close()
    lib.cuda_zstd_destroy_dictionary(d)
