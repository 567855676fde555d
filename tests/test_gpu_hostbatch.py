"""Host-resident batch calls of the C ABI (cuda_zstd_batch_decompress_host / cuda_zstd_batch_compress_host_packed): payloads in
host memory, staged in waves inside the library.  Checked against stock libzstd and the oracle in both directions; covers
pinned and pageable buffers, contiguous and scattered layouts, ragged sizes, damaged items and small workspaces."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ptrs(buf, offs):
    base = buf.data_ptr() if isinstance(buf, torch.Tensor) else buf.ctypes.data
    return (np.uint64(base) + np.asarray(offs, dtype=np.uint64)).astype(np.uint64)


@pytest.mark.parametrize("n,pinned", [(3000, True), (700, False)])
def test_decompress_host_matches_libzstd(oracle, libzstd, pkg, n, pinned):
    chunk = 65536
    codec = pkg.ZstdBatchCodec(level=3)
    data = oracle.gen_batch(chunk, n, 2, 0)                                 # mixed classes: raw, RLE and compressed blocks
    blob, offs, sizes = libzstd.compress_chunks(data, chunk, 3)
    h_comp = torch.from_numpy(blob.copy())
    h_out = torch.zeros(n * chunk, dtype=torch.uint8)
    if pinned:
        h_comp, h_out = h_comp.pin_memory(), h_out.pin_memory()
    sizes = sizes.astype(np.uint64)
    caps = np.full(n, chunk, np.uint64)
    ws = torch.empty(codec.host_decompress_temp_size(sizes, caps), dtype=torch.uint8, device="cuda")
    out_sizes = caps.copy()
    st = np.full(n, 77, np.uint32)
    rc = codec.decompress_host(_ptrs(h_comp, offs), sizes, _ptrs(h_out, np.arange(n) * chunk), out_sizes, ws, st)
    assert rc == 0 and (st == 0).all() and (out_sizes == chunk).all()
    assert np.array_equal(h_out.numpy(), data)
    assert codec.last_launch_count() > 0
    # too small a workspace is refused before anything is copied
    assert codec.decompress_host(_ptrs(h_comp, offs), sizes, _ptrs(h_out, np.arange(n) * chunk), caps.copy(), ws[:1 << 20], st) == 7


def test_decompress_host_scattered_ragged_and_damaged(oracle, libzstd, pkg):
    codec = pkg.ZstdBatchCodec(level=3)
    rng = np.random.default_rng(3)
    n = 300
    lens = rng.integers(1, 70000, n)
    chunks = [oracle.gen_batch(int(l), 1, 0, 40000, first_idx=i) for i, l in enumerate(lens)]
    frames = [libzstd.compress(c, 3) for c in chunks]
    frames[7] = frames[7].copy(); frames[7][0] ^= 0xFF                       # bad magic
    # scattered: every frame and every output in its own host array (no two adjacent)
    h_in = [torch.from_numpy(f.copy()) for f in frames]
    h_out = [torch.zeros(int(l) + 13, dtype=torch.uint8) for l in lens]
    in_ptrs = np.array([t.data_ptr() for t in h_in], np.uint64)
    out_ptrs = np.array([t.data_ptr() for t in h_out], np.uint64)
    sizes = np.array([f.size for f in frames], np.uint64)
    caps = np.array([int(l) + 13 for l in lens], np.uint64)
    ws = torch.empty(codec.host_decompress_temp_size(sizes, caps), dtype=torch.uint8, device="cuda")
    out_sizes = caps.copy()
    st = np.zeros(n, np.uint32)
    rc = codec.decompress_host(in_ptrs, sizes, out_ptrs, out_sizes, ws, st)
    assert rc == 1 and st[7] == pkg.Status.ERROR_INVALID_MAGIC and out_sizes[7] == 0
    for i in range(n):
        if i == 7:
            continue
        assert st[i] == 0 and out_sizes[i] == lens[i], i
        assert np.array_equal(h_out[i].numpy()[: int(lens[i])], chunks[i]), i


@pytest.mark.parametrize("level,chunk,n", [(3, 65536, 2500), (1, 65536, 1100), (9, 131072, 96)])
def test_compress_host_packed_decodes_everywhere(oracle, libzstd, pkg, level, chunk, n):
    codec = pkg.ZstdBatchCodec(level=level, checksum=(level == 9))
    data = oracle.gen_batch(chunk, n, 2, 0)
    h_in = torch.from_numpy(data.copy()).pin_memory()
    sizes = np.full(n, chunk, np.uint64)
    sizes[-1] = chunk - 777                                                  # ragged tail
    cap = int(sum(codec.max_compressed_size(int(s)) for s in sizes))
    h_packed = torch.zeros(cap, dtype=torch.uint8).pin_memory()
    offs = np.zeros(n + 1, np.uint64)
    st = np.full(n, 77, np.uint32)
    ws = torch.empty(codec.host_compress_temp_size(sizes), dtype=torch.uint8, device="cuda")
    rc = codec.compress_host_packed(_ptrs(h_in, np.arange(n) * chunk), sizes, h_packed, cap, offs, ws, st)
    assert rc == 0 and (st == 0).all()
    assert offs[0] == 0 and (np.diff(offs.astype(np.int64)) > 0).all()
    packed = h_packed.numpy()
    for i in list(range(0, n, max(1, n // 40))) + [n - 1]:
        f = packed[int(offs[i]): int(offs[i + 1])]
        want = data[i * chunk: i * chunk + int(sizes[i])]
        assert np.array_equal(libzstd.decompress(f, int(sizes[i])), want), i
        rc2, out = oracle.decompress(f, int(sizes[i]))
        assert rc2 == 0 and np.array_equal(out, want), i
    # and back through the host-resident decode: packed frames in, original bytes out
    fsz = np.diff(offs).astype(np.uint64)
    h_back = torch.zeros(n * chunk, dtype=torch.uint8).pin_memory()
    ws2 = torch.empty(codec.host_decompress_temp_size(fsz, sizes), dtype=torch.uint8, device="cuda")
    out_sizes = sizes.copy()
    assert codec.decompress_host(_ptrs(h_packed, offs[:-1]), fsz, _ptrs(h_back, np.arange(n) * chunk), out_sizes, ws2) == 0
    assert (out_sizes == sizes).all()
    assert np.array_equal(h_back.numpy()[: (n - 1) * chunk + int(sizes[-1])], data[: (n - 1) * chunk + int(sizes[-1])])
    # packed capacity too small -> 7
    assert codec.compress_host_packed(_ptrs(h_in, np.arange(n) * chunk), sizes, h_packed, 1000, offs, ws, st) == 7
