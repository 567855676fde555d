"""cuda_zstd_batch_{compress,decompress}_sharded: one batch sharded by chunk index over the GPUs of this process, one host
thread per shard, sizes exchanged device to device and scanned on every device.  Runs with two shards on ONE device
everywhere (the exchange then is a plain device copy) and with one shard per GPU when the box has more than one."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(oracle, libzstd, pkg, devices):
    chunk, total, level = 65536, 1200, 3
    G = len(devices)
    lib = pkg.load_library()
    data = oracle.gen_batch(chunk, total, 2, 0)
    bounds = [pkg.shard_range(total, g, G) for g in range(G)]
    keep, shards = [], (pkg.ShardC * G)()
    for g, dev in enumerate(devices):
        lo, hi = bounds[g]
        n = hi - lo
        with torch.cuda.device(dev):
            codec = pkg.ZstdBatchCodec(level=level)
            d_in = torch.from_numpy(data[lo * chunk: hi * chunk]).to(f"cuda:{dev}")
            stride = (codec.max_compressed_size(chunk) + 15) // 16 * 16
            d_out = torch.zeros(n * stride, dtype=torch.uint8, device=f"cuda:{dev}")
            idx = np.arange(n, dtype=np.uint64)
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int64)).to(f"cuda:{dev}")      # noqa: E731
            in_ptrs, in_sizes = t(np.uint64(d_in.data_ptr()) + idx * np.uint64(chunk)), t(np.full(n, chunk))
            out_ptrs, out_sizes = t(np.uint64(d_out.data_ptr()) + idx * np.uint64(stride)), t(np.full(n, stride))
            status = torch.full((n,), 77, dtype=torch.int32, device=f"cuda:{dev}")
            ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device=f"cuda:{dev}")
            all_sizes = torch.zeros(total, dtype=torch.int64, device=f"cuda:{dev}")
            all_offs = torch.zeros(total + 1, dtype=torch.int64, device=f"cuda:{dev}")
            stream = torch.cuda.Stream(device=dev)
        keep.append((codec, d_in, d_out, in_ptrs, in_sizes, out_ptrs, out_sizes, status, ws, all_sizes, all_offs, stream, stride, n))
        s = shards[g]
        s.device, s.mgr = dev, codec.h
        s.d_in_ptrs, s.d_in_sizes, s.d_out_ptrs, s.d_out_sizes = in_ptrs.data_ptr(), in_sizes.data_ptr(), out_ptrs.data_ptr(), out_sizes.data_ptr()
        s.d_statuses, s.num_chunks, s.d_temp, s.temp_bytes = status.data_ptr(), n, ws.data_ptr(), ws.numel()
        s.stream, s.d_all_sizes, s.d_all_offsets = stream.cuda_stream, all_sizes.data_ptr(), all_offs.data_ptr()
    torch.cuda.synchronize()
    assert lib.cuda_zstd_batch_compress_sharded(shards, G) == 0
    sizes_ref = None
    for g in range(G):
        codec, d_in, d_out, in_ptrs, in_sizes, out_ptrs, out_sizes, status, ws, all_sizes, all_offs, stream, stride, n = keep[g]
        assert int(status.max().item()) == 0
        a = all_sizes.cpu().numpy()
        lo, hi = bounds[g]
        assert np.array_equal(a[lo:hi], out_sizes.cpu().numpy())                      # my own slice sits at my base index
        if sizes_ref is None:
            sizes_ref = a
        assert np.array_equal(a, sizes_ref)                                           # every shard sees the same global table
        assert np.array_equal(all_offs.cpu().numpy(), np.concatenate([[0], np.cumsum(a)]))
        oh = d_out.cpu().numpy()
        for i in (0, n // 2, n - 1):
            f = oh[i * stride: i * stride + int(a[lo + i])]
            assert np.array_equal(libzstd.decompress(f, chunk), data[(lo + i) * chunk:(lo + i + 1) * chunk])
    # decompress the shards in place of their inputs: frames -> a second buffer, compared with the original
    backs = []
    for g, dev in enumerate(devices):
        codec, d_in, d_out, in_ptrs, in_sizes, out_ptrs, out_sizes, status, ws, all_sizes, all_offs, stream, stride, n = keep[g]
        with torch.cuda.device(dev):
            back = torch.zeros(n * chunk, dtype=torch.uint8, device=f"cuda:{dev}")
            idx = np.arange(n, dtype=np.uint64)
            bp = torch.from_numpy((np.uint64(back.data_ptr()) + idx * np.uint64(chunk)).astype(np.int64)).to(f"cuda:{dev}")
            bsz = torch.full((n,), chunk, dtype=torch.int64, device=f"cuda:{dev}")
            fsz = out_sizes.clone()
        backs.append((back, bp, bsz, fsz))
        s = shards[g]
        s.d_in_ptrs, s.d_in_sizes, s.d_out_ptrs, s.d_out_sizes = out_ptrs.data_ptr(), fsz.data_ptr(), bp.data_ptr(), bsz.data_ptr()
        s.d_all_sizes, s.d_all_offsets = None, None
    torch.cuda.synchronize()
    assert lib.cuda_zstd_batch_decompress_sharded(shards, G) == 0
    for g in range(G):
        assert torch.equal(backs[g][0], keep[g][1])
        assert bool((backs[g][2] == chunk).all().item())
    # a damaged frame in one shard: overall 1, the other shard's results intact
    out0 = keep[0][2]
    out0[3] ^= 0xFF
    assert lib.cuda_zstd_batch_decompress_sharded(shards, G) == 1
    assert int(keep[0][7][0].item()) != 0 and int(keep[-1][7].max().item()) == (0 if G > 1 else int(keep[-1][7].max().item()))


def test_two_shards_on_one_device(oracle, libzstd, pkg):
    _run(oracle, libzstd, pkg, [0, 0])


def test_one_shard_per_gpu(oracle, libzstd, pkg):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    _run(oracle, libzstd, pkg, list(range(min(torch.cuda.device_count(), 4))))
