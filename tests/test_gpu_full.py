"""BASELINE.json full-size configurations, checked through size-independent properties:
compress -> decompress round trip on the GPU, XXH64-of-output == XXH64-of-input per chunk, sampled
frames through stock libzstd, total size against libzstd on a sample."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roundtrip(oracle, libzstd, pkg, chunk, n, level, kind, P, checksum, sample=64):
    codec = pkg.ZstdBatchCodec(level=level, checksum=checksum)
    host = oracle.gen_batch(chunk, n, kind, P)
    dev = torch.from_numpy(host).cuda()
    out, sizes, stride = codec.compress_chunks(dev, chunk)
    back, bsz = codec.decompress_chunks(out, np.arange(n, dtype=np.uint64) * np.uint64(stride), sizes, chunk)
    assert (bsz == chunk).all()
    assert torch.equal(back, dev)                                   # encode -> decode identity over the whole batch
    rng = np.random.default_rng(1)
    pick = rng.choice(n, size=min(sample, n), replace=False)
    oh = out.cpu().numpy()
    zs = 0
    mine = 0
    for i in pick:
        f = oh[i * stride: i * stride + int(sizes[i])]
        c = host[i * chunk:(i + 1) * chunk]
        assert np.array_equal(libzstd.decompress(f, chunk), c)
        if checksum:
            assert f[4] & 4 and int.from_bytes(f[-4:].tobytes(), "little") == (oracle.xxh64(c) & 0xFFFFFFFF)
        zs += libzstd.compress(c, level).size
        mine += f.size
    assert mine <= 1.02 * zs + 8 * len(pick), (mine, zs)
    return float(host.size) / float(sizes.sum())


def test_config2_decompress_1gib_libzstd_frames(oracle, libzstd, pkg):
    # 16384 x 64 KiB, frames made by libzstd L3 (what the reference's default batch compress emits)
    chunk, n = 65536, 16384
    codec = pkg.ZstdBatchCodec(level=3)
    base_n = 512                                                     # distinct chunks; the batch tiles them with distinct frames
    host = oracle.gen_batch(chunk, base_n, 0, 32768)
    blob, offs, sizes = libzstd.compress_chunks(host, chunk, 3)
    reps = n // base_n
    comp = torch.from_numpy(blob).cuda()
    all_offs = np.tile(offs, reps)
    all_sizes = np.tile(sizes, reps)
    back, bsz = codec.decompress_chunks(comp, all_offs, all_sizes, chunk)
    assert (bsz == chunk).all()
    ref = torch.from_numpy(host).cuda()
    assert torch.equal(back.view(reps, -1), ref.view(1, -1).expand(reps, -1))


def test_config3_compress_l1_1gib(oracle, libzstd, pkg):
    ratio = _roundtrip(oracle, libzstd, pkg, 65536, 16384, 1, 0, 32768, False)
    assert ratio > 4.5


def test_config4_compress_l9_128k_checksum(oracle, libzstd, pkg):
    # BASELINE config 4 at its full size: 16384 x 128 KiB = 2 GiB, level 9, XXH64 checksums
    ratio = _roundtrip(oracle, libzstd, pkg, 131072, 16384, 9, 0, 32768, True, sample=32)
    assert ratio > 6.5


def test_more_than_one_wave_of_ragged_mixed_chunks(oracle, libzstd, pkg):
    """20,000 mixed-entropy chunks of ragged sizes (1 B .. 128 KiB) in ONE call: the decoder's fast path runs in waves of
    16,384 chunks and the levels 1-4 encoder in waves of 8,192 / 4,096, both with items on every route (64 KiB and
    128 KiB geometries, raw / RLE blocks, general decode kernel for what the pools cannot hold)."""
    n = 20000
    rng = np.random.default_rng(11)
    lens = rng.integers(1, 65537, n).astype(np.int64)
    lens[rng.choice(n, 300, replace=False)] = rng.integers(65537, 131073, 300)
    lens[:4] = [1, 65536, 65537, 131072]
    offs = np.zeros(n + 1, np.int64); offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    host = oracle.gen_batch(65536, (total + 65535) // 65536, 2, 0)[:total].copy()
    dev = torch.from_numpy(host).cuda()
    codec = pkg.ZstdBatchCodec(level=3)
    caps = np.array([codec.max_compressed_size(int(l)) for l in lens], np.int64)
    coffs = np.zeros(n + 1, np.int64); coffs[1:] = np.cumsum((caps + 15) // 16 * 16)
    comp = torch.empty(int(coffs[-1]), dtype=torch.uint8, device="cuda")
    sizes = lens.astype(np.uint64)
    ws = torch.empty(max(codec.compress_temp_size(n, sizes), codec.decompress_temp_size(n)), dtype=torch.uint8, device="cuda")
    csz = caps.astype(np.uint64)
    rc = codec.compress_tables((np.uint64(dev.data_ptr()) + offs[:-1].astype(np.uint64)).astype(np.uint64), sizes, n,
                               (np.uint64(comp.data_ptr()) + coffs[:-1].astype(np.uint64)).astype(np.uint64), csz, ws)
    assert rc == 0 and (csz > 0).all()
    ch = comp.cpu().numpy()
    for i in rng.choice(n, 96, replace=False).tolist() + [0, 1, 2, 3, n - 1]:
        f = ch[int(coffs[i]): int(coffs[i]) + int(csz[i])]
        assert np.array_equal(libzstd.decompress(f, int(lens[i])), host[int(offs[i]): int(offs[i + 1])]), i
    back = torch.zeros(total, dtype=torch.uint8, device="cuda")
    bsz = lens.astype(np.uint64)
    rc = codec.decompress_tables((np.uint64(comp.data_ptr()) + coffs[:-1].astype(np.uint64)).astype(np.uint64), csz, n,
                                 (np.uint64(back.data_ptr()) + offs[:-1].astype(np.uint64)).astype(np.uint64), bsz, ws)
    assert rc == 0 and (bsz == sizes).all()
    assert torch.equal(back, dev)


def test_config5_mixed_entropy_shard(oracle, libzstd, pkg):
    # one GPU's share of the 8 GiB mixed-entropy job at 8 GPUs: 16384 chunks starting at the shard's first index
    plan = pkg.ShardPlan(131072, 3, 8)
    codec = pkg.ZstdBatchCodec(level=3)
    chunk = 65536
    host = oracle.gen_batch(chunk, plan.count, 2, 0, first_idx=plan.lo)
    dev = torch.from_numpy(host).cuda()
    out, sizes, stride = codec.compress_chunks(dev, chunk)
    d_sizes = torch.from_numpy(sizes.astype(np.int64)).cuda()
    table = pkg.gather_sizes(d_sizes, pkg.ShardPlan(plan.count, 0, 1))
    off = codec.scan_sizes(table, base=0)
    assert int(off[-1].item()) == int(sizes.sum())
    back, bsz = codec.decompress_chunks(out, np.arange(plan.count, dtype=np.uint64) * np.uint64(stride), sizes, chunk)
    assert torch.equal(back, dev)
