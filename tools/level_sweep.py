"""Compress / decompress throughput and size vs libzstd per level (device-resident, CUDA events)."""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle.oracle import Oracle, LibZstd
from concurrent.futures import ThreadPoolExecutor
pkg = ge.import_package(); orc = Oracle(); z = LibZstd()
out = []
for chunk, n, levels in ((65536, 8192, (3, 1, 3, 5, 9)), (131072, 4096, (9,))):
    host = np.concatenate(list(ThreadPoolExecutor(8).map(lambda i: orc.gen_batch(chunk, 256, 0, 32768, first_idx=i * 256), range(n // 256))))
    dev = torch.from_numpy(host).cuda()
    for lvl in levels:
        codec = pkg.ZstdBatchCodec(level=lvl, checksum=(lvl == 9))
        ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device='cuda')
        o, sizes, stride = codec.compress_chunks(dev, chunk, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o, sizes, stride = codec.compress_chunks(dev, chunk, ws); e1.record(); torch.cuda.synchronize()
        cms = e0.elapsed_time(e1)
        offs = np.arange(n, dtype=np.uint64) * np.uint64(stride)
        wsd = torch.empty(codec.decompress_temp_size(n, sizes), dtype=torch.uint8, device='cuda')
        back, _ = codec.decompress_chunks(o, offs, sizes, chunk, wsd)
        e0.record(); back, _ = codec.decompress_chunks(o, offs, sizes, chunk, wsd); e1.record(); torch.cuda.synchronize()
        dms = e0.elapsed_time(e1)
        assert torch.equal(back, dev)
        sample = range(0, n, n // 64)
        zs = sum(z.compress(host[i * chunk:(i + 1) * chunk], lvl).size for i in sample)
        mine = sum(int(sizes[i]) for i in sample)
        r = dict(chunk=chunk, n=n, level=lvl, compress_gbs=round(host.size / cms / 1e6, 2), decompress_gbs=round(host.size / dms / 1e6, 1),
                 ratio=round(host.size / float(sizes.sum()), 3), size_vs_libzstd=round(mine / zs, 4))
        print(json.dumps(r)); out.append(r)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/level_sweep.json', 'w'), indent=1)
