"""Encoder timing by level on the bench's P=0.50 class: device-resident, CUDA events, best of 3; checks every frame with
libzstd on a sample and reports size vs libzstd at the same level.
usage: python tools/enc_time.py [--chunks 16384] [--levels 1,3] [--chunk 65536]"""
import argparse
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as ge
from oracle.oracle import LibZstd, Oracle

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=16384)
ap.add_argument("--chunk", type=int, default=65536)
ap.add_argument("--levels", default="1,3")
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--P", type=int, default=32768)
ap.add_argument("--checksum", action="store_true")
a = ap.parse_args()
pkg = ge.import_package()
orc, z = Oracle(), LibZstd()
n, chunk = a.chunks, a.chunk
with ThreadPoolExecutor(8) as ex:
    parts = list(ex.map(lambda i: orc.gen_batch(chunk, 256, a.kind, a.P, first_idx=i * 256), range((n + 255) // 256)))
host = np.concatenate(parts)[: n * chunk]
dev = torch.from_numpy(host).cuda()
res = []
for level in [int(x) for x in a.levels.split(",")]:
    codec = pkg.ZstdBatchCodec(level=level, checksum=a.checksum)
    ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device="cuda")
    out, sizes, stride = codec.compress_chunks(dev, chunk, ws)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out, sizes, stride = codec.compress_chunks(dev, chunk, ws); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    oh = out.cpu().numpy()
    sample = list(range(0, n, max(1, n // 64)))
    zs = 0
    for i in sample:
        f = oh[i * stride: i * stride + int(sizes[i])]
        assert np.array_equal(z.decompress(f, chunk), host[i * chunk:(i + 1) * chunk]), (level, i)
        zs += z.compress(host[i * chunk:(i + 1) * chunk], level, a.checksum).size
    r = dict(level=level, chunks=n, chunk=chunk, ms=round(best, 3), gbs=round(host.size / best / 1e6, 2), ratio=round(host.size / float(sizes.sum()), 3),
             size_vs_libzstd=round(float(sum(int(sizes[i]) for i in sample)) / zs, 4), ws_mb=ws.numel() >> 20)
    print(json.dumps(r), flush=True)
    res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/enc_time.json", "w"), indent=1)
