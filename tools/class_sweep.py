"""Per data class: decode GB/s of libzstd frames and compress GB/s, device-resident, CUDA events (best of 3).
Classes are SURVEY.md 8(d)'s generator settings; "mixed" is config 5's idx-mod-8 mix.  Looks for throughput cliffs
(chunks leaving the decode fast path, raw / RLE blocks, literal-only blocks, 128 KB chunks at L9).
usage: python tools/class_sweep.py [--chunks 8192]"""
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
ap.add_argument("--chunks", type=int, default=8192)
a = ap.parse_args()
pkg = ge.import_package()
orc, z = Oracle(), LibZstd()
CASES = [  # name, chunk, kind, P, frame level, checksum
    ("P=0 literals only", 65536, 0, 0, 3, False),
    ("P=0.25", 65536, 0, 16384, 3, False),
    ("P=0.50 (config 2)", 65536, 0, 32768, 3, False),
    ("P=0.75", 65536, 0, 49152, 3, False),
    ("P=0.90", 65536, 0, 58982, 3, False),
    ("uniform random", 65536, 1, 0, 3, False),
    ("zeros", 65536, 3, 0, 3, False),
    ("mixed (config 5)", 65536, 2, 0, 3, False),
    ("P=0.50 L1 frames", 65536, 0, 32768, 1, False),
    ("P=0.50 128K L9 + checksum (config 4)", 131072, 0, 32768, 9, True),
]


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


out = []
for name, chunk, kind, P, level, ck in CASES:
    n = a.chunks if chunk == 65536 else a.chunks // 2
    piece = 256

    def work(i, chunk=chunk, kind=kind, P=P, level=level, ck=ck):
        d = orc.gen_batch(chunk, piece, kind, P, first_idx=i * piece)
        blob, offs, sizes = z.compress_chunks(d, chunk, level, ck)
        return d, blob, sizes

    with ThreadPoolExecutor(8) as ex:
        parts = list(ex.map(work, range(n // piece)))
    host = np.concatenate([p[0] for p in parts])
    sizes = np.concatenate([p[2] for p in parts]).astype(np.uint64)
    blob = np.concatenate([p[1] for p in parts])
    offs = np.zeros(n, np.uint64); offs[1:] = np.cumsum(sizes)[:-1]
    dev = torch.from_numpy(host).cuda()
    comp = torch.from_numpy(blob).cuda()
    codec = pkg.ZstdBatchCodec(level=level, checksum=ck)
    wsd = torch.empty(codec.decompress_temp_size(n, sizes), dtype=torch.uint8, device="cuda")
    back, _ = codec.decompress_chunks(comp, offs, sizes, chunk, wsd)
    assert torch.equal(back, dev), name
    dms, _ = timed(lambda: codec.decompress_chunks(comp, offs, sizes, chunk, wsd))
    ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device="cuda")
    codec.compress_chunks(dev, chunk, ws)
    cms, (o, msz, stride) = timed(lambda: codec.compress_chunks(dev, chunk, ws))
    r = dict(case=name, chunk=chunk, n=n, level=level, decode_gbs=round(host.size / dms / 1e6, 1), compress_gbs=round(host.size / cms / 1e6, 2),
             libzstd_ratio=round(host.size / float(sizes.sum()), 3), ratio=round(host.size / float(msz.sum()), 3),
             size_vs_libzstd=round(float(msz.sum()) / float(sizes.sum()), 4))
    print(json.dumps(r), flush=True)
    out.append(r)
    del dev, comp, wsd, ws, back, o
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/class_sweep.json", "w"), indent=1)
