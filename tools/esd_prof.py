"""Cycle accounting of the levels 1-4 encoder by warp role (instrumented build: make -C .../csrc prof).
usage: CUDA_ZSTD_B200_LIB=custom-nvcomp-with-zstd_b200/libcuda_zstd_b200_prof.so python tools/esd_prof.py [--level 3] [--chunks 16384]"""
import argparse
import ctypes as C
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as ge
from oracle.oracle import Oracle

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=16384)
ap.add_argument("--level", type=int, default=3)
ap.add_argument("--P", type=int, default=32768)
a = ap.parse_args()
pkg = ge.import_package()
lib = pkg.load_library()
orc = Oracle()
n, chunk = a.chunks, 65536
with ThreadPoolExecutor(8) as ex:
    parts = list(ex.map(lambda i: orc.gen_batch(chunk, 256, 0, a.P, first_idx=i * 256), range((n + 255) // 256)))
dev = torch.from_numpy(np.concatenate(parts)[: n * chunk]).cuda()
codec = pkg.ZstdBatchCodec(level=a.level)
ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device="cuda")
codec.compress_chunks(dev, chunk, ws)
buf = (C.c_ulonglong * 16)()
lib.cuda_zstd_b200_esd_prof(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); codec.compress_chunks(dev, chunk, ws); e1.record(); torch.cuda.synchronize()
lib.cuda_zstd_b200_esd_prof(buf, 0)
v = list(buf)
names = ["S busy", "S waits V", "block total", "finish busy", "H total", "H waits ring", "V total", "V waits H", "blocks", "load wait", "S steps",
         "sequences", "H exact-path windows", "open extensions"]
blocks = max(v[8], 1)
print(f"ms {e0.elapsed_time(e1):.2f}  blocks {v[8]}")
for i, nm in enumerate(names):
    print(f"{nm:16s} {v[i] / blocks:14.1f} per block")
json.dump({nm: v[i] / blocks for i, nm in enumerate(names)}, open("gpurun_out/esd_prof.json", "w"), indent=1)
