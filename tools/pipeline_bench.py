"""Host-to-host throughput of PipelinedBatchManager (include/pipeline_manager.hpp) over the C ABI: pageable host data in,
concatenated frames out, per batch size.  Mirrors the reference's benchmarks/benchmark_pipeline.cu:77-120 (its data set
is 256 MB of half-random bytes; here the bench's P=0.50 class, 1 GiB) -- wall clock around compress_stream_pipeline.
usage: python tools/pipeline_bench.py [--mib 1024] [--level 3]"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import __graft_entry__ as ge
from oracle.oracle import LibZstd, Oracle

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--level", type=int, default=3)
a = ap.parse_args()
pkg = ge.import_package()
orc, z = Oracle(), LibZstd()
n = a.mib * 16
with ThreadPoolExecutor(8) as ex:
    data = np.concatenate(list(ex.map(lambda i: orc.gen_batch(65536, 256, 0, 32768, first_idx=i * 256), range(n // 256))))
out = np.empty(data.size // 2, dtype=np.uint8)
res = []
for batch_mib, slots in ((16, 3), (16, 8), (64, 3), (64, 6), (128, 3), (128, 5)):
    pipe = pkg.ZstdPipeline(level=a.level, batch_bytes=batch_mib << 20, slots=slots)
    pipe.compress(data[: 2 * (batch_mib << 20)], out)                      # warm-up: kernels loaded, pinned buffers touched
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter(); blob, sizes = pipe.compress(data, out); best = min(best, time.perf_counter() - t0)
    pipe.close()
    # spot check: first and last frame through stock libzstd
    first = z.decompress(blob[: sizes[0]], batch_mib << 20)
    assert np.array_equal(first, data[: batch_mib << 20])
    last_n = data.size - (len(sizes) - 1) * (batch_mib << 20)
    last = z.decompress(blob[blob.size - sizes[-1]:], last_n)
    assert np.array_equal(last, data[data.size - last_n:])
    r = dict(batch_mib=batch_mib, slots=slots, frames=len(sizes), host_to_host_gbs=round(data.size / best / 1e9, 2), ratio=round(data.size / blob.size, 3), level=a.level)
    print(json.dumps(r), flush=True)
    res.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/pipeline_bench.json", "w"), indent=1)
