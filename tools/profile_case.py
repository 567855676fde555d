"""Small fixed workload for ncu captures: one kernel family per invocation.
usage: python tools/profile_case.py {decode|encode} [--chunks N] [--level L] [--iters K]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402
from bench import CHUNK, make_workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("mode", choices=["decode", "encode"])
ap.add_argument("--chunks", type=int, default=2368)
ap.add_argument("--level", type=int, default=3)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--kind", type=int, default=0, help="generator: 0 tunable(P), 1 random, 2 mixed (config 5), 3 zeros")
ap.add_argument("--P", type=int, default=None, help="match probability knob 0..65536 (default: the bench's 32768)")
ap.add_argument("--no-check", action="store_true", help="timing experiments with deliberately wrong kernels")
a = ap.parse_args()
pkg = ge.import_package()
if a.kind == 0 and a.P is None:
    data, blob, offs, sizes = make_workload(0, a.chunks, 8)
else:
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import LibZstd, Oracle
    orc, z = Oracle(), LibZstd()

    def work(i):
        d = orc.gen_batch(CHUNK, 256, a.kind, a.P if a.P is not None else 32768, first_idx=i * 256)
        b, _, sz = z.compress_chunks(d, CHUNK, a.level)
        return d, b, sz

    with ThreadPoolExecutor(8) as ex:
        parts = list(ex.map(work, range((a.chunks + 255) // 256)))
    data = np.concatenate([p[0] for p in parts])[: a.chunks * CHUNK]
    sizes = np.concatenate([p[2] for p in parts]).astype(np.uint64)[: a.chunks]
    blob = np.concatenate([p[1] for p in parts])
    offs = np.zeros(a.chunks, np.uint64); offs[1:] = np.cumsum(sizes)[:-1]
    blob = blob[: int(offs[-1] + sizes[-1])]
codec = pkg.ZstdBatchCodec(level=a.level)
n = a.chunks
if a.mode == "decode":
    comp = torch.from_numpy(blob).cuda()
    for _ in range(a.iters):
        try:
            out, osz = codec.decompress_chunks(comp, offs, sizes, CHUNK)
        except RuntimeError:
            if not a.no_check:
                raise
            out = None
    assert a.no_check or np.array_equal(out.cpu().numpy(), data)
else:
    dev = torch.from_numpy(data).cuda()
    for _ in range(a.iters):
        out, osz, stride = codec.compress_chunks(dev, CHUNK)
    print("ratio", data.size / float(osz.sum()))
torch.cuda.synchronize()
print("ok")
