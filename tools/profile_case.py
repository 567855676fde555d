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
ap.add_argument("--no-check", action="store_true", help="timing experiments with deliberately wrong kernels")
a = ap.parse_args()
pkg = ge.import_package()
data, blob, offs, sizes = make_workload(0, a.chunks, 8)
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
