"""Encoder residency sweep: resident one-warp CTAs per SM vs compress throughput (device-resident, CUDA events)."""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as ge
from oracle.oracle import Oracle
from concurrent.futures import ThreadPoolExecutor
pkg = ge.import_package(); lib = pkg.load_library(); orc = Oracle()
chunk, n = 65536, 16384
host = np.concatenate(list(ThreadPoolExecutor(8).map(lambda i: orc.gen_batch(chunk, 256, 0, 32768, first_idx=i * 256), range(n // 256))))
dev = torch.from_numpy(host).cuda()
for lvl in (1, 3, 5):
    for cap in (16, 20, 24, 28, 32):
        lib.cuda_zstd_b200_tune_enc_ctas(cap)
        codec = pkg.ZstdBatchCodec(level=lvl)
        ws = torch.empty(codec.compress_temp_size(n), dtype=torch.uint8, device='cuda')
        for _ in range(2): codec.compress_chunks(dev, chunk, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): o, sizes, stride = codec.compress_chunks(dev, chunk, ws)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps(dict(level=lvl, ctas_per_sm=cap, ms=round(ms, 2), gbs=round(host.size / ms / 1e6, 2), ws_mb=ws.numel() >> 20)))
