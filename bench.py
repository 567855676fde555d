#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native batched Zstandard codec on BASELINE.json's configurations.

Metric (BASELINE.json): batch ZSTD compress & decompress GB/s on 64 KiB chunks at level 3.  GB = 1e9 bytes of
UNCOMPRESSED data.  The headline `value` is the configuration the metric is quoted on that fits one GPU (configs[1],
"config 2"): batch DECOMPRESS of 16,384 x 64 KiB chunks per GPU whose frames were produced by libzstd level 3 (exactly
what the reference's default batch compress emits at this chunk size).  The same JSON line carries, under "configs",
one block per remaining BASELINE configuration, each with its own roofline, CPU baseline at the same level, and e2e:

  compress_l3   level-3 batch compress of the headline's chunks (the other half of the metric)
  config3       ZstdBatchManager batch compress, level 1, 1 GiB in 64 KiB chunks
  config4       level 9 + XXH64 checksums, 2 GiB in 128 KiB chunks: compress, and decompress of those frames
  config5       8 GiB mixed-entropy batch (131,072 x 64 KiB) sharded by chunk over the N ranks (STRONG scaling: total work
                fixed, N = 1 holds all of it): compress -> all-gather of the per-chunk sizes -> device scan -> decompress

Top-level fields follow the contract:
  value      whole-job decompress throughput, inputs (frames + pointer/size tables) resident in HBM, CUDA events around
             K calls on the launching stream, max over ranks (weak scaling: every rank decodes its own 16,384 chunks)
  e2e        the same step through the library's host-resident batch call (cuda_zstd_batch_decompress_host, the call a
             user with host data makes): frames and results in pinned HOST memory, H2D / D2H inside the timed region
  roofline   HBM roofline of the decode pipeline: algorithmic bytes (U + C per chunk) / measured time
  cpu_baseline  the reference's own CPU path (HybridEngine FORCE_CPU == libzstd) on this box's host cores

`--impl reference` times the reference's CPU implementation (oracle/_ref, else libzstd through the oracle bindings) with
all host threads on the SAME config (16,384 chunks per step) and prints the same JSON shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK = 65536
N_CHUNKS = 16384
LEVEL = 3
P_KNOB = 32768          # tunable-entropy class P = 0.50 (configs 2-4)
METRIC = "batch ZSTD decompress GB/s (16384 x 64 KiB chunks, libzstd L3 frames)"
C5_TOTAL = 131072       # config 5: 8 GiB of 64 KiB chunks over all ranks


def headline_config(n_chunks: int, world: int) -> dict:
    """The `config` object: identical in the repo arm and the reference arm (no measured values inside)."""
    return {"workload": "config2: NvcompV5BatchManager-style batch decompress of 16384 x 64 KiB chunks per GPU, frames by libzstd L3, tunable-entropy P=0.50",
            "chunk_bytes": CHUNK, "chunks_per_gpu": n_chunks, "level": LEVEL,
            "l2": "inputs+outputs per step (1.2 GB) exceed the 126 MB L2; no explicit flush", "parallelism": f"chunk-sharded x{world}"}


def host_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def gen_chunks(first_idx: int, n_chunks: int, chunk: int, kind: int, P: int, threads: int) -> np.ndarray:
    from oracle.oracle import Oracle
    orc = Oracle()
    piece = 256
    jobs = [(i, min(piece, n_chunks - i)) for i in range(0, n_chunks, piece)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda j: orc.gen_batch(chunk, j[1], kind, P, first_idx=first_idx + j[0]), jobs))
    return np.concatenate(parts)


def make_workload(first_idx: int, n_chunks: int, threads: int):
    """Synthetic tunable-entropy chunks + their libzstd L3 frames (the reference's batch output)."""
    from oracle.oracle import LibZstd, Oracle
    orc, z = Oracle(), LibZstd()
    piece = 256
    jobs = [(i, min(piece, n_chunks - i)) for i in range(0, n_chunks, piece)]

    def work(job):
        i, m = job
        d = orc.gen_batch(CHUNK, m, Oracle.KIND_TUNABLE, P_KNOB, first_idx=first_idx + i)
        blob, offs, sizes = z.compress_chunks(d, CHUNK, LEVEL)
        return d, blob, sizes

    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(work, jobs))
    data = np.concatenate([p[0] for p in parts])
    sizes = np.concatenate([p[2] for p in parts]).astype(np.uint64)
    # frames at 16-byte aligned offsets inside one blob
    aligned = (sizes + np.uint64(15)) // np.uint64(16) * np.uint64(16)
    offs = np.zeros(n_chunks, np.uint64)
    offs[1:] = np.cumsum(aligned)[:-1]
    blob = np.zeros(int(offs[-1] + aligned[-1]), np.uint8)
    k = 0
    for p in parts:
        pos = 0
        for s in p[2]:
            s = int(s)
            blob[int(offs[k]): int(offs[k]) + s] = p[1][pos: pos + s]
            pos += s
            k += 1
    return data, blob, offs, sizes


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kind: str):
    """dram bytes per launch from the committed ncu summary, if one exists (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kind)
        except Exception:
            return None
    return None


# ---------------------------------------------------------------------------------------------------------------------
# CPU side: the reference's own path (HybridEngine FORCE_CPU, oracle/_ref) or, without it, libzstd through the oracle
# ---------------------------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self, threads: int):
        from oracle.oracle import LibZstd, RefHybrid
        self.threads = threads
        self.ref = RefHybrid() if RefHybrid.available() else None
        self.z = LibZstd()
        self.kind = "reference" if self.ref else "port"

    def decompress_secs(self, comp: np.ndarray, stride: int, sizes: np.ndarray, chunk: int) -> float:
        if self.ref:
            return self.ref.decompress(comp, stride, sizes, chunk, self.threads)[0]
        m, T = len(sizes), self.threads
        parts = [range(m * t // T, m * (t + 1) // T) for t in range(T)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(lambda r: [self.z.decompress(comp[i * stride: i * stride + int(sizes[i])], chunk) for i in r], parts))
        return time.perf_counter() - t0

    def compress_secs(self, data: np.ndarray, chunk: int, level: int, checksum: bool = False):
        """returns (seconds, compressed bytes)"""
        if self.ref and not checksum:
            s, _, _, osz = self.ref.compress(data, chunk, level, self.threads)
            return s, int(osz.sum())
        m, T = data.size // chunk, self.threads
        parts = [range(m * t // T, m * (t + 1) // T) for t in range(T)]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(T) as ex:
            tot = sum(ex.map(lambda r: sum(self.z.compress(data[i * chunk:(i + 1) * chunk], level, checksum).size for i in r), parts))
        return time.perf_counter() - t0, int(tot)


def strided_frames(blob, offs, sizes, m):
    stride = (int(sizes[:m].max()) + 15) // 16 * 16
    comp = np.zeros(stride * m, np.uint8)
    for i in range(m):
        comp[i * stride: i * stride + int(sizes[i])] = blob[int(offs[i]): int(offs[i]) + int(sizes[i])]
    return comp, stride


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n = args.chunks
    data, blob, offs, sizes = make_workload(0, n, threads)
    comp, stride = strided_frames(blob, offs, sizes, n)
    cpu = CpuArm(threads)
    step = lambda: cpu.decompress_secs(comp, stride, sizes, CHUNK)       # noqa: E731
    for _ in range(args.warmup):
        step()
    secs = [step() for _ in range(args.steps)]
    tot = sum(secs)
    val = n * CHUNK * args.steps / tot / 1e9
    cs, cb = cpu.compress_secs(data, CHUNK, LEVEL)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": headline_config(n, args.gpus),
            "ratio": float(data.size) / float(sizes.sum()),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": cpu.kind,
                             "sample": f"the full config: {n} x 64 KiB chunks per step, {threads} host threads",
                             "compress_l3_gbs": data.size / cs / 1e9, "compress_l3_ratio": data.size / cb},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------------------------
class Timer:
    """CUDA-event timing of `steps` calls of fn on the current stream, max over ranks."""

    def __init__(self, torch, dist, world, dev):
        self.torch, self.dist, self.world, self.dev = torch, dist, world, dev

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        if self.world > 1:
            t = self.torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def time(self, fn, steps: int, warmup: int, stream):
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [self.torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for k in range(steps):
            fn()
            ev[k + 1].record(stream)
        self.barrier()
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        return self.max_over_ranks(ev[0].elapsed_time(ev[-1])), per

    def wall(self, fn, steps: int, warmup: int):
        """wall clock around calls that synchronise themselves (host-resident batch calls), max over ranks"""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        self.barrier()
        return self.max_over_ranks(ms)


def device_tables(torch, dev, base_in, in_offs, in_sizes, base_out, out_offs, caps):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int64)).to(dev)          # noqa: E731
    return (t(np.uint64(base_in) + in_offs.astype(np.uint64)), t(in_sizes), t(np.uint64(base_out) + out_offs.astype(np.uint64)), t(caps))


def compress_block(torch, pkg, T, dev, world, rank, data, chunk, level, checksum, steps, peak, cpu, cpu_sample_chunks, label, e2e=True,
                   decode_too=False):
    """Batch compress of `data` in `chunk`-byte pieces at `level`: device-resident timing, roofline, e2e through
    cuda_zstd_batch_compress_host_packed, CPU baseline at the same level; optionally the decode of the produced frames."""
    from oracle.oracle import LibZstd
    n = data.size // chunk
    U = int(data.size)
    codec = pkg.ZstdBatchCodec(level=level, checksum=checksum)
    stream = torch.cuda.current_stream()
    d_in = torch.from_numpy(data).to(dev)
    stride = (codec.max_compressed_size(chunk) + 15) // 16 * 16
    c_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    idx = np.arange(n, dtype=np.uint64)
    in_ptrs, in_sizes, out_ptrs, caps = device_tables(torch, dev, d_in.data_ptr(), idx * np.uint64(chunk), np.full(n, chunk), c_out.data_ptr(),
                                                      idx * np.uint64(stride), np.full(n, stride))
    c_sizes = caps.clone()
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    sizes_h = np.full(n, chunk, np.uint64)
    ws = torch.empty(max(codec.compress_temp_size(n, sizes_h), codec.decompress_temp_size(n)), dtype=torch.uint8, device=dev)
    plan = pkg.ShardPlan(n * world, rank, world)
    launches = [0]

    def step():
        c_sizes.copy_(caps)
        rc = codec.compress_nosync(in_ptrs, in_sizes, n, out_ptrs, c_sizes, status, ws, stream)
        assert rc == 0
        table = pkg.gather_sizes(c_sizes, plan)                      # the one cross-GPU step (no-op at N=1)
        launches[0] += codec.last_launch_count() + 1
        return codec.scan_sizes(table, 0, stream)                    # device-side exclusive scan -> packed offsets

    total_ms, per = T.time(step, steps, 2, stream)
    assert int(status.max().item()) == 0
    csz = int(c_sizes.sum().item())
    z = LibZstd()
    szh = c_sizes.cpu().numpy()
    for i in (0, n // 2, n - 1):                                      # spot-check: frames decode in stock libzstd
        f = c_out[i * stride: i * stride + int(szh[i])].cpu().numpy()
        assert np.array_equal(z.decompress(f, chunk), data[i * chunk:(i + 1) * chunk])
    ms = total_ms / steps
    gbs = world * U / (ms * 1e-3) / 1e9
    blk = {"workload": label, "value": gbs, "unit": "GB/s", "ms_per_step": ms, "steps": steps, "ratio": U / csz, "level": level,
           "chunk_bytes": chunk, "chunks_per_gpu": n, "checksum": bool(checksum),
           "roofline": {"bound": "hbm", "achieved": (U + csz) / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": ((U + csz) / (ms * 1e-3) / 1e9) / peak, "traffic": ncu_traffic(f"compress_l{level}"),
                        "algorithmic_bytes_per_launch": U + csz}}
    if decode_too:
        d_back = torch.empty(U, dtype=torch.uint8, device=dev)
        bp = torch.from_numpy((np.uint64(d_back.data_ptr()) + idx * np.uint64(chunk)).astype(np.int64)).to(dev)
        bcaps = torch.full((n,), chunk, dtype=torch.int64, device=dev)
        bsz = bcaps.clone()
        fsz = c_sizes.clone()

        def dstep():
            bsz.copy_(bcaps)
            assert codec.decompress_nosync(out_ptrs, fsz, n, bp, bsz, status, ws, stream) == 0
            launches[0] += codec.last_launch_count()

        dms, _ = T.time(dstep, steps, 2, stream)
        assert int(status.max().item()) == 0 and torch.equal(d_back, d_in), "round trip differs"
        dms /= steps
        blk["decompress"] = {"value": world * U / (dms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": dms,
                             "roofline_frac": ((U + csz) / (dms * 1e-3) / 1e9) / peak, "verified": "bit-exact vs input (checksums verified on the device)"}
        del d_back
    if e2e:
        h_in = torch.from_numpy(data).pin_memory()
        h_packed = torch.empty(csz + csz // 8 + (1 << 20), dtype=torch.uint8).pin_memory()
        h_offs = np.zeros(n + 1, np.uint64)
        hws = torch.empty(codec.host_compress_temp_size(sizes_h), dtype=torch.uint8, device=dev)
        hp = (np.uint64(h_in.data_ptr()) + idx * np.uint64(chunk)).astype(np.uint64)

        def hstep():
            rc = codec.compress_host_packed(hp, sizes_h, h_packed, h_packed.numel(), h_offs, hws, None, stream)
            assert rc == 0
            launches[0] += codec.last_launch_count()

        esteps = max(2, min(steps, 5))
        ems = T.wall(hstep, esteps, 1) / esteps
        assert int(h_offs[-1]) == csz
        f = h_packed.numpy()[int(h_offs[n - 1]): int(h_offs[n])]
        assert np.array_equal(z.decompress(f, chunk), data[(n - 1) * chunk:])
        blk["e2e"] = {"value": world * U / (ems * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ems, "steps": esteps,
                      "h2d_bytes_per_step": U + 4 * n * 8, "d2h_bytes_per_step": csz + (n + 1) * 8 + 4 * n,
                      "call": "cuda_zstd_batch_compress_host_packed: pinned host chunks in, packed frames + offsets out (library-side wave pipeline)"}
        del h_in, h_packed, hws
    if cpu is not None:
        m = min(cpu_sample_chunks, n)
        secs, cb = cpu.compress_secs(data[: m * chunk], chunk, level, checksum)
        blk["cpu_baseline"] = {"value": m * chunk / secs / 1e9, "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind if not checksum else "port",
                               "sample": f"{m} x {chunk // 1024} KiB chunks, level {level}, {cpu.threads} threads", "ratio": m * chunk / cb}
        blk["size_vs_cpu_same_level"] = (float(szh[:m].sum()) / cb)
    blk["gpu_launches"] = launches[0]
    del d_in, c_out, ws
    torch.cuda.empty_cache()
    return blk


def config5_block(torch, dist, pkg, T, dev, world, rank, steps, peak, threads):
    """8 GiB mixed-entropy batch sharded by chunk index over the ranks (strong scaling)."""
    plan = pkg.ShardPlan(C5_TOTAL, rank, world)
    n = plan.count
    data = gen_chunks(plan.lo, n, CHUNK, 2, 0, threads)
    U = int(data.size)
    codec = pkg.ZstdBatchCodec(level=LEVEL, checksum=False)
    stream = torch.cuda.current_stream()
    d_in = torch.from_numpy(data).to(dev)
    stride = (codec.max_compressed_size(CHUNK) + 15) // 16 * 16
    c_out = torch.empty(n * stride, dtype=torch.uint8, device=dev)
    d_back = torch.empty(U, dtype=torch.uint8, device=dev)
    idx = np.arange(n, dtype=np.uint64)
    in_ptrs, in_sizes, out_ptrs, caps = device_tables(torch, dev, d_in.data_ptr(), idx * np.uint64(CHUNK), np.full(n, CHUNK), c_out.data_ptr(),
                                                      idx * np.uint64(stride), np.full(n, stride))
    bp = torch.from_numpy((np.uint64(d_back.data_ptr()) + idx * np.uint64(CHUNK)).astype(np.int64)).to(dev)
    bcaps = torch.full((n,), CHUNK, dtype=torch.int64, device=dev)
    c_sizes, bsz = caps.clone(), bcaps.clone()
    status = torch.zeros(n, dtype=torch.int32, device=dev)
    sizes_h = np.full(n, CHUNK, np.uint64)
    ws = torch.empty(max(codec.compress_temp_size(n, sizes_h), codec.decompress_temp_size(n)), dtype=torch.uint8, device=dev)
    launches = [0]
    last = {}

    def cstep():
        c_sizes.copy_(caps)
        assert codec.compress_nosync(in_ptrs, in_sizes, n, out_ptrs, c_sizes, status, ws, stream) == 0
        table = pkg.gather_sizes(c_sizes, plan)                      # all ranks learn all sizes: the one exchange of the job
        last["offsets"] = codec.scan_sizes(table, 0, stream)         # global packed offsets, on the device
        launches[0] += codec.last_launch_count() + 1

    def dstep():
        bsz.copy_(bcaps)
        assert codec.decompress_nosync(out_ptrs, c_sizes, n, bp, bsz, status, ws, stream) == 0
        launches[0] += codec.last_launch_count()

    def both():
        cstep()
        dstep()

    cms, _ = T.time(cstep, steps, 3, stream)          # (the first calls of the size all-gather set NCCL's channels up)
    assert int(status.max().item()) == 0
    csz_local = int(c_sizes.sum().item())
    total_c = int(last["offsets"][-1].item())
    dms, _ = T.time(dstep, steps, 1, stream)
    assert int(status.max().item()) == 0 and torch.equal(d_back, d_in), "config 5 round trip differs"
    rms, _ = T.time(both, steps, 0, stream)
    cms, dms, rms = cms / steps, dms / steps, rms / steps
    tot_U = C5_TOTAL * CHUNK
    blk = {"workload": "config5: 8 GiB mixed-entropy batch (131072 x 64 KiB, classes by chunk index), level 3, sharded by chunk index over the ranks; "
                       "compress -> all-gather of sizes -> device scan -> decompress",
           "scaling": "strong", "chunks_total": C5_TOTAL, "chunks_this_rank": n, "n_gpus": world,
           "compress_gbs": tot_U / (cms * 1e-3) / 1e9, "decompress_gbs": tot_U / (dms * 1e-3) / 1e9,
           "roundtrip_gbs": tot_U / (rms * 1e-3) / 1e9, "compress_ms": cms, "decompress_ms": dms, "roundtrip_ms": rms,
           "ratio": tot_U / total_c, "global_offsets_total": total_c,
           "roofline_frac_compress": ((U + csz_local) / (cms * 1e-3) / 1e9) / peak, "roofline_frac_decompress": ((U + csz_local) / (dms * 1e-3) / 1e9) / peak,
           "verified": "every rank: decode(encode(x)) == x bit-exact over its whole shard", "gpu_launches": launches[0]}
    del d_in, c_out, d_back, ws
    torch.cuda.empty_cache()
    return blk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=N_CHUNKS, help="chunks per GPU (default: the BASELINE config)")
    ap.add_argument("--skip-configs", action="store_true", help="headline only (no compress / config 3-5 blocks)")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.import_package()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    T = Timer(torch, dist, world, dev)

    n = args.chunks
    threads = max(1, host_threads() // world)
    data, blob, offs, sizes = make_workload(rank * n, n, threads)
    U, Cb = int(data.size), int(sizes.sum())
    peak, peak_src = measured_peak_hbm()

    codec = pkg.ZstdBatchCodec(level=LEVEL, checksum=False)
    # ---- headline: resident inputs ----
    d_comp = torch.from_numpy(blob).to(dev)
    d_out = torch.empty(U, dtype=torch.uint8, device=dev)
    idx = np.arange(n, dtype=np.uint64)
    t_in_ptrs, t_in_sizes, t_out_ptrs, t_caps = device_tables(torch, dev, d_comp.data_ptr(), offs, sizes, d_out.data_ptr(), idx * np.uint64(CHUNK),
                                                              np.full(n, CHUNK))
    t_out_sizes = t_caps.clone()
    t_status = torch.zeros(n, dtype=torch.int32, device=dev)
    ws = torch.empty(codec.decompress_temp_size(n, sizes), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    launches = [0]

    def dec_step():
        t_out_sizes.copy_(t_caps)      # sizes are in/out: restore the capacities (tiny D2D, inside the timed region)
        rc = codec.decompress_nosync(t_in_ptrs, t_in_sizes, n, t_out_ptrs, t_out_sizes, t_status, ws, stream)
        assert rc == 0
        launches[0] += codec.last_launch_count()

    for _ in range(args.warmup):
        dec_step()
    torch.cuda.synchronize()
    assert int(t_status.max().item()) == 0 and bool((t_out_sizes == CHUNK).all().item()), "decode failed"
    ref_dev = torch.from_numpy(data).to(dev)
    assert torch.equal(d_out, ref_dev), "decoded bytes differ from the input"        # parity inside the bench run
    del ref_dev
    launches[0] = 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, step_ms = T.time(dec_step, args.steps, 0, stream)
    clocks = sampler.stop() if rank == 0 else None
    value = world * U * args.steps / (total_ms * 1e-3) / 1e9
    kern_ms = float(np.mean(step_ms))      # one decode pipeline per step (+ a 128 KiB D2D of capacities)
    achieved = (U + Cb) / (kern_ms * 1e-3) / 1e9

    # ---- e2e: the library's host-resident batch call on pinned host buffers (H2D of the frames, D2H of the results inside) ----
    h_comp = torch.from_numpy(blob).pin_memory()
    h_out = torch.empty(U, dtype=torch.uint8).pin_memory()
    hp_in = (np.uint64(h_comp.data_ptr()) + offs).astype(np.uint64)
    hp_out = (np.uint64(h_out.data_ptr()) + idx * np.uint64(CHUNK)).astype(np.uint64)
    caps_h = np.full(n, CHUNK, np.uint64)
    hws = torch.empty(codec.host_decompress_temp_size(sizes, caps_h), dtype=torch.uint8, device=dev)
    h_status = np.zeros(n, np.uint32)
    osz = caps_h.copy()

    def e2e_step():
        osz[:] = caps_h
        rc = codec.decompress_host(hp_in, sizes, hp_out, osz, hws, h_status, stream)
        assert rc == 0
        launches[0] += codec.last_launch_count()

    e2e_ms = T.wall(e2e_step, args.steps, 1) / args.steps
    e2e_val = world * U / (e2e_ms * 1e-3) / 1e9
    assert int(h_status.max()) == 0 and int(osz.min()) == CHUNK
    assert np.array_equal(h_out.numpy(), data), "e2e output differs from the input"
    h2d_bytes = int(blob.size) + 4 * n * 8
    d2h_bytes = U + n * 8 + n * 4
    # what the link itself gives: one plain pinned D2H copy of the same output buffer (the e2e step cannot beat this)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    l0.record()
    for _ in range(3):
        h_out.copy_(d_out, non_blocking=True)
    l1.record()
    torch.cuda.synchronize()
    link_gbs = 3 * U / (l0.elapsed_time(l1) * 1e-3) / 1e9
    del h_comp, h_out, hws, d_comp, d_out, ws
    torch.cuda.empty_cache()

    # ---- CPU baseline of the headline (rank 0, N = 1 only) ----
    cpu_arm = CpuArm(host_threads()) if (rank == 0 and world == 1 and not args.skip_cpu) else None
    cpu = None
    if cpu_arm:
        m = min(4096, n)
        comp_s, stride_s = strided_frames(blob, offs, sizes, m)
        best = min(cpu_arm.decompress_secs(comp_s, stride_s, sizes[:m], CHUNK) for _ in range(3))
        cpu = {"value": m * CHUNK / best / 1e9, "unit": "GB/s", "cores": cpu_arm.threads, "kind": cpu_arm.kind,
               "sample": f"{m} x 64 KiB chunks (P=0.50), libzstd L3 frames, best of 3, {cpu_arm.threads} threads"}

    # ---- the other BASELINE configurations ----
    configs = {}
    csteps = max(2, min(args.steps, 5))
    if not args.skip_configs:
        def guarded(name, fn):
            try:
                configs[name] = fn()
            except Exception as ex:          # a failing side block must not take the headline with it: report it
                configs[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                torch.cuda.empty_cache()
        guarded("compress_l3", lambda: compress_block(torch, pkg, T, dev, world, rank, data, CHUNK, 3, False, csteps, peak, cpu_arm, 4096,
                                                      "level-3 batch compress of the headline's 16384 x 64 KiB chunks per GPU (+ size all-gather and device scan)"))
        guarded("config3", lambda: compress_block(torch, pkg, T, dev, world, rank, data, CHUNK, 1, False, csteps, peak, cpu_arm, 4096,
                                                  "config3: ZstdBatchManager batch compress level 1 of 1 GiB in 64 KiB chunks per GPU"))
        if world == 1:
            def c4():
                d4 = gen_chunks(1 << 20, 16384, 131072, 0, P_KNOB, threads)
                return compress_block(torch, pkg, T, dev, world, rank, d4, 131072, 9, True, 2, peak, cpu_arm, 512,
                                      "config4: batch compress level 9 of 2 GiB in 128 KiB chunks with XXH64 checksums, and decompress of those frames",
                                      decode_too=True)
            guarded("config4", c4)
        guarded("config5", lambda: config5_block(torch, dist, pkg, T, dev, world, rank, 2, peak, threads))

    if rank == 0:
        total_launches = launches[0] + sum(int(b.get("gpu_launches", 0)) for b in configs.values() if isinstance(b, dict))
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": headline_config(n, world), "ratio": U / Cb,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("decode"), "peak_source": peak_src,
                         "kernel": "decode pipeline of one batch call: zstd_fast_prep_kernel, zstd_fast_order_kernel, zstd_fast_lit_kernel, 6 x (zstd_fast_seq_kernel || zstd_fast_exec_kernel)",
                         "algorithmic_bytes_per_launch": U + Cb, "kernel_ms": kern_ms},
            "e2e": {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "steps": args.steps, "ms_per_step": e2e_ms, "d2h_link_gbs": link_gbs,
                    "host_limit": f"the step moves {d2h_bytes / 1e9:.2f} GB device->host per rank; a plain pinned D2H copy of that buffer ran at "
                                  f"{link_gbs:.1f} GB/s on rank 0 right after the timed region ({world} rank(s) on this box share its PCIe / host memory)",
                    "call": "cuda_zstd_batch_decompress_host: frames + tables H2D and output + sizes + statuses D2H from / to pinned host memory, "
                            "staged in 4 waves on the library's copy streams beside the decode kernels"},
            "gpu_launches": total_launches, "clocks": clocks,
        }
        if configs:
            line["configs"] = configs
            if "compress_l3" in configs and "value" in configs["compress_l3"]:
                c = configs["compress_l3"]
                line["compress"] = {"l3_gbs": c["value"], "ms_per_step": c["ms_per_step"], "ratio": c["ratio"], "libzstd_l3_ratio": U / Cb,
                                    "size_vs_libzstd": (U / c["ratio"]) / Cb, "roofline_frac": c["roofline"]["frac"],
                                    "e2e_gbs": c.get("e2e", {}).get("value")}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
