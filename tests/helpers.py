"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    with open(os.path.join(GOLDEN, "manifest.json")) as f:
        return json.load(f)["cases"]


def golden_input(orc, case):
    g = case["gen"]
    if g["kind"] == "textlike":
        t = orc.gen_textlike(g["total"])
        return t[g["idx"] * g["chunk"]:(g["idx"] + 1) * g["chunk"]].copy()
    return orc.gen_batch(g["chunk"], 1, g["kind"], g["P"], first_idx=g["idx"])


def golden_frame(case):
    return np.fromfile(os.path.join(GOLDEN, case["name"] + ".zst"), dtype=np.uint8)


CLASSES = [("p0", 0, 0), ("p25", 0, 16384), ("p50", 0, 32768), ("p75", 0, 49152), ("p90", 0, 58982), ("random", 1, 0),
           ("zeros", 3, 0), ("mixed", 2, 0)]


def edge_inputs(orc, sizes=(1, 2, 3, 7, 8, 9, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 4096, 5000, 65535, 65536, 65537, 131072)):
    rng = np.random.default_rng(7)
    out = {}
    for n in sizes:
        out[f"text{n}"] = orc.gen_textlike(n)
        out[f"gen{n}"] = orc.gen_batch(n, 1, 0, 40000)
        out[f"mod{n}"] = (np.arange(n) % 256).astype(np.uint8)
        out[f"rnd{n}"] = rng.integers(0, 256, n, dtype=np.uint8)
        out[f"zero{n}"] = np.zeros(n, np.uint8)
        out[f"two{n}"] = rng.integers(0, 2, n, dtype=np.uint8)
        out[f"per7_{n}"] = (np.arange(n) % 7).astype(np.uint8)
        if n in (5000, 65536, 131072):
            # the pattern of the reference's single-buffer benchmarks (benchmarks/benchmark_nvcomp_interface.cu:40-42):
            # two offsets alternate for the whole buffer, so the third repeat-offset entry never leaves the history
            i = np.arange(n, dtype=np.int64)
            out[f"drift{n}"] = ((i * 17 + i // 256) % 256).astype(np.uint8)
    return out
