"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF, run in the authoring
container: oracle/_ref/libref_hybrid.so is the unmodified reference (built by oracle/build_ref.sh from
/root/reference) and HybridEngine{FORCE_CPU} is its own implementation of the batch path at BASELINE
chunk sizes (src/cuda_zstd_hybrid.cu:402-458, 779, 838).  /root/reference does not exist on the GPU
box, so the frames are committed; tests/test_oracle.py and tests/test_gpu_decode.py replay them.

Each fixture: <name>.zst (reference output frame) and an entry in manifest.json with the generator
call that reproduces the input (oracle.gen_batch / gen_textlike arguments), its XXH64 and sizes.
Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import LibZstd, Oracle, RefHybrid  # noqa: E402


def main():
    orc, ref, z = Oracle(), RefHybrid(), LibZstd()
    cases = []
    K = Oracle
    spec = [
        ("tunable_p50_64k_l3", dict(kind=K.KIND_TUNABLE, P=32768, chunk=65536, idx=0), 3),
        ("tunable_p50_64k_l1", dict(kind=K.KIND_TUNABLE, P=32768, chunk=65536, idx=1), 1),
        ("tunable_p25_64k_l9", dict(kind=K.KIND_TUNABLE, P=16384, chunk=65536, idx=2), 9),
        ("tunable_p75_128k_l9", dict(kind=K.KIND_TUNABLE, P=49152, chunk=131072, idx=3), 9),
        ("tunable_p0_64k_l3", dict(kind=K.KIND_TUNABLE, P=0, chunk=65536, idx=4), 3),
        ("tunable_p90_4k_l3", dict(kind=K.KIND_TUNABLE, P=58982, chunk=4096, idx=5), 3),
        ("random_64k_l3", dict(kind=K.KIND_RANDOM, P=0, chunk=65536, idx=6), 3),
        ("zeros_64k_l3", dict(kind=K.KIND_ZEROS, P=0, chunk=65536, idx=7), 3),
        ("tunable_p50_300_l3", dict(kind=K.KIND_TUNABLE, P=32768, chunk=300, idx=8), 3),
        ("tunable_p50_40_l3", dict(kind=K.KIND_TUNABLE, P=32768, chunk=40, idx=9), 3),
    ]
    for name, g, level in spec:
        data = orc.gen_batch(g["chunk"], 1, g["kind"], g["P"], first_idx=g["idx"])
        secs, out, stride, sizes = ref.compress(data, g["chunk"], level, 1)
        frame = out[: int(sizes[0])].copy()
        # the reference's CPU path IS libzstd: its bytes must equal a direct ZSTD_compress
        assert np.array_equal(frame, z.compress(data, level)), name
        frame.tofile(os.path.join(HERE, name + ".zst"))
        cases.append(dict(name=name, gen=g, level=level, input_xxh64=f"{orc.xxh64(data):016x}", input_size=int(data.size),
                          frame_size=int(frame.size)))
    # config 1 shape: first two 64 KiB chunks of the reference's own text-like generator
    text = orc.gen_textlike(2 * 65536)
    secs, out, stride, sizes = ref.compress(text, 65536, 3, 1)
    for i in range(2):
        frame = out[i * stride: i * stride + int(sizes[i])].copy()
        name = f"textlike_64k_l3_chunk{i}"
        frame.tofile(os.path.join(HERE, name + ".zst"))
        cases.append(dict(name=name, gen=dict(kind="textlike", total=2 * 65536, chunk=65536, idx=i), level=3,
                          input_xxh64=f"{orc.xxh64(text[i * 65536:(i + 1) * 65536]):016x}", input_size=65536,
                          frame_size=int(frame.size)))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(dict(generator="tests/golden/make_golden.py", reference="HybridEngine FORCE_CPU (oracle/_ref), libzstd 1.5.5",
                       cases=cases), f, indent=1)
    print(f"wrote {len(cases)} fixtures, {sum(c['frame_size'] for c in cases)} bytes")


if __name__ == "__main__":
    main()
