"""The host-side model of the GPU compressor (tests/model/enc_model.cpp) shares the entropy-stage code
with the kernel.  These CPU tests prove that what that algorithm emits is valid Zstandard (stock
libzstd and the oracle decode it bit-exactly) and meets the size bar against libzstd, per level."""
import numpy as np
import pytest

from helpers import CLASSES, edge_inputs

TOL = 1.02      # north_star: compressed size within 2 % of the reference's batch output (= libzstd at that level)


@pytest.mark.parametrize("level", [1, 3, 5, 7, 9, 12])
def test_model_edge_cases_decode(oracle, libzstd, model, level):
    for name, d in edge_inputs(oracle).items():
        for ck in (False, True):
            f = model.compress(d, level, ck)
            assert f.size > 0, name
            assert np.array_equal(libzstd.decompress(f, d.size), d), (name, level, ck)
            rc, out = oracle.decompress(f, d.size)
            assert rc == 0 and np.array_equal(out, d), (name, level, ck, rc)
            assert libzstd.frame_content_size(f) == d.size


@pytest.mark.parametrize("chunk,level", [(65536, 1), (65536, 3), (65536, 9), (131072, 9), (65536, 5), (65536, 7)])
def test_model_size_within_tolerance_of_libzstd(oracle, libzstd, model, chunk, level):
    for name, kind, P in CLASSES:
        d = oracle.gen_batch(chunk, 6, kind, P)
        mine = sum(model.compress(d[i * chunk:(i + 1) * chunk], level).size for i in range(6))
        _, _, sizes = libzstd.compress_chunks(d, chunk, level)
        assert mine <= TOL * float(sizes.sum()) + 6 * 8, (name, level, mine, int(sizes.sum()))


def test_model_multiblock(oracle, libzstd, model):
    d = oracle.gen_batch(400000, 1, 0, 30000)
    for level in (1, 3, 9):
        f = model.compress(d, level, True)
        assert np.array_equal(libzstd.decompress(f, d.size), d)
        rc, out, info = oracle.decompress(f, d.size, want_info=True)
        assert rc == 0 and np.array_equal(out, d) and info.n_blocks == 4 and info.has_checksum == 1


def test_model_uses_huffman_fse_and_repcodes(oracle, model):
    d = oracle.gen_batch(65536, 1, 0, 32768)
    rc, out, info = oracle.decompress(model.compress(d, 3), 65536, want_info=True)
    assert rc == 0 and info.lit_mode[2] == 1                      # Huffman-compressed literals
    assert info.seq_mode[0][2] == info.seq_mode[1][2] == info.seq_mode[2][2] == 1   # FSE-compressed tables
    t = oracle.gen_textlike(65536)
    f = model.compress(t, 3)
    assert f.size < 3600                                          # libzstd -3 needs 3288 on this chunk


@pytest.mark.parametrize("level", [1, 3, 5, 9])
def test_model_stitching_shortcut_equals_plain_rule(oracle, model, level, monkeypatch):
    """The lanes' walks are joined where position and the two youngest repeat offsets agree, the sequences up to
    the point where the third one agrees too being re-coded (zstd_encode_lz.cuh: select_rewalk / select_recode).
    The plain rule -- keep walking until all three agree -- must give the same frame, byte for byte."""
    inputs = {k: v for k, v in edge_inputs(oracle).items() if v.size >= 4096}
    for name, kind, P in CLASSES:
        inputs[name] = oracle.gen_batch(65536, 1, kind, P)
    rng = np.random.default_rng(11)
    # a few offsets in rotation: repeat codes of all three history entries, long stretches without a new offset
    base = rng.integers(0, 256, 700, dtype=np.uint8)
    rot = np.concatenate([np.roll(base, int(s))[: int(l)] for s, l in zip(rng.integers(0, 3, 400) * 231, rng.integers(20, 400, 400))])
    inputs["rotating"] = rot[:131072].copy()
    for name, d in inputs.items():
        monkeypatch.delenv("ENC_MODEL_EXACT_STITCH", raising=False)
        fast = model.compress(d, level)
        monkeypatch.setenv("ENC_MODEL_EXACT_STITCH", "1")
        plain = model.compress(d, level)
        assert np.array_equal(fast, plain), (name, level)
