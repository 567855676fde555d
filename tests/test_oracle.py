"""Pins the CPU oracle (oracle/zstd_oracle.c): known answers, golden frames produced by the
reference's own CPU path, and agreement with the system libzstd the reference links against."""
import numpy as np
import pytest

from helpers import CLASSES, golden_cases, golden_frame, golden_input


def test_xxh64_known_answers(oracle):
    # SURVEY.md 8c; the string is the one tests/test_checksum_validation.cu:62 of the reference hashes
    assert oracle.xxh64(b"The quick brown fox jumps over the lazy dog") == 0x0B242D361FDA71BC
    assert oracle.xxh64(b"") == 0xEF46DB3751D8E999
    import xxhash
    rng = np.random.default_rng(0)
    for n in (1, 3, 4, 7, 8, 31, 32, 33, 63, 64, 100, 4096, 65536):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert oracle.xxh64(d) == xxhash.xxh64(d.tobytes(), seed=0).intdigest()


def test_libzstd_known_answers(oracle, libzstd):
    assert libzstd.version() == 10505
    z = np.zeros(65536, np.uint8)
    assert libzstd.compress(z, 3).tobytes().hex() == "28b52ffd6000ff5500001000000100fb7f1d6001"
    assert libzstd.compress(oracle.gen_batch(65536, 1, oracle.KIND_RANDOM), 3).size == 65546
    assert libzstd.compress((np.arange(65536) % 256).astype(np.uint8), 1).size == 277
    assert libzstd.compress(oracle.gen_batch(65536, 1, 0, 32768), 3)[:7].tobytes().hex() == "28b52ffd6000ff"


def test_size_formula(oracle):
    # reference estimate_compressed_size, src/cuda_zstd_types.cpp:831-853 (probe values in SURVEY.md 8a)
    assert oracle.max_compressed_size(65536) == 66308
    assert oracle.max_compressed_size(131072) == 132101


def test_generator_matches_survey_table(oracle, libzstd):
    # mean libzstd size over 6 chunks of 64 KiB
    expect = {(0, 3): 43032, (16384, 3): 16390, (32768, 1): 12877, (32768, 3): 12157, (32768, 9): 9360, (49152, 3): 9932}
    for (P, level), want in expect.items():
        d = oracle.gen_batch(65536, 6, 0, P)
        _, _, sizes = libzstd.compress_chunks(d, 65536, level)
        assert abs(sizes.mean() - want) < 1.0, (P, level, sizes.mean())


def test_textlike_generator_is_the_references(oracle):
    # first bytes and the total libzstd size match the reference's generator (SURVEY.md section 6 probe: 3,350,936 B)
    t = oracle.gen_textlike(1 << 20)
    assert t[:36].tobytes() == b"hello world hello world hello world "


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_golden_frames(oracle, libzstd, case):
    data = golden_input(oracle, case)
    assert f"{oracle.xxh64(data):016x}" == case["input_xxh64"]
    frame = golden_frame(case)
    assert frame.size == case["frame_size"]
    rc, out, info = oracle.decompress(frame, data.size, want_info=True)
    assert rc == 0 and np.array_equal(out, data)
    assert info.content_size == data.size
    # the reference's CPU path is libzstd: same bytes today
    assert np.array_equal(frame, libzstd.compress(data, case["level"]))


@pytest.mark.parametrize("cls", CLASSES, ids=lambda c: c[0])
@pytest.mark.parametrize("level", [1, 3, 5, 9, 19])
@pytest.mark.parametrize("chunk", [65536, 131072])
def test_oracle_decodes_libzstd(oracle, libzstd, cls, level, chunk):
    _, kind, P = cls
    d = oracle.gen_batch(chunk, 3, kind, P)
    blob, off, sz = libzstd.compress_chunks(d, chunk, level, checksum=(level == 9))
    rc, out, osz = oracle.decompress_batch(blob, off, sz, chunk)
    assert rc == 0 and np.array_equal(out, d) and (osz == chunk).all()


def test_oracle_multiblock_and_repeat_modes(oracle, libzstd):
    # > 128 KB inputs produce multi-block frames that use Repeat_Mode / Treeless literals
    d = oracle.gen_batch(700000, 1, 0, 30000)
    for level in (1, 3, 7, 12):
        f = libzstd.compress(d, level)
        rc, out, info = oracle.decompress(f, d.size, want_info=True)
        assert rc == 0 and np.array_equal(out, d) and info.n_blocks > 1
    t = oracle.gen_textlike(1 << 20)
    f = libzstd.compress(t, 3)
    rc, out, info = oracle.decompress(f, t.size, want_info=True)
    assert rc == 0 and np.array_equal(out, t)
    assert info.seq_mode[0][3] + info.seq_mode[1][3] + info.seq_mode[2][3] + info.lit_mode[3] > 0   # repeat/treeless seen


def test_oracle_rejects_damage(oracle, libzstd):
    d = oracle.gen_batch(65536, 1, 0, 32768)
    f = libzstd.compress(d, 3, checksum=True)
    rc, _ = oracle.decompress(f[:-1], 65536)
    assert rc != 0
    bad = f.copy(); bad[0] ^= 1
    assert oracle.decompress(bad, 65536)[0] == 5                 # ERROR_INVALID_MAGIC
    bad = f.copy(); bad[-1] ^= 0x55
    assert oracle.decompress(bad, 65536)[0] == 10                # ERROR_CHECKSUM_FAILED
    assert oracle.decompress(f, 65535)[0] == 7                   # ERROR_BUFFER_TOO_SMALL
    rng = np.random.default_rng(3)
    for _ in range(200):                                         # random damage never crashes and is never silently accepted
        bad = f.copy()
        i = int(rng.integers(7, f.size - 4))
        bad[i] ^= int(rng.integers(1, 256))
        rc, out = oracle.decompress(bad, 65536)
        assert rc != 0 or np.array_equal(out, d)


def test_store_frame(oracle, libzstd):
    for n in (1, 255, 256, 65535, 65536, 65537, 131072, 300000):
        d = oracle.gen_batch(n, 1, 1, 0)
        for ck in (False, True):
            f = oracle.store_frame(d, ck)
            assert np.array_equal(libzstd.decompress(f, n), d)
            assert libzstd.frame_content_size(f) == n


def test_reference_cpu_path_is_libzstd(oracle, libzstd):
    from oracle.oracle import RefHybrid
    if not RefHybrid.available():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = RefHybrid()
    d = oracle.gen_batch(65536, 4, 2, 0)
    _, out, stride, sizes = ref.compress(d, 65536, 3, 2)
    for i in range(4):
        frame = out[i * stride: i * stride + int(sizes[i])]
        assert np.array_equal(frame, libzstd.compress(d[i * 65536:(i + 1) * 65536], 3))
        rc, back = oracle.decompress(frame, 65536)
        assert rc == 0 and np.array_equal(back, d[i * 65536:(i + 1) * 65536])
    _, dec, osz = ref.decompress(out, stride, sizes, 65536, 2)
    assert np.array_equal(dec, d)
