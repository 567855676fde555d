"""PipelinedBatchManager (reference src/pipeline_manager.hpp:35-66): host data streamed through
H2D -> batch compress -> D2H.  Same cases as the reference's tests/test_pipeline_integration.cu:118-133 (smaller than
a batch, exactly one batch, several aligned batches, an odd tail); where that program shells out to `zstd -t`, this
test decodes every frame with stock libzstd and with the plain-C oracle and compares with the input bit for bit."""
import numpy as np
import pytest

MB = 1 << 20


def _frames(blob, sizes):
    off = 0
    for n in sizes:
        yield blob[off:off + n]
        off += n
    assert off == blob.size


@pytest.mark.gpu
@pytest.mark.parametrize("total,batch,level,checksum", [
    (4 * MB, 16 * MB, 1, False),              # "Small Data"
    (16 * MB, 16 * MB, 1, False),             # "Exact Batch"
    (32 * MB, 16 * MB, 3, False),             # "Multi-Batch Aligned"
    (50 * MB + 12345, 16 * MB, 3, True),      # "Multi-Batch Odd" (+ ragged tail, + whole-frame checksums)
    (3 * MB + 100_000, 1 * MB, 5, False),     # small batches; the last one is below one 128 KB block
])
def test_pipeline_frames_decode_in_libzstd_and_oracle(pkg, oracle, libzstd, total, batch, level, checksum):
    n_chunks = (total + 65535) // 65536
    data = oracle.gen_batch(65536, n_chunks, kind=0, P=32768).reshape(-1)[:total].copy()
    pipe = pkg.ZstdPipeline(level=level, checksum=checksum, batch_bytes=batch, slots=3)
    blob, sizes = pipe.compress(data)
    pipe.close()
    assert len(sizes) == (total + batch - 1) // batch           # one frame per batch, in order
    assert blob.size < total // 3                                # this class compresses > 5x; a stored frame would not
    pos = 0
    for k, fr in enumerate(_frames(blob, sizes)):
        want = data[pos:pos + min(batch, total - pos)]
        assert libzstd.frame_content_size(fr) == want.size
        got = libzstd.decompress(fr, want.size)
        assert np.array_equal(got, want), f"libzstd: frame {k} differs"
        if k == 0 or k == len(sizes) - 1:                        # the oracle is slower; first and last frame are enough
            rc, got2 = oracle.decompress(fr, want.size, verify_checksum=True)
            assert rc == 0 and np.array_equal(got2, want), f"oracle: frame {k} differs (rc {rc})"
        if checksum:
            assert fr[4] & 0x04                                  # Content_Checksum_flag (RFC 8878 3.1.1.1.1)
        pos += want.size
    assert pos == total


@pytest.mark.gpu
def test_pipeline_incompressible_and_reuse(pkg, oracle, libzstd):
    """The reference test feeds uniform random bytes (tests/test_pipeline_integration.cu:25-30): frames of raw blocks.
    The same manager object is then used for a second stream."""
    pipe = pkg.ZstdPipeline(level=1, batch_bytes=2 * MB, slots=2)
    rnd = oracle.gen_batch(65536, 80, kind=1).reshape(-1)        # 5 MiB uniform random
    blob, sizes = pipe.compress(rnd)
    assert len(sizes) == 3 and rnd.size < blob.size < rnd.size + 3 * 1200
    pos = 0
    for fr in _frames(blob, sizes):
        n = min(2 * MB, rnd.size - pos)
        assert np.array_equal(libzstd.decompress(fr, n), rnd[pos:pos + n])
        pos += n
    zeros = np.zeros(3 * MB + 17, dtype=np.uint8)
    blob, sizes = pipe.compress(zeros)
    assert len(sizes) == 2 and blob.size < 4096
    pos = 0
    for fr in _frames(blob, sizes):
        n = min(2 * MB, zeros.size - pos)
        assert np.array_equal(libzstd.decompress(fr, n), zeros[pos:pos + n])
        pos += n
    pipe.close()


@pytest.mark.gpu
def test_pipeline_empty_input_and_bad_arguments(pkg):
    pipe = pkg.ZstdPipeline(level=3, batch_bytes=1 * MB)
    blob, sizes = pipe.compress(np.zeros(0, dtype=np.uint8))
    assert blob.size == 0 and sizes == []
    lib = pkg.load_library()
    assert lib.cuda_zstd_pipeline_create(3, 0, 0, 3) is None      # zero batch size
    assert lib.cuda_zstd_pipeline_create(3, 0, 1 << 20, 1) is None  # a ring needs two slots
    pipe.close()
