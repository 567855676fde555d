"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the checkers.

Three checkers live here; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (never the product package):

* ``Oracle``    -- oracle/liboracle_zstd.so, the plain-C restatement (oracle/zstd_oracle.c).
* ``LibZstd``   -- the system libzstd.so.1 (1.5.5).  This IS the arithmetic the reference runs at
                   BASELINE chunk sizes (reference: src/cuda_zstd_manager.cu:1604-1668, 3277-3341),
                   so it pins the oracle and defines the compressed-size target.
* ``RefHybrid`` -- oracle/_ref/libref_hybrid.so: the unmodified reference built from
                   /root/reference (oracle/build_ref.sh), driven through HybridEngine{FORCE_CPU}.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


class FrameInfo(C.Structure):
    _fields_ = [
        ("content_size", C.c_uint64), ("window_size", C.c_uint64),
        ("has_checksum", C.c_uint32), ("single_segment", C.c_uint32), ("dict_id", C.c_uint32),
        ("header_size", C.c_uint32),
        ("n_blocks", C.c_uint32), ("n_raw", C.c_uint32), ("n_rle", C.c_uint32), ("n_comp", C.c_uint32),
        ("lit_mode", C.c_uint32 * 4), ("seq_mode", (C.c_uint32 * 4) * 3),
        ("n_seq", C.c_uint64), ("n_lit", C.c_uint64),
        ("seq_log", (C.c_uint32 * 10) * 3),
    ]


def build_oracle() -> str:
    so = os.path.join(HERE, "liboracle_zstd.so")
    src = os.path.join(HERE, "zstd_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s", "liboracle_zstd.so"])
    return so


class Oracle:
    KIND_TUNABLE, KIND_RANDOM, KIND_MIXED, KIND_ZEROS = 0, 1, 2, 3

    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [_u8p, C.c_size_t, C.c_uint64]
        L.orc_max_compressed_size.restype = C.c_size_t
        L.orc_max_compressed_size.argtypes = [C.c_size_t]
        L.orc_gen_batch.argtypes = [_u8p, C.c_size_t, C.c_uint64, C.c_size_t, C.c_int, C.c_uint32]
        L.orc_gen_textlike.argtypes = [_u8p, C.c_size_t]
        L.orc_decompress.restype = C.c_int
        L.orc_decompress.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int,
                                     C.POINTER(FrameInfo)]
        L.orc_decompress_batch.restype = C.c_int
        L.orc_decompress_batch.argtypes = [_u8p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_size_t, _u8p,
                                           C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
        L.orc_store_frame.restype = C.c_size_t
        L.orc_store_frame.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_int]

    def xxh64(self, data: bytes | np.ndarray, seed: int = 0) -> int:
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        a = np.ascontiguousarray(a)
        return int(self.lib.orc_xxh64(_ptr(a) if a.size else None, a.size, seed))

    def max_compressed_size(self, n: int) -> int:
        return int(self.lib.orc_max_compressed_size(n))

    def gen_batch(self, chunk: int, n: int, kind: int = 0, P: int = 32768, first_idx: int = 0) -> np.ndarray:
        out = np.empty(chunk * n, dtype=np.uint8)
        self.lib.orc_gen_batch(_ptr(out), chunk, first_idx, n, kind, P)
        return out

    def gen_textlike(self, size: int) -> np.ndarray:
        out = np.empty(size, dtype=np.uint8)
        self.lib.orc_gen_textlike(_ptr(out), size)
        return out

    def decompress(self, frame, capacity: int, verify_checksum: bool = True, want_info: bool = False):
        src = np.ascontiguousarray(np.frombuffer(bytes(frame), dtype=np.uint8) if not isinstance(frame, np.ndarray) else frame)
        dst = np.empty(max(capacity, 1), dtype=np.uint8)
        n = C.c_size_t(0)
        info = FrameInfo()
        rc = self.lib.orc_decompress(_ptr(src), src.size, _ptr(dst), capacity, C.byref(n), int(verify_checksum), C.byref(info))
        out = dst[: n.value].copy()
        return (rc, out, info) if want_info else (rc, out)

    def decompress_batch(self, blob: np.ndarray, offsets: np.ndarray, sizes: np.ndarray, stride: int, verify: bool = True):
        n = len(sizes)
        off = np.ascontiguousarray(offsets, dtype=np.uint64)
        sz = np.ascontiguousarray(sizes, dtype=np.uint64)
        dst = np.empty(stride * n, dtype=np.uint8)
        out_sizes = np.zeros(n, dtype=np.uint64)
        rc = self.lib.orc_decompress_batch(_ptr(blob), off.ctypes.data_as(C.POINTER(C.c_size_t)),
                                           sz.ctypes.data_as(C.POINTER(C.c_size_t)), n, _ptr(dst), stride,
                                           out_sizes.ctypes.data_as(C.POINTER(C.c_size_t)), int(verify))
        return rc, dst, out_sizes

    def store_frame(self, data: np.ndarray, checksum: bool = False) -> np.ndarray:
        cap = data.size + 64 + 3 * (data.size // (128 * 1024) + 1)
        dst = np.empty(cap, dtype=np.uint8)
        n = self.lib.orc_store_frame(_ptr(data) if data.size else None, data.size, _ptr(dst), cap, int(checksum))
        return dst[:n].copy()


class LibZstd:
    """System libzstd 1.5.5 through its stable public ABI (no header needed)."""
    ZSTD_c_compressionLevel = 100
    ZSTD_c_checksumFlag = 201

    def __init__(self):
        self.lib = C.CDLL("libzstd.so.1")
        L = self.lib
        L.ZSTD_compress.restype = C.c_size_t
        L.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ZSTD_decompress.restype = C.c_size_t
        L.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ZSTD_compressBound.restype = C.c_size_t
        L.ZSTD_compressBound.argtypes = [C.c_size_t]
        L.ZSTD_isError.restype = C.c_uint
        L.ZSTD_isError.argtypes = [C.c_size_t]
        L.ZSTD_getErrorName.restype = C.c_char_p
        L.ZSTD_getErrorName.argtypes = [C.c_size_t]
        L.ZSTD_getFrameContentSize.restype = C.c_ulonglong
        L.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
        L.ZSTD_versionNumber.restype = C.c_uint
        L.ZSTD_createCCtx.restype = C.c_void_p
        L.ZSTD_freeCCtx.argtypes = [C.c_void_p]
        L.ZSTD_CCtx_setParameter.restype = C.c_size_t
        L.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ZSTD_compress2.restype = C.c_size_t
        L.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]

    def version(self) -> int:
        return int(self.lib.ZSTD_versionNumber())

    def is_error(self, code: int) -> bool:
        return bool(self.lib.ZSTD_isError(code))

    def compress(self, data: np.ndarray, level: int = 3, checksum: bool = False) -> np.ndarray:
        data = np.ascontiguousarray(data)
        cap = int(self.lib.ZSTD_compressBound(data.size))
        dst = np.empty(cap, dtype=np.uint8)
        if checksum:
            cctx = self.lib.ZSTD_createCCtx()
            self.lib.ZSTD_CCtx_setParameter(cctx, self.ZSTD_c_compressionLevel, level)
            self.lib.ZSTD_CCtx_setParameter(cctx, self.ZSTD_c_checksumFlag, 1)
            n = self.lib.ZSTD_compress2(cctx, dst.ctypes.data, cap, data.ctypes.data, data.size)
            self.lib.ZSTD_freeCCtx(cctx)
        else:
            n = self.lib.ZSTD_compress(dst.ctypes.data, cap, data.ctypes.data, data.size, level)
        if self.is_error(n):
            raise RuntimeError(self.lib.ZSTD_getErrorName(n).decode())
        return dst[:n].copy()

    def decompress(self, frame: np.ndarray, capacity: int) -> np.ndarray:
        """Raises RuntimeError with libzstd's message when the frame does not decode."""
        frame = np.ascontiguousarray(frame)
        dst = np.empty(max(capacity, 1), dtype=np.uint8)
        n = self.lib.ZSTD_decompress(dst.ctypes.data, capacity, frame.ctypes.data, frame.size)
        if self.is_error(n):
            raise RuntimeError(self.lib.ZSTD_getErrorName(n).decode())
        return dst[:n].copy()

    def frame_content_size(self, frame: np.ndarray) -> int:
        frame = np.ascontiguousarray(frame)
        return int(self.lib.ZSTD_getFrameContentSize(frame.ctypes.data, frame.size))

    def compress_chunks(self, data: np.ndarray, chunk: int, level: int, checksum: bool = False):
        """Per-chunk ZSTD_compress == what the reference's default batch path emits.  Returns
        (blob, offsets, sizes) with frames packed back to back."""
        n = (data.size + chunk - 1) // chunk
        frames = [self.compress(data[i * chunk:(i + 1) * chunk], level, checksum) for i in range(n)]
        sizes = np.array([f.size for f in frames], dtype=np.uint64)
        offsets = np.zeros(n, dtype=np.uint64)
        if n > 1:
            offsets[1:] = np.cumsum(sizes)[:-1]
        return np.concatenate(frames) if frames else np.empty(0, np.uint8), offsets, sizes


class RefHybrid:
    """The reference's own CPU implementation (HybridEngine FORCE_CPU), built into oracle/_ref."""

    @staticmethod
    def path() -> str:
        return os.path.join(HERE, "_ref", "libref_hybrid.so")

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(cls.path())

    def __init__(self):
        self.lib = C.CDLL(self.path())
        L = self.lib
        sz = C.POINTER(C.c_size_t)
        L.ref_hybrid_cpu_compress.restype = C.c_double
        L.ref_hybrid_cpu_compress.argtypes = [_u8p, sz, C.c_size_t, C.c_size_t, _u8p, C.c_size_t, sz, C.c_int, C.c_int]
        L.ref_hybrid_cpu_decompress.restype = C.c_double
        L.ref_hybrid_cpu_decompress.argtypes = [_u8p, sz, C.c_size_t, C.c_size_t, _u8p, C.c_size_t, sz, C.c_int]
        L.ref_hybrid_hw_threads.restype = C.c_int

    def hw_threads(self) -> int:
        return int(self.lib.ref_hybrid_hw_threads())

    def compress(self, data: np.ndarray, chunk: int, level: int, threads: int = 1):
        """returns (seconds, out buffer with stride, out_stride, sizes)"""
        n = data.size // chunk
        sizes = np.full(n, chunk, dtype=np.uint64)
        stride = chunk + chunk // 255 + 515
        out = np.empty(stride * n, dtype=np.uint8)
        out_sizes = np.zeros(n, dtype=np.uint64)
        p = C.POINTER(C.c_size_t)
        s = self.lib.ref_hybrid_cpu_compress(_ptr(data), sizes.ctypes.data_as(p), chunk, n, _ptr(out), stride,
                                             out_sizes.ctypes.data_as(p), level, threads)
        if s < 0:
            raise RuntimeError("reference HybridEngine FORCE_CPU compress failed")
        return s, out, stride, out_sizes

    def decompress(self, comp: np.ndarray, stride: int, sizes: np.ndarray, chunk: int, threads: int = 1):
        n = len(sizes)
        out = np.empty(chunk * n, dtype=np.uint8)
        out_sizes = np.zeros(n, dtype=np.uint64)
        p = C.POINTER(C.c_size_t)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        s = self.lib.ref_hybrid_cpu_decompress(_ptr(comp), sizes.ctypes.data_as(p), stride, n, _ptr(out), chunk,
                                               out_sizes.ctypes.data_as(p), threads)
        if s < 0:
            raise RuntimeError("reference HybridEngine FORCE_CPU decompress failed")
        return s, out, out_sizes
