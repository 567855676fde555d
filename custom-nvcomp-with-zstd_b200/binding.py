"""ctypes binding of include/cuda_zstd_batch_c.h plus torch-tensor conveniences.

Mirrors the reference's pointer-array batch interface (NvcompV5BatchManager,
include/cuda_zstd_nvcomp.h:93-134 in the reference): same argument meaning (sizes are in = capacity,
out = bytes written), same error codes (cuda_zstd::Status through status_to_nvcomp_error).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
from typing import Optional, Tuple

import numpy as np
import torch

# CUDA_ZSTD_B200_LIB selects an instrumented build of the same library (tools/esd_prof.py); there is no other fallback
LIB_PATH = os.environ.get("CUDA_ZSTD_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcuda_zstd_b200.so")
_LIB: Optional[C.CDLL] = None


class Status(enum.IntEnum):
    SUCCESS = 0
    ERROR_GENERIC = 1
    ERROR_INVALID_PARAMETER = 2
    ERROR_OUT_OF_MEMORY = 3
    ERROR_CUDA_ERROR = 4
    ERROR_INVALID_MAGIC = 5
    ERROR_CORRUPT_DATA = 6
    ERROR_BUFFER_TOO_SMALL = 7
    ERROR_DICTIONARY_MISMATCH = 9
    ERROR_CHECKSUM_FAILED = 10
    ERROR_COMPRESSION = 12
    ERROR_NOT_IMPLEMENTED = 24
    ERROR_UNSUPPORTED_FORMAT = 28


def status_to_nvcomp_error(status: int) -> int:
    """The reference's lossy Status -> int map (src/cuda_zstd_nvcomp.cpp:75-96)."""
    return status if status in (0, 2, 3, 4, 6, 7, 10, 12) else 1


# every symbol include/*.h declares extern "C"
EXPORTS = [
    "cuda_zstd_batch_create", "cuda_zstd_batch_destroy", "cuda_zstd_batch_get_max_compressed_size",
    "cuda_zstd_batch_get_compress_temp_size", "cuda_zstd_batch_get_decompress_temp_size", "cuda_zstd_batch_compress",
    "cuda_zstd_batch_decompress", "cuda_zstd_batch_compress_nosync", "cuda_zstd_batch_decompress_nosync",
    "cuda_zstd_batch_scan_sizes", "cuda_zstd_batch_pack", "cuda_zstd_batch_last_launch_count", "cuda_zstd_batch_error_string",
    "cuda_zstd_batch_get_host_decompress_temp_size", "cuda_zstd_batch_get_host_compress_temp_size", "cuda_zstd_batch_decompress_host",
    "cuda_zstd_batch_compress_host_packed", "cuda_zstd_batch_compress_sharded", "cuda_zstd_batch_decompress_sharded",
    "cuda_zstd_create_manager", "cuda_zstd_destroy_manager", "cuda_zstd_compress", "cuda_zstd_decompress",
    "cuda_zstd_get_compress_workspace_size", "cuda_zstd_get_decompress_workspace_size", "cuda_zstd_train_dictionary",
    "cuda_zstd_destroy_dictionary", "cuda_zstd_set_dictionary", "cuda_zstd_get_error_string", "cuda_zstd_is_error",
    "nvcomp_zstd_create_manager_v5", "nvcomp_zstd_destroy_manager_v5", "nvcomp_zstd_compress_async_v5",
    "nvcomp_zstd_decompress_async_v5", "nvcomp_zstd_get_compress_temp_size_v5", "nvcomp_zstd_get_decompress_temp_size_v5",
    "nvcomp_zstd_get_metadata_v5",
    "cuda_zstd_pipeline_create", "cuda_zstd_pipeline_destroy", "cuda_zstd_pipeline_compress",
    "cuda_zstd_hybrid_create", "cuda_zstd_hybrid_create_default", "cuda_zstd_hybrid_destroy", "cuda_zstd_hybrid_compress",
    "cuda_zstd_hybrid_decompress", "cuda_zstd_hybrid_max_compressed_size", "cuda_zstd_hybrid_query_routing",
]


class ShardC(C.Structure):                 # cuda_zstd_shard_t (include/cuda_zstd_batch_c.h)
    _fields_ = [("device", C.c_int), ("mgr", C.c_void_p), ("d_in_ptrs", C.c_void_p), ("d_in_sizes", C.c_void_p), ("d_out_ptrs", C.c_void_p),
                ("d_out_sizes", C.c_void_p), ("d_statuses", C.c_void_p), ("num_chunks", C.c_size_t), ("d_temp", C.c_void_p),
                ("temp_bytes", C.c_size_t), ("stream", C.c_void_p), ("d_all_sizes", C.c_void_p), ("d_all_offsets", C.c_void_p)]


class HybridConfigC(C.Structure):       # cuda_zstd_hybrid_config_t (include/cuda_zstd_hybrid.h)
    _fields_ = [("mode", C.c_uint), ("cpu_size_threshold", C.c_size_t), ("gpu_device_threshold", C.c_size_t),
                ("compression_level", C.c_int), ("enable_profiling", C.c_int), ("cpu_thread_count", C.c_uint)]


class HybridResultC(C.Structure):       # cuda_zstd_hybrid_result_t
    _fields_ = [("backend_used", C.c_uint), ("input_location", C.c_uint), ("output_location", C.c_uint),
                ("total_time_ms", C.c_double), ("transfer_time_ms", C.c_double), ("compute_time_ms", C.c_double),
                ("throughput_mbps", C.c_double), ("input_bytes", C.c_size_t), ("output_bytes", C.c_size_t),
                ("compression_ratio", C.c_float)]


# callbacks of include/pipeline_manager.hpp
PIPELINE_INPUT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t))
PIPELINE_OUTPUT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_size_t)


def load_library() -> C.CDLL:
    """Loads the in-tree CUDA library.  Raises if it is missing: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    lib.cuda_zstd_batch_create.restype = vp
    lib.cuda_zstd_batch_create.argtypes = [i32, i32]
    lib.cuda_zstd_batch_destroy.argtypes = [vp]
    lib.cuda_zstd_batch_get_max_compressed_size.restype = sz
    lib.cuda_zstd_batch_get_max_compressed_size.argtypes = [vp, sz]
    for f in (lib.cuda_zstd_batch_get_compress_temp_size, lib.cuda_zstd_batch_get_decompress_temp_size):
        f.restype = sz
        f.argtypes = [vp, vp, sz]
    for f in (lib.cuda_zstd_batch_compress, lib.cuda_zstd_batch_decompress):
        f.restype = i32
        f.argtypes = [vp, vp, vp, sz, vp, vp, vp, sz, vp]
    for f in (lib.cuda_zstd_batch_compress_nosync, lib.cuda_zstd_batch_decompress_nosync):
        f.restype = i32
        f.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, sz, vp]
    lib.cuda_zstd_batch_get_host_decompress_temp_size.restype = sz
    lib.cuda_zstd_batch_get_host_decompress_temp_size.argtypes = [vp, vp, vp, sz]
    lib.cuda_zstd_batch_get_host_compress_temp_size.restype = sz
    lib.cuda_zstd_batch_get_host_compress_temp_size.argtypes = [vp, vp, sz]
    lib.cuda_zstd_batch_decompress_host.restype = i32
    lib.cuda_zstd_batch_decompress_host.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, sz, vp]
    lib.cuda_zstd_batch_compress_host_packed.restype = i32
    lib.cuda_zstd_batch_compress_host_packed.argtypes = [vp, vp, vp, sz, vp, sz, vp, vp, vp, sz, vp]
    for f in (lib.cuda_zstd_batch_compress_sharded, lib.cuda_zstd_batch_decompress_sharded):
        f.restype = i32
        f.argtypes = [C.POINTER(ShardC), i32]
    lib.cuda_zstd_batch_scan_sizes.restype = i32
    lib.cuda_zstd_batch_scan_sizes.argtypes = [vp, sz, u64, vp, vp]
    lib.cuda_zstd_batch_pack.restype = i32
    lib.cuda_zstd_batch_pack.argtypes = [vp, vp, vp, sz, vp, vp]
    lib.cuda_zstd_batch_last_launch_count.restype = i32
    lib.cuda_zstd_batch_last_launch_count.argtypes = [vp]
    lib.cuda_zstd_batch_error_string.restype = C.c_char_p
    lib.cuda_zstd_batch_error_string.argtypes = [i32]
    # single-buffer C APIs
    for name in ("cuda_zstd_create_manager", "nvcomp_zstd_create_manager_v5"):
        getattr(lib, name).restype = vp
        getattr(lib, name).argtypes = [i32]
    for name in ("cuda_zstd_destroy_manager", "nvcomp_zstd_destroy_manager_v5"):
        getattr(lib, name).argtypes = [vp]
    for name in ("cuda_zstd_compress", "cuda_zstd_decompress", "nvcomp_zstd_compress_async_v5", "nvcomp_zstd_decompress_async_v5"):
        getattr(lib, name).restype = i32
        getattr(lib, name).argtypes = [vp, vp, sz, vp, C.POINTER(sz), vp, sz, vp]
    for name in ("cuda_zstd_get_compress_workspace_size", "cuda_zstd_get_decompress_workspace_size",
                 "nvcomp_zstd_get_compress_temp_size_v5", "nvcomp_zstd_get_decompress_temp_size_v5"):
        getattr(lib, name).restype = sz
        getattr(lib, name).argtypes = [vp, sz]
    lib.cuda_zstd_get_error_string.restype = C.c_char_p
    lib.cuda_zstd_get_error_string.argtypes = [i32]
    lib.cuda_zstd_is_error.restype = i32
    lib.cuda_zstd_is_error.argtypes = [i32]
    lib.cuda_zstd_set_dictionary.restype = i32
    lib.cuda_zstd_set_dictionary.argtypes = [vp, vp]
    lib.cuda_zstd_train_dictionary.restype = vp
    lib.cuda_zstd_train_dictionary.argtypes = [vp, vp, sz, sz]
    lib.cuda_zstd_pipeline_create.restype = vp
    lib.cuda_zstd_pipeline_create.argtypes = [i32, i32, sz, i32]
    lib.cuda_zstd_pipeline_destroy.argtypes = [vp]
    lib.cuda_zstd_pipeline_compress.restype = i32
    lib.cuda_zstd_pipeline_compress.argtypes = [vp, PIPELINE_INPUT_FN, PIPELINE_OUTPUT_FN, vp]
    lib.cuda_zstd_hybrid_create.restype = vp
    lib.cuda_zstd_hybrid_create.argtypes = [C.POINTER(HybridConfigC)]
    lib.cuda_zstd_hybrid_create_default.restype = vp
    lib.cuda_zstd_hybrid_create_default.argtypes = []
    lib.cuda_zstd_hybrid_destroy.argtypes = [vp]
    for f in (lib.cuda_zstd_hybrid_compress, lib.cuda_zstd_hybrid_decompress):
        f.restype = i32
        f.argtypes = [vp, vp, sz, vp, C.POINTER(sz), C.c_uint, C.c_uint, C.POINTER(HybridResultC), vp]
    lib.cuda_zstd_hybrid_max_compressed_size.restype = sz
    lib.cuda_zstd_hybrid_max_compressed_size.argtypes = [vp, sz]
    lib.cuda_zstd_hybrid_query_routing.restype = C.c_uint
    lib.cuda_zstd_hybrid_query_routing.argtypes = [vp, sz, C.c_uint, C.c_uint, i32]
    _LIB = lib
    return lib


def _addr(x) -> Optional[int]:
    """Address of a numpy array (host) or torch tensor (device or host); None passes NULL."""
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return int(x)


def _stream_handle(stream) -> int:
    if stream is None:
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, torch.cuda.Stream):
        return stream.cuda_stream
    return int(stream)


class ZstdBatchCodec:
    """Host-side mirror of NvcompV5BatchManager over the C ABI (one per thread, like the reference)."""

    def __init__(self, level: int = 3, checksum: bool = False):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("ZstdBatchCodec needs a CUDA device: the batch path has no CPU route")
        self.level, self.checksum = level, checksum
        self.h = self.lib.cuda_zstd_batch_create(level, int(checksum))
        if not self.h:
            raise RuntimeError("cuda_zstd_batch_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cuda_zstd_batch_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- size queries -------------------------------------------------------------------------
    def max_compressed_size(self, n: int) -> int:
        return int(self.lib.cuda_zstd_batch_get_max_compressed_size(self.h, n))

    def compress_temp_size(self, num_chunks: int, sizes: Optional[np.ndarray] = None) -> int:
        """`sizes`: host array of uncompressed chunk sizes (optional)."""
        p = np.ascontiguousarray(sizes, dtype=np.uint64).ctypes.data if sizes is not None else None
        return int(self.lib.cuda_zstd_batch_get_compress_temp_size(self.h, p, num_chunks))

    def decompress_temp_size(self, num_chunks: int, sizes: Optional[np.ndarray] = None) -> int:
        """`sizes`: host array of COMPRESSED sizes; it sizes the fast path's literal/sequence pools
        (without it 16 KB frames are assumed; a small workspace only routes more chunks to the slower kernel)."""
        p = np.ascontiguousarray(sizes, dtype=np.uint64).ctypes.data if sizes is not None else None
        return int(self.lib.cuda_zstd_batch_get_decompress_temp_size(self.h, p, num_chunks))

    # ---- raw pointer-table calls (tables: numpy uint64 on the host, or torch int64 on the device) ----
    def compress_tables(self, in_ptrs, in_sizes, n, out_ptrs, out_sizes, workspace: Optional[torch.Tensor], stream=None) -> int:
        return int(self.lib.cuda_zstd_batch_compress(self.h, _addr(in_ptrs), _addr(in_sizes), n, _addr(out_ptrs), _addr(out_sizes),
                                                     _addr(workspace), workspace.numel() if workspace is not None else 0,
                                                     _stream_handle(stream)))

    def decompress_tables(self, in_ptrs, in_sizes, n, out_ptrs, out_sizes, workspace: Optional[torch.Tensor], stream=None) -> int:
        return int(self.lib.cuda_zstd_batch_decompress(self.h, _addr(in_ptrs), _addr(in_sizes), n, _addr(out_ptrs), _addr(out_sizes),
                                                       _addr(workspace), workspace.numel() if workspace is not None else 0,
                                                       _stream_handle(stream)))

    def compress_nosync(self, d_in_ptrs, d_in_sizes, n, d_out_ptrs, d_out_sizes, d_status, workspace, stream=None) -> int:
        return int(self.lib.cuda_zstd_batch_compress_nosync(self.h, _addr(d_in_ptrs), _addr(d_in_sizes), n, _addr(d_out_ptrs),
                                                            _addr(d_out_sizes), _addr(d_status), _addr(workspace), workspace.numel(),
                                                            _stream_handle(stream)))

    def decompress_nosync(self, d_in_ptrs, d_in_sizes, n, d_out_ptrs, d_out_sizes, d_status, workspace, stream=None) -> int:
        return int(self.lib.cuda_zstd_batch_decompress_nosync(self.h, _addr(d_in_ptrs), _addr(d_in_sizes), n, _addr(d_out_ptrs),
                                                              _addr(d_out_sizes), _addr(d_status), _addr(workspace), workspace.numel(),
                                                              _stream_handle(stream)))

    def last_launch_count(self) -> int:
        return int(self.lib.cuda_zstd_batch_last_launch_count(self.h))

    # ---- host-resident batches: payloads in (pinned) HOST memory, staged in waves by the library ----
    def host_decompress_temp_size(self, comp_sizes: np.ndarray, caps: np.ndarray) -> int:
        a, b = np.ascontiguousarray(comp_sizes, dtype=np.uint64), np.ascontiguousarray(caps, dtype=np.uint64)
        return int(self.lib.cuda_zstd_batch_get_host_decompress_temp_size(self.h, a.ctypes.data, b.ctypes.data, len(a)))

    def host_compress_temp_size(self, sizes: np.ndarray) -> int:
        a = np.ascontiguousarray(sizes, dtype=np.uint64)
        return int(self.lib.cuda_zstd_batch_get_host_compress_temp_size(self.h, a.ctypes.data, len(a)))

    def decompress_host(self, h_in_ptrs: np.ndarray, in_sizes: np.ndarray, h_out_ptrs: np.ndarray, out_sizes: np.ndarray,
                        workspace: torch.Tensor, statuses: Optional[np.ndarray] = None, stream=None) -> int:
        """cuda_zstd_batch_decompress_host: all four tables are host uint64 arrays of HOST addresses; out_sizes is in/out."""
        n = len(in_sizes)
        return int(self.lib.cuda_zstd_batch_decompress_host(self.h, h_in_ptrs.ctypes.data, in_sizes.ctypes.data, n, h_out_ptrs.ctypes.data,
                                                            out_sizes.ctypes.data, _addr(statuses), workspace.data_ptr(), workspace.numel(),
                                                            _stream_handle(stream)))

    def compress_host_packed(self, h_in_ptrs: np.ndarray, in_sizes: np.ndarray, h_packed, packed_cap: int, h_offsets: np.ndarray,
                             workspace: torch.Tensor, statuses: Optional[np.ndarray] = None, stream=None) -> int:
        """cuda_zstd_batch_compress_host_packed: frames packed back to back in h_packed, h_offsets[0..n] their exclusive scan."""
        n = len(in_sizes)
        return int(self.lib.cuda_zstd_batch_compress_host_packed(self.h, h_in_ptrs.ctypes.data, in_sizes.ctypes.data, n, _addr(h_packed), packed_cap,
                                                                 h_offsets.ctypes.data, _addr(statuses), workspace.data_ptr(), workspace.numel(),
                                                                 _stream_handle(stream)))

    # ---- tensor conveniences ----------------------------------------------------------------------
    def compress_chunks(self, data: torch.Tensor, chunk: int, workspace: Optional[torch.Tensor] = None
                        ) -> Tuple[torch.Tensor, np.ndarray, int]:
        """Compress a contiguous uint8 device tensor in `chunk`-byte pieces.  Returns (out, sizes, stride):
        frame i is out[i*stride : i*stride + sizes[i]].  Raises on any failure."""
        assert data.is_cuda and data.dtype == torch.uint8 and data.is_contiguous()
        total = data.numel()
        n = (total + chunk - 1) // chunk
        stride = (self.max_compressed_size(chunk) + 15) // 16 * 16
        out = torch.empty(n * stride, dtype=torch.uint8, device=data.device)
        idx = np.arange(n, dtype=np.uint64)
        in_ptrs = data.data_ptr() + idx * np.uint64(chunk)
        in_sizes = np.minimum(np.uint64(chunk), np.uint64(total) - idx * np.uint64(chunk)).astype(np.uint64)
        out_ptrs = out.data_ptr() + idx * np.uint64(stride)
        out_sizes = np.full(n, stride, dtype=np.uint64)
        if workspace is None:
            workspace = torch.empty(self.compress_temp_size(n), dtype=torch.uint8, device=data.device)
        rc = self.compress_tables(in_ptrs, in_sizes, n, out_ptrs, out_sizes, workspace)
        if rc != 0:
            raise RuntimeError(f"batch compress failed: {self.lib.cuda_zstd_batch_error_string(rc).decode()}")
        return out, out_sizes, stride

    def decompress_chunks(self, comp: torch.Tensor, offsets: np.ndarray, sizes: np.ndarray, chunk: int,
                          workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, np.ndarray]:
        """Decompress frames comp[offsets[i] : offsets[i]+sizes[i]] into a contiguous tensor with stride `chunk`."""
        assert comp.is_cuda and comp.dtype == torch.uint8
        n = len(sizes)
        out = torch.empty(n * chunk, dtype=torch.uint8, device=comp.device)
        idx = np.arange(n, dtype=np.uint64)
        in_ptrs = (comp.data_ptr() + np.asarray(offsets, dtype=np.uint64)).astype(np.uint64)
        in_sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        out_ptrs = out.data_ptr() + idx * np.uint64(chunk)
        out_sizes = np.full(n, chunk, dtype=np.uint64)
        if workspace is None:
            workspace = torch.empty(self.decompress_temp_size(n, in_sizes), dtype=torch.uint8, device=comp.device)
        rc = self.decompress_tables(in_ptrs, in_sizes, n, out_ptrs, out_sizes, workspace)
        if rc != 0:
            raise RuntimeError(f"batch decompress failed: {self.lib.cuda_zstd_batch_error_string(rc).decode()}")
        return out, out_sizes

    def scan_sizes(self, d_sizes: torch.Tensor, base: int = 0, stream=None) -> torch.Tensor:
        """Device-side exclusive scan: returns int64 offsets[n+1] (offsets[n] = base + total)."""
        n = d_sizes.numel()
        off = torch.empty(n + 1, dtype=torch.int64, device=d_sizes.device)
        rc = self.lib.cuda_zstd_batch_scan_sizes(d_sizes.data_ptr(), n, base, off.data_ptr(), _stream_handle(stream))
        if rc != 0:
            raise RuntimeError("scan_sizes failed")
        return off

    def pack(self, d_ptrs: torch.Tensor, d_sizes: torch.Tensor, d_offsets: torch.Tensor, packed: torch.Tensor, stream=None):
        rc = self.lib.cuda_zstd_batch_pack(d_ptrs.data_ptr(), d_sizes.data_ptr(), d_offsets.data_ptr(), d_sizes.numel(),
                                           packed.data_ptr(), _stream_handle(stream))
        if rc != 0:
            raise RuntimeError("pack failed")


class ZstdSingle:
    """The reference's single-buffer C API (cuda_zstd_* / nvcomp_zstd_*_v5) over device tensors."""

    def __init__(self, level: int = 3, flavor: str = "cuda_zstd"):
        self.lib = load_library()
        self.flavor = flavor
        create = self.lib.cuda_zstd_create_manager if flavor == "cuda_zstd" else self.lib.nvcomp_zstd_create_manager_v5
        self.h = create(level)
        if not self.h:
            raise RuntimeError("create_manager failed")

    def close(self):
        if getattr(self, "h", None):
            (self.lib.cuda_zstd_destroy_manager if self.flavor == "cuda_zstd" else self.lib.nvcomp_zstd_destroy_manager_v5)(self.h)
            self.h = None

    __del__ = close

    def compress_workspace(self, n: int) -> int:
        f = self.lib.cuda_zstd_get_compress_workspace_size if self.flavor == "cuda_zstd" else self.lib.nvcomp_zstd_get_compress_temp_size_v5
        return int(f(self.h, n))

    def decompress_workspace(self, n: int) -> int:
        f = self.lib.cuda_zstd_get_decompress_workspace_size if self.flavor == "cuda_zstd" else self.lib.nvcomp_zstd_get_decompress_temp_size_v5
        return int(f(self.h, n))

    def compress(self, src, n, dst, cap, ws, ws_bytes, stream=None) -> Tuple[int, int]:
        f = self.lib.cuda_zstd_compress if self.flavor == "cuda_zstd" else self.lib.nvcomp_zstd_compress_async_v5
        size = C.c_size_t(cap)
        rc = f(self.h, _addr(src), n, _addr(dst), C.byref(size), _addr(ws), ws_bytes, _stream_handle(stream))
        return int(rc), int(size.value)

    def decompress(self, src, n, dst, cap, ws, ws_bytes, stream=None) -> Tuple[int, int]:
        f = self.lib.cuda_zstd_decompress if self.flavor == "cuda_zstd" else self.lib.nvcomp_zstd_decompress_async_v5
        size = C.c_size_t(cap)
        rc = f(self.h, _addr(src), n, _addr(dst), C.byref(size), _addr(ws), ws_bytes, _stream_handle(stream))
        return int(rc), int(size.value)


class ZstdPipeline:
    """PipelinedBatchManager (reference src/pipeline_manager.hpp:35-66) over the C ABI: host-resident data in,
    concatenated Zstandard frames out (one frame per batch), H2D / compress / D2H of neighbouring batches overlapped."""

    def __init__(self, level: int = 3, checksum: bool = False, batch_bytes: int = 64 << 20, slots: int = 3):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("ZstdPipeline needs a CUDA device: there is no CPU route")
        self.batch_bytes = batch_bytes
        self.h = self.lib.cuda_zstd_pipeline_create(level, int(checksum), batch_bytes, slots)
        if not self.h:
            raise RuntimeError("cuda_zstd_pipeline_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cuda_zstd_pipeline_destroy(self.h)
            self.h = None

    __del__ = close

    def compress(self, data: np.ndarray, out: Optional[np.ndarray] = None):
        """Streams the host array through the pipeline.  Returns (frames, sizes): the concatenated frames (a view of
        `out` when given, else a new array) and the list of per-batch frame sizes."""
        src = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        total = src.size
        if out is None:
            nb = (total + self.batch_bytes - 1) // self.batch_bytes
            out = np.empty(total + total // 255 + nb * 1100 + 4096, dtype=np.uint8)
        state = {"rd": 0, "wr": 0, "sizes": [], "ovf": False}
        src_addr, out_addr = src.ctypes.data, out.ctypes.data

        def fill(_user, buf, cap, out_len):
            n = min(cap, total - state["rd"])
            if n:
                C.memmove(buf, src_addr + state["rd"], n)
            state["rd"] += n
            out_len[0] = n
            return 1 if state["rd"] < total else 0

        def sink(_user, buf, n):
            if state["wr"] + n > out.size:
                state["ovf"] = True
                return
            C.memmove(out_addr + state["wr"], buf, n)
            state["wr"] += n
            state["sizes"].append(n)

        rc = self.lib.cuda_zstd_pipeline_compress(self.h, PIPELINE_INPUT_FN(fill), PIPELINE_OUTPUT_FN(sink), None)
        if rc != 0:
            raise RuntimeError(f"pipeline compress failed: {self.lib.cuda_zstd_batch_error_string(rc).decode()}")
        if state["ovf"]:
            raise RuntimeError("pipeline output buffer too small")
        return out[: state["wr"]], state["sizes"]


class ZstdHybrid:
    """HybridEngine C API (include/cuda_zstd_hybrid.h): host or device buffers in, host or device buffers out, always on the
    GPU.  Locations: 0 host, 1 device, 3 detect."""
    HOST, DEVICE, UNKNOWN = 0, 1, 3

    def __init__(self, level: int = 3, mode: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("ZstdHybrid needs a CUDA device (there is no CPU path in this library)")
        self.lib = load_library()
        cfg = HybridConfigC(mode, 1 << 20, 64 << 10, level, 0, 0)
        self.h = self.lib.cuda_zstd_hybrid_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("cuda_zstd_hybrid_create failed")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cuda_zstd_hybrid_destroy(self.h)
            self.h = None

    __del__ = close

    def max_compressed_size(self, n: int) -> int:
        return self.lib.cuda_zstd_hybrid_max_compressed_size(self.h, n)

    def query_routing(self, n: int, in_loc: int = 0, out_loc: int = 0, compress: bool = True) -> int:
        return self.lib.cuda_zstd_hybrid_query_routing(self.h, n, in_loc, out_loc, int(compress))

    def _call(self, fn, src, n, dst, cap, in_loc, out_loc, stream):
        size = C.c_size_t(cap)
        res = HybridResultC()
        rc = fn(self.h, _addr(src), n, _addr(dst), C.byref(size), in_loc, out_loc, C.byref(res), _stream_handle(stream))
        return rc, size.value, res

    def compress(self, src, n, dst, cap, in_loc=3, out_loc=3, stream=None):
        return self._call(self.lib.cuda_zstd_hybrid_compress, src, n, dst, cap, in_loc, out_loc, stream)

    def decompress(self, src, n, dst, cap, in_loc=3, out_loc=3, stream=None):
        return self._call(self.lib.cuda_zstd_hybrid_decompress, src, n, dst, cap, in_loc, out_loc, stream)
