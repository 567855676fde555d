"""B200-native batched Zstandard codec: Python host-side mirror of the C-ABI boundary.

The directory name follows the task layout (it contains hyphens), so import it through
``__graft_entry__.import_package()`` (module name ``custom_nvcomp_with_zstd_b200``).
PyTorch is used for device memory, streams and torch.distributed plumbing only; all compute is in
``libcuda_zstd_b200.so`` (hand-written CUDA for sm_100a).  There is NO CPU fallback: loading fails
loudly if the library has not been built.
"""
from .binding import (  # noqa: F401
    LIB_PATH,
    Status,
    ShardC,
    ZstdBatchCodec,
    ZstdHybrid,
    ZstdPipeline,
    ZstdSingle,
    load_library,
    status_to_nvcomp_error,
)
from .sharding import ShardPlan, gather_sizes, shard_range  # noqa: F401
