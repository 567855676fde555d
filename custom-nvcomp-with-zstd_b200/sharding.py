"""Multi-GPU plumbing: chunks shard by index, the only exchange is the sizes/offsets gather.

SURVEY.md section 8(e): GPU g owns chunks [g*N/G, (g+1)*N/G); no payload crosses GPUs.  After the
local batch compress each rank has its per-chunk compressed sizes on the device; one all-gather of
those sizes (<= 1 MiB even for 131,072 chunks) over the NCCL/NVLink communicator gives every rank
the global size table, from which the device-side exclusive scan (cuda_zstd_batch_scan_sizes)
produces the global packed offsets.  Works with the gloo backend on CPU tensors for the tests.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.distributed as dist


def shard_range(n_chunks: int, rank: int, world: int):
    """Contiguous, balanced partition by chunk index."""
    lo = n_chunks * rank // world
    hi = n_chunks * (rank + 1) // world
    return lo, hi


@dataclass
class ShardPlan:
    n_chunks: int
    rank: int
    world: int

    @property
    def lo(self) -> int:
        return shard_range(self.n_chunks, self.rank, self.world)[0]

    @property
    def hi(self) -> int:
        return shard_range(self.n_chunks, self.rank, self.world)[1]

    @property
    def count(self) -> int:
        return self.hi - self.lo

    def counts(self) -> List[int]:
        return [shard_range(self.n_chunks, r, self.world)[1] - shard_range(self.n_chunks, r, self.world)[0] for r in range(self.world)]


def gather_sizes(local_sizes: torch.Tensor, plan: ShardPlan, group=None) -> torch.Tensor:
    """All-gather of per-chunk compressed sizes (int64).  Returns the global size table [n_chunks] on
    every rank, on the same device as `local_sizes`.  Ragged shards are padded to the largest shard."""
    assert local_sizes.dtype == torch.int64 and local_sizes.numel() == plan.count
    if plan.world == 1 or not (dist.is_available() and dist.is_initialized()):
        return local_sizes.clone()
    counts = plan.counts()
    width = max(counts)
    padded = torch.zeros(width, dtype=torch.int64, device=local_sizes.device)
    padded[: plan.count] = local_sizes
    out = torch.empty(plan.world * width, dtype=torch.int64, device=local_sizes.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = [out[r * width: r * width + counts[r]] for r in range(plan.world)]
    return torch.cat(parts)


def global_offsets_from_sizes(sizes: torch.Tensor) -> torch.Tensor:
    """Reference (torch) exclusive scan used by the CPU tests; the GPU path uses the CUDA scan kernel."""
    off = torch.zeros(sizes.numel() + 1, dtype=torch.int64, device=sizes.device)
    off[1:] = torch.cumsum(sizes, 0)
    return off
