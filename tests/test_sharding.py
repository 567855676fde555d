"""Multi-rank host logic on CPU: chunk sharding and the sizes/offsets gather (gloo, world_size 2)."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_chunks, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.import_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = pkg.ShardPlan(n_chunks, rank, world)
    # pretend compressed size of global chunk i is 1000 + 7*i
    local = torch.tensor([1000 + 7 * i for i in range(plan.lo, plan.hi)], dtype=torch.int64)
    table = pkg.gather_sizes(local, plan)
    from custom_nvcomp_with_zstd_b200.sharding import global_offsets_from_sizes
    off = global_offsets_from_sizes(table)
    q.put((rank, plan.lo, plan.hi, table.tolist(), off.tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [8, 13])
def test_gather_sizes_world2(n_chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_chunks
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [1000 + 7 * i for i in range(n_chunks)]
    acc, off = 0, []
    for w in want:
        off.append(acc); acc += w
    off.append(acc)
    los = sorted((r[1], r[2]) for r in res)
    assert los[0][0] == 0 and los[0][1] == los[1][0] and los[1][1] == n_chunks
    for r in res:
        assert r[3] == want and r[4] == off


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg = ge.import_package()
    for n in (0, 1, 7, 8, 131072, 131073):
        for world in (1, 2, 4, 8):
            covered = 0
            for r in range(world):
                lo, hi = pkg.shard_range(n, r, world)
                assert lo == covered and hi >= lo
                covered = hi
            assert covered == n
