"""Host-side logic of the N>1 path on CPU: world_size-2 gloo. Rows of C|D are dealt r % world, every rank
contributes a padded block, the all-gather lays the blocks out as the engine's echelon expects
(gamba_b200/parallel.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_bot, nnp, out):
    sys.path.insert(0, ROOT)
    from gamba_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    full = rng.integers(0, 32003, size=(n_bot, nnp), dtype=np.int64).astype(np.uint32)   # the D' every rank would compute
    mine = list(parallel.shard_rows(n_bot, rank, world))
    per = parallel.rows_per_rank(n_bot, world)
    block = np.zeros((per, nnp), dtype=np.uint32)
    block[: len(mine)] = full[mine]
    g = parallel.allgather_rows(torch.from_numpy(block.astype(np.int64))).numpy().astype(np.uint32)
    ok = g.shape == (per * world, nnp)
    for r in range(n_bot):
        ok &= bool(np.array_equal(g[parallel.gathered_row_of(r, n_bot, world)], full[r]))
    # padding rows stay zero
    used = {parallel.gathered_row_of(r, n_bot, world) for r in range(n_bot)}
    for i in range(per * world):
        if i not in used:
            ok &= not g[i].any()
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t))
    dist.destroy_process_group()


def _run(n_bot, nnp, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_bot, nnp, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return q.get(timeout=10)


def test_shards_partition_the_rows():
    from gamba_b200 import parallel
    for n_bot in (0, 1, 7, 2000, 2001):
        for world in (1, 2, 4, 8):
            rows = [r for k in range(world) for r in parallel.shard_rows(n_bot, k, world)]
            assert sorted(rows) == list(range(n_bot))
            assert all(len(parallel.shard_rows(n_bot, k, world)) <= parallel.rows_per_rank(n_bot, world) for k in range(world))
            slots = {parallel.gathered_row_of(r, n_bot, world) for r in range(n_bot)}
            assert len(slots) == n_bot and (not slots or max(slots) < world * parallel.rows_per_rank(n_bot, world))


def test_allgather_layout_world2_gloo():
    assert _run(n_bot=13, nnp=37) == 1
    assert _run(n_bot=8, nnp=5) == 1
