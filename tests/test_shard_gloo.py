"""N>1 path on CPU: world_size-2 gloo.  Each rank derives its flat output range
with shard_range, produces its shard of the C5 inputs with the counter-based
generator, computes its shard (the oracle stands in for the kernels -- there is
no GPU here), and the shards are all-gathered and compared with the
single-process result.  No collective touches the compute path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n, align, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        import simplemath_b200 as smb
        from simplemath_b200 import shard
        orc = oracle.c_oracle()
        b, e = smb.shard_range(n, rank, world, align)
        a_loc = orc.fill_uniform_f32(b, e - b, 1, -1.0, 1.0)
        b_loc = orc.fill_uniform_f32(b, e - b, 2, -1.0, 1.0)
        add_loc = orc.elementwise("add", a_loc, [1], b_loc, [1], [e - b]) if e > b else np.empty(0, np.float32)
        pow_loc = orc.array_scalar("pow", np.abs(a_loc) + np.float32(0.01), 2.5)
        full_add = shard.gather_shards(torch.from_numpy(add_loc), n, align).numpy()
        full_pow = shard.gather_shards(torch.from_numpy(pow_loc), n, align).numpy()
        worst = shard.max_over_ranks(float(rank + 1))
        # sharded dot product: per-rank partial (oracle) + one scalar all-reduce; int32 wraps exactly
        ia = (a_loc * 1e6).astype(np.int32)
        ib = (b_loc * 1e6).astype(np.int32)
        part = int(orc.dot(ia, ib)) if e > b else 0
        total = int(shard.allreduce_scalar_sum(part, dtype=torch.int64)) & 0xFFFFFFFF
        if rank == 0:
            a = orc.fill_uniform_f32(0, n, 1, -1.0, 1.0)
            bb = orc.fill_uniform_f32(0, n, 2, -1.0, 1.0)
            ok = np.array_equal(full_add, orc.elementwise("add", a, [1], bb, [1], [n]))
            ok &= np.array_equal(full_pow, orc.array_scalar("pow", np.abs(a) + np.float32(0.01), 2.5))
            ok &= worst == float(world)
            ok &= total == (int(orc.dot((a * 1e6).astype(np.int32), (bb * 1e6).astype(np.int32))) & 0xFFFFFFFF)
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,align", [(100_003, 1024), (65536, 4096), (1000, 8)])
def test_two_rank_sharded_add_pow_matches_single_process(n, align):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, align, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
