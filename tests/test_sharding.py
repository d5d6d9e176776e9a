"""Multi-process path on CPU (gloo, world_size 2 and 3): frame-range sharding + the single counter all-reduce.
The device decoder is replaced by the CPU oracle here (it consumes the same (seed, global frame id) streams), so the
test checks exactly what the multi-GPU run relies on: shards are disjoint, cover everything, and sum to the whole."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    from qec_ldpc_b200.sharding import shard_range
    for total in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == total
            for (a, ca), (b, _) in zip(parts, parts[1:]):
                assert a + ca == b
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    from qec_ldpc_b200.sharding import run_sharded
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oc = Oracle().code_qc(3, 3, 6, 7, 2, 3)
    imp = np.load(os.path.join(ROOT, "tests", "golden", "codes.npz"))
    shp = imp["C1_iMinusP_shape"]
    oc.set_logical(np.unpackbits(imp["C1_iMinusP"], axis=1)[:, :shp[1]].astype(np.int32))
    got = run_sharded(lambda first, n: oc.run_depolarizing(99, first, n, 0.05, 20, 1)["counters"], total, first_frame=5)
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counters_equal_unsharded(oracle, world):
    import torch.multiprocessing as mp
    from util import oracle_code
    total = 3001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = oracle_code(oracle, "C1").run_depolarizing(99, 5, total, 0.05, 20)["counters"]
    assert np.array_equal(got, want)
