"""Multi-GPU parity on hardware: N ranks (one process per GPU, NCCL) decode contiguous shares of one global frame
range with the DEVICE decoder; the all-reduced counters must equal the 1-GPU counters of the same range and the CPU
oracle's.  Same through the C++ CLI (`--gpus N`, one host thread per device).  Skipped below 2 devices; the CPU twin
of this test (gloo, decoder replaced by the oracle) is tests/test_sharding.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, FIRST, SEED, P, MAXIT = 20_011, 7, 4242, 0.05, 50


def _ndev():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import qec_ldpc_b200 as ql
    from qec_ldpc_b200.sharding import run_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    code = ql.Code.qc(4, 5, 10, 61, 9, 49)
    dec = ql.Decoder(code, rank, 8192)  # smaller than a shard: chunking is exercised too
    got = run_sharded(lambda first, n: dec.get_statistics_depolarizing(SEED, first, n, P, MAXIT)["counters"], TOTAL,
                      first_frame=FIRST, device="cuda")
    if rank == 0:
        q.put(got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_n_rank_counters_equal_one_rank_and_oracle(qldpc, oracle, world):
    if _ndev() < world:
        pytest.skip("needs %d CUDA devices" % world)
    import torch.multiprocessing as mp
    from util import oracle_code
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 300) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    code = qldpc.Code.qc(4, 5, 10, 61, 9, 49)
    one = qldpc.Decoder(code, 0, 1 << 15).get_statistics_depolarizing(SEED, FIRST, TOTAL, P, MAXIT)["counters"]
    assert np.array_equal(got, one)
    want = oracle_code(oracle, "C2").run_depolarizing(SEED, FIRST, TOTAL, P, MAXIT)["counters"]
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_cli_counters_independent_of_gpu_count(qldpc, tmp_path):
    if _ndev() < 2:
        pytest.skip("needs 2 CUDA devices")
    from qec_ldpc_b200 import build as b
    cli = b.build_cli()
    texts = []
    for gpus in (1, 2):
        d = tmp_path / ("g%d" % gpus)
        d.mkdir()
        (d / "init.txt").write_text("qc:4,5,10,61,9,49\n0\n0\n30000\n50\n0.05")
        subprocess.run([cli, "init.txt", "--depolarizing", "--seed", "123", "--gpus", str(gpus)], cwd=d, check=True,
                       capture_output=True, text=True)
        files = list((d / "results").glob("*depolarizing*.txt"))
        assert len(files) == 1
        texts.append([ln for ln in files[0].read_text().splitlines() if not ln.startswith("Duration")])
    assert texts[0] == texts[1]


@pytest.mark.gpu
def test_reference_library_agrees_on_the_gpu_box(qldpc, oracle):
    """The unmodified reference (oracle/_ref, prebuilt in the authoring container and shipped with the snapshot) is
    loaded on the GPU box as well, so that the port the other GPU tests check against cannot drift from it unnoticed."""
    from oracle.pyoracle import Reference
    from bench import golden_code_file
    from util import oracle_code
    if not Reference.available():
        pytest.skip("oracle/_ref/libqldpc_ref.so not present")
    rc = Reference().code_from_file(golden_code_file("C2"))
    oc = oracle_code(oracle, "C2")
    nf = 96
    x, z = oc.depolarizing_bulk(SEED, 0, nf, P)
    ref = rc.run_frames(x, z, P, MAXIT, 0, want_out=True)
    port = oc.run_frames(x, z, P, MAXIT, 0, want_out=True)
    assert np.array_equal(ref["outX"], port["outX"]) and np.array_equal(ref["outZ"], port["outZ"])
    code = qldpc.Code.qc(4, 5, 10, 61, 9, 49)
    dec = qldpc.Decoder(code, 0, 4096)
    sx = np.stack([oc.syndrome(0, x[f]) for f in range(nf)])
    sz = np.stack([oc.syndrome(1, z[f]) for f in range(nf)])
    ox, oz, fl, _ = dec.decode_batch(sx, sz, P, MAXIT)
    assert np.array_equal(ox, ref["outX"]) and np.array_equal(oz, ref["outZ"])
    got = dec.get_stats_from_errors(x, z, P, MAXIT)["counters"]
    for name, idx in (("xTested", 1), ("zTested", 2), ("corrected", 3), ("synX", 4), ("synZ", 5), ("logical", 6),
                      ("cvX", 7), ("cvZ", 8)):
        assert int(got[idx]) == ref["counters"][name], name
