"""CPU: the N > 1 path with world_size 2 over gloo.  The path shards by profile with no data-path collective
(SURVEY §8e); what is tested is exactly what a multi-GPU run adds: the contiguous split, per-profile RNG streams
that make a profile's result independent of the shard it lands in, the host-side gather, and the
max-over-ranks / sum-over-ranks aggregation bench.py reports.  The sampler standing in for the GPU here is the
CPU oracle (test infrastructure) — the GPU counterpart of this test is tests/test_gpu_parity.py::test_shard_invariance."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    import torch.distributed as dist

    from fitoct_b200 import _abi as abi
    from fitoct_b200 import shard, synth
    from oracle import oracle as O

    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = shard.shard_range(n_total, rank, world)
    S = synth.make_profiles(last - first, first_id=first, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=5, ids=S["ids"])
    cfg = abi.default_cfg(n_warmup=30, n_iter=60, seed=77, chains=2)
    out = O.sample(abi.FOCT_EXPGP, b, last - first, abi.default_spec(), cfg, n_threads=2)
    dist.barrier()
    t_local = 1.0 + rank  # fake per-rank step time: aggregation must take the max
    mx, sm = shard.aggregate([t_local], [float(last - first), float(out["n_leapfrog"].sum())], dist)
    rows = shard.gather_rows(out["summary"][:, :, 0], dist)
    if rank == 0:
        q.put((mx.tolist(), sm.tolist(), rows))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    import torch.multiprocessing as mp

    from fitoct_b200 import _abi as abi
    from fitoct_b200 import synth
    from oracle import oracle as O

    n_total, world = 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    mx, sm, rows = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank run of the whole batch
    S = synth.make_profiles(n_total, modulated_only=True)
    b = abi.make_problems_dense(S["x"], S["Y"], S["UY"], S["theta0"], S["Sigma0"], Nn=5, ids=S["ids"])
    cfg = abi.default_cfg(n_warmup=30, n_iter=60, seed=77, chains=2)
    one = O.sample(abi.FOCT_EXPGP, b, n_total, abi.default_spec(), cfg, n_threads=4)
    assert mx == [2.0]                                   # max over ranks of the step time
    assert sm[0] == n_total                              # units processed by all ranks
    assert sm[1] == one["n_leapfrog"].sum()              # identical chains whatever the sharding
    np.testing.assert_array_equal(rows, one["summary"][:, :, 0])   # gathered in profile order, bit-identical
