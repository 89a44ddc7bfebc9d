"""N > 1 host logic on CPU: contiguous sharding of independent subints and the
host-side gather, over torch.distributed with the gloo backend (world size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from pulseportraiture_b200.multigpu import shard_range
    for n in (0, 1, 7, 8, 10000, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from pulseportraiture_b200.multigpu import shard_range, gather_results
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    a, b = shard_range(n, rank, world)
    # stand-in for a per-rank fit: results are functions of the global subint index
    idx = np.arange(a, b)
    local = {"params": np.stack([idx * 1.0, idx * 2.0], axis=1), "chi2": idx * 10.0,
             "nfeval": (idx % 3).astype(np.int32)}
    merged = gather_results(local, dst=0)
    dist.barrier()
    if rank == 0:
        q.put({k: v.tolist() for k, v in merged.items()})
    dist.destroy_process_group()


def test_gloo_world2_gather_in_rank_order():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(11)
    assert np.array_equal(np.array(merged["chi2"]), idx * 10.0)
    assert np.array_equal(np.array(merged["params"])[:, 1], idx * 2.0)
    assert np.array_equal(np.array(merged["nfeval"]), idx % 3)


def _shm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from pulseportraiture_b200.multigpu import shard_range, SharedGather
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    a, b = shard_range(n, rank, world)
    sg = SharedGather(6, 3, tag="test")
    outs = []
    for rep in range(3):            # slots are reused call after call
        idx = np.arange(a, b, dtype=np.float64)
        local = np.stack([idx, idx * 2 + rep, idx * idx], axis=1)
        outs.append(sg.gather(local))
    sg.close()
    if rank == 0:
        q.put(outs)
    dist.destroy_process_group()


def test_shared_memory_gather_world2():
    import multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(11, dtype=np.float64)
    for rep, o in enumerate(outs):
        assert o.shape == (11, 3)
        assert np.array_equal(o, np.stack([idx, idx * 2 + rep, idx * idx], axis=1))
