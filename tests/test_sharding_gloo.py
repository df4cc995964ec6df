"""CPU, world_size 2 over gloo: the N>1 host logic of the request-sharded layout (no kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vllm_kvcompress_b200 import sharding
    mine = sharding.shard_requests(513, world, rank)
    cfg = sharding.broadcast_config([256, 32, 128] if rank == 0 else [0, 0, 0], "cpu")
    ms = sharding.max_over_ranks([1.5 + rank, 2.0 - rank], "cpu")
    tput = sharding.aggregate_throughput(256, ms[0])
    # every rank must see all other ranks' shards to prove they are disjoint and complete
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, cfg, ms, tput, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_request_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, mine, cfg, ms, tput, gathered in out:
        assert cfg == [256, 32, 128]                       # rank 0's header everywhere
        assert ms == [2.5, 2.0]                            # max over ranks, element-wise
        assert tput == pytest.approx(256 * 2 / 2.5e-3)     # whole-job aggregate
        assert mine == list(range(rank, 513, world))
        allr = sorted(i for part in gathered for i in part)
        assert allr == list(range(513))                    # disjoint + complete


def test_shard_requests_validation():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from vllm_kvcompress_b200 import sharding
    assert sharding.shard_requests(5, 1, 0) == [0, 1, 2, 3, 4]
    assert sharding.shard_requests(2, 4, 3) == []
    with pytest.raises(ValueError):
        sharding.shard_requests(5, 2, 2)
