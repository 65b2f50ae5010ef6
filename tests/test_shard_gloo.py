"""N > 1 host path on CPU: two gloo ranks compute the utterance partition independently, agree without
communicating, cover every utterance exactly once, and the bench's max-over-ranks time reduction works."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jatts_b200.shard import shard_utterances


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lens, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_utterances(lens, world)[rank]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # the bench reduction: elapsed = max over ranks, work = sum over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        a = torch.tensor([float(sum(lens[i] for i in mine))], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        if rank == 0:
            q.put((gathered, float(t), float(a)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    lens = [int(x) for x in torch.randint(200, 400, (37,), generator=torch.Generator().manual_seed(0))]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, tmax, total = q.get()
    assert gathered == shard_utterances(lens, 2)
    assert sorted(i for s in gathered for i in s) == list(range(len(lens)))
    assert tmax == 11.0 and total == float(sum(lens))
