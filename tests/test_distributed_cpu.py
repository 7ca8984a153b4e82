"""world_size-2 gloo test of the N > 1 host logic: per-rank shards of one corpus,
per-rank encode (CPU oracle standing in for the GPU kernel), counts gathered,
global offsets by exclusive scan, concatenation == encoding the whole text."""
from __future__ import annotations

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, result_path: str):
    sys.path.insert(0, HERE)
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from _oracle import Oracle
    from wordpiece_b200 import synth
    from wordpiece_b200 import global_offsets

    g = synth.generator("en")
    blocks = 2
    shard = g.generate(blocks * synth.BLOCK, seed=4, first_block=rank * blocks, n_threads=1)
    o = Oracle(g.spec.vocab)
    ids = o.encode(shard)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([ids.size], dtype=torch.int64))
    offs = global_offsets([int(c.item()) for c in counts])
    total = sum(int(c.item()) for c in counts)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t.item()) == float(world)
    # gather the ids on rank 0 at their global offsets
    out = torch.zeros(total, dtype=torch.int32)
    out[offs[rank]:offs[rank] + ids.size] = torch.from_numpy(ids)
    dist.all_reduce(out, op=dist.ReduceOp.SUM)
    if rank == 0:
        whole = g.generate(world * blocks * synth.BLOCK, seed=4, n_threads=2)
        expect = o.encode(whole)
        np.save(result_path, np.array([int(np.array_equal(out.numpy(), expect)), total, expect.size]))
    dist.destroy_process_group()


def test_two_rank_sharded_encode(tmp_path):
    result = str(tmp_path / "result.npy")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, result), nprocs=2, join=True)
    ok, total, expect = np.load(result).tolist()
    assert ok == 1 and total == expect
