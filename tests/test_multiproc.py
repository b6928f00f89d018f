"""N > 1 path on CPU: two processes over gloo exercise the proof all-gather and the rank -> seed sharding that bench.py
uses with NCCL on the GPUs (the only exchange step of the path: proofs of different pictures are independent)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    assert bench.dist_env() == (rank, world, rank)
    proof = bytes([rank + 1]) * (1000 + 10 * rank)          # ragged on purpose
    got = bench.gather_proofs(proof, torch.device("cpu"), dist)
    seeds = [10_000 + (k * world + rank) for k in range(3)]
    q.put((rank, [len(g) for g in got], [g[:1] for g in got], seeds))
    dist.destroy_process_group()


def test_gather_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lens, heads, seeds in res:
        assert lens == [1000, 1010] and heads == [b"\x01", b"\x02"]
    all_seeds = res[0][3] + res[1][3]
    assert len(set(all_seeds)) == 6            # every (step, rank) proves under its own challenge seed
