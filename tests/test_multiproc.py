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


def _input_worker(rank, world, port, path, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    os.environ["ZKCNN_B200_TEST_EMU"] = "1"
    import zkcnn_b200
    from zkcnn_b200._binding import HostLib
    zkcnn_b200._host = HostLib(os.path.join(ROOT, "tests", "emu", "_build", "libzkcnn_host_emu.so"))   # (no GPU here: the parser lives in the host library)
    v = zkcnn_b200.load_input(path if rank == 0 else "/nonexistent/only-rank-0-reads-the-file", dist=dist)
    q.put((rank, len(v), float(v[:5].sum()), float(v[-3:].sum())))
    dist.destroy_process_group()


def test_input_cache_and_broadcast_world2(tmp_path):
    """the text input is parsed once (binary cache next to it) and only rank 0 reads it; the other rank gets the
    numbers through the process group (gloo here, NCCL on the GPUs)"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_synthetic_input as gen
    vals = gen.generate("lenet", 11)
    path = str(tmp_path / "lenet.csv")
    gen.write_text(vals, path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_input_worker, args=(r, 2, port, path, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = vals.astype(np.float64)
    for rank, n, head, tail in res:
        assert n == len(want) and abs(head - want[:5].sum()) < 1e-9 and abs(tail - want[-3:].sum()) < 1e-9
    cache = np.load(path + ".f64.npy")
    assert len(cache) == len(want) + 2 and (cache[2:].astype(np.float32) == vals).all()      # %.9g round-trips float32


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
