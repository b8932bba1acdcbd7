"""CPU, world_size 2 over gloo: the host-side N>1 logic (clip dealing, max-over-ranks timing,
result gather).  No kernels are involved — the data path has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mivos_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.clips_of_rank(num_clips, rank, world)
        local = [(c, {"rank": rank, "checksum": c * c + 1}) for c in mine]
        merged = sharding.gather_clip_results(local, num_clips)
        tmax = sharding.max_over_ranks(10.0 + rank)
        dist.barrier()
        q.put((rank, mine, [m["rank"] for m in merged], [m["checksum"] for m in merged], tmax))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    world, num_clips = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    for rank, mine, owners, sums, tmax in out:
        assert owners == [sharding.owner_of_clip(c, world) for c in range(num_clips)]
        assert sums == [c * c + 1 for c in range(num_clips)]
        assert tmax == 11.0  # max over ranks


def test_single_process_paths():
    assert sharding.clips_of_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sharding.gather_clip_results([(0, "a"), (1, "b")], 2) == ["a", "b"]
    assert sharding.max_over_ranks(3.5) == 3.5
