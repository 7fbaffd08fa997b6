"""Multi-process (gloo, world_size 2, CPU) test of the batch-axis sharding + the single all-gather of finished
latents: uneven batch, padded shards, rank order preserved."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from invertible_cd_b200 import dist_utils
    dist_utils.init("gloo")
    assert dist_utils.get_world_size() == world and dist_utils.get_rank() == rank
    start, stop, per = dist_utils.shard_batch(n_total)
    # every "latent" is filled with its global prompt index
    local = torch.stack([torch.full((4, 8, 8), float(i)) for i in range(start, stop)]) if stop > start \
        else torch.zeros(0, 4, 8, 8)
    full = dist_utils.gather_latents(local, n_total, per)
    out_q.put((rank, start, stop, per, full[:, 0, 0, 0].tolist()))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_shard_and_gather_world2():
    world, n_total = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort()
    assert [(r[1], r[2], r[3]) for r in results] == [(0, 3, 3), (3, 5, 3)]
    for r in results:
        assert r[4] == [0.0, 1.0, 2.0, 3.0, 4.0]          # same gathered batch, global order, padding dropped


def test_shard_batch_arithmetic():
    from invertible_cd_b200.dist_utils import shard_batch
    assert [shard_batch(32, r, 8)[:2] for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert shard_batch(3, 3, 4) == (3, 3, 1) and shard_batch(0, 0, 2) == (0, 0, 0)
    assert shard_batch(7, 0, 1) == (0, 7, 7)
