"""Process-group setup for batch-axis data parallelism: one process per GPU, NCCL over NVLink
(mirrors utils/dist_utils.py:8-24; the path has no collective inside the sampling loop)."""
import os

import torch
import torch.distributed as dist


def get_world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def get_rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init(backend=None):
    """Fills env defaults so a single process forms a world of 1 (as the reference does), then joins the group."""
    for key, val in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29500"), ("RANK", "0"), ("LOCAL_RANK", "0"),
                     ("WORLD_SIZE", "1")):
        os.environ.setdefault(key, val)
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, init_method="env://")
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))


def shard_batch(n_items, rank=None, world=None):
    """Contiguous, padded batch-axis shard: every rank gets ceil(n/world) slots so the final all-gather has equal
    counts (the reference's np.array_split shards would hang its all_gather on uneven counts,
    running/sd1.5/generate.py:29-39,375-383). Returns (start, stop, per_rank)."""
    rank = get_rank() if rank is None else rank
    world = get_world_size() if world is None else world
    per = (n_items + world - 1) // world
    start = min(rank * per, n_items)
    return start, min(start + per, n_items), per


def gather_latents(local, n_total, per_rank):
    """One all-gather of the finished latents [per_rank, C, H, W] (padded) -> [n_total, C, H, W] on every rank.
    Replaces the reference's all_gather of decoded uint8 images (running/sd1.5/generate.py:375-383)."""
    world = get_world_size()
    if local.shape[0] < per_rank:
        pad = torch.zeros((per_rank - local.shape[0],) + tuple(local.shape[1:]), device=local.device,
                          dtype=local.dtype)
        local = torch.cat([local, pad], 0)
    if world == 1:
        return local[:n_total]
    out = torch.empty((world * per_rank,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:n_total]
