"""Batch sharding over the GPUs of one box: problems are independent, so rank g solves the
contiguous slice [g*ceil(B/G), min(B,(g+1)*ceil(B/G))) with no traffic during the solve; the only collective
is the final gather of the packed command records to rank 0 (NCCL on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(B, world_size, rank):
    per = (B + world_size - 1) // world_size
    lo = min(B, rank * per)
    hi = min(B, (rank + 1) * per)
    return lo, hi


def gather_records(local, B, world_size, rank):
    """local: [n_local, rec] tensor of this rank's shard (device of the backend). Returns the [B, rec] tensor
    on rank 0 (None elsewhere). Shards may be ragged (last ranks shorter or empty)."""
    per = (B + world_size - 1) // world_size
    rec = local.shape[1]
    if world_size == 1:
        return local
    pad = torch.zeros((per, rec), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if rank == 0:
        bufs = [torch.empty_like(pad) for _ in range(world_size)]
        dist.gather(pad, bufs, dst=0)
        out = torch.cat(bufs, dim=0)[:B]
        return out
    dist.gather(pad, None, dst=0)
    return None
