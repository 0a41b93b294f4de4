"""Multi-GPU plumbing: forward scene completion shards over INDEPENDENT 64^3 blocks -- rows of
different batch indices share no grid cell and no rule, eval-mode BatchNorm is per row -- so block i goes to
rank i mod world, there is no halo and no data-path collective.  The only collective is one broadcast of
the flattened parameter buffer (2.57 MB fp32) from rank 0; timings are reduced with MAX over ranks.
The reference itself has no distributed code (one process, one GPU: train.py:77, test_scene.py:53).
"""
import torch
import torch.distributed as dist


def shard_blocks(n_blocks, rank, world):
    """Round-robin: block i -> rank i mod world (BASELINE.json configs[3])."""
    return list(range(rank, n_blocks, world))


def flatten_state(model):
    ts = [t for t in model.state_dict().values() if t.is_floating_point()]
    return ts, torch.cat([t.detach().reshape(-1).float() for t in ts]) if ts else torch.zeros(0)


def broadcast_parameters(model, src=0):
    """One collective: rank `src` -> all (NCCL on GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    ts, flat = flatten_state(model)
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
            off += n
    return model


def max_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device):
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
