"""Ciphertext-parallel sharding across one process per GPU.

The path shards by independent units -- ciphertexts of a batch, samples of a dataset -- with no data-path collective:
every rank owns a contiguous block of units, evaluates them with its own context and keys, and only the timing is
combined (MAX over ranks of the device time, SUM of the units).  Works with any torch.distributed backend: NCCL on the
GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def my_units(total, rank, world):
    """Contiguous block partition of `total` units: range of the units owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def combine(units_done, seconds, device="cpu"):
    """Whole-job throughput: all ranks' units / the slowest rank's time.  Returns (total_units, max_seconds, units_per_s)."""
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_done)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    total, worst = float(u.item()), float(t.item())
    return total, worst, total / worst


def gather_logits(logits, device="cpu"):
    """Rank 0 receives every rank's 20 logits (the only data that ever leaves a rank in sample-parallel inference)."""
    x = torch.as_tensor(logits, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [x.cpu().numpy()]
    out = [torch.empty_like(x) for _ in range(dist.get_world_size())]
    dist.all_gather(out, x)
    return [o.cpu().numpy() for o in out]
