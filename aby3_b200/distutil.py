"""Rank bookkeeping for bench.py's multi-GPU mode.  The hot path shards by independent
output row blocks (one process per GPU, no data-path collective); torch.distributed is
only used for the barrier around the timed region and to combine per-rank results."""
import os


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def row_block(total_rows, rank, world):
    """Rows [r0, r1) of the global product owned by `rank` (contiguous, balanced)."""
    base, rem = divmod(total_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def combine(dist, device, ms, launches, units):
    """max over ranks of the device time, sums of launches and of the units processed."""
    if dist is None:
        return ms, launches, units
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([float(launches), float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), int(s[0].item()), float(s[1].item())


def throughput(units_total, ms_max):
    return units_total / (ms_max * 1e-3)
