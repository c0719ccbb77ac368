"""Multi-GPU plumbing of the receiver bank: one process per GPU, receivers sharded by contiguous ranges.

Virtual receivers are independent: rank r of N owns receivers
[r * S / N, (r + 1) * S / N) in its own t41rx context and no collective touches the data path.
torch.distributed is used for exactly two things: reducing timings (max over ranks) and the
optional gather of the spectrum / waterfall rows to rank 0 (what a display front end would read).
Backend: "nccl" on GPUs (bench.py), "gloo" in the CPU tests (tests/test_sharding_cpu.py).
"""
import torch
import torch.distributed as dist


def shard_range(n_streams, rank, world):
    """Receivers [first, first + count) owned by `rank`; ranges are contiguous, disjoint and cover 0..n_streams."""
    if not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad shard request")
    first = n_streams * rank // world
    last = n_streams * (rank + 1) // world
    return first, last - first


def max_over_ranks(value, device=None):
    """max of a Python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(rows, n_streams_total, dst=0):
    """Gather per-rank row tensors [count_r, n_rows, 512] (int16 spectrum or uint16-as-int16 waterfall) to
    rank `dst` in receiver order.  Returns the [n_streams_total, n_rows, 512] tensor on dst, None elsewhere.
    Shards may be ragged (shard_range): every rank pads to the largest shard for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = [shard_range(n_streams_total, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = torch.zeros((pad,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    buf[: rows.shape[0]] = rows
    wire = buf.view(torch.uint8)                      # byte view: every backend moves uint8
    out = [torch.empty_like(wire) for _ in range(world)] if rank == dst else None
    dist.gather(wire, out, dst=dst)
    if rank != dst:
        return None
    out = [o.view(rows.dtype) for o in out]
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
