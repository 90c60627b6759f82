"""Instance-batch data parallelism: the only parallel axis of the path.

Instances (independent NLP decision vectors: multi-start / Monte-Carlo guesses) share
only read-only problem data, so rank r of G evaluates the contiguous slice
[r*B/G, (r+1)*B/G) with no data-path collective; the one exchange
step is the gather of per-instance results (converged decision vectors, costs) at the
end, done with torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""


def shard_range(total, rank, world):
    """Contiguous, balanced slice [lo, hi) of `total` instances owned by `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total, world):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_rows(local, total, group=None):
    """All-gather row-sharded results: `local` is this rank's (b_r, ...) slice in
    shard_range order; returns the (total, ...) tensor on every rank.  Uneven shards are
    padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        assert local.shape[0] == total
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    assert local.shape[0] == sizes[dist.get_rank(group)]
    cap = max(sizes)
    pad = local
    if local.shape[0] < cap:
        pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    pieces = [out[r * cap:r * cap + sizes[r]] for r in range(world)]
    return torch.cat(pieces, dim=0)


def run_sharded(fn, rows, group=None, device=None):
    """Data-parallel driver of the instance axis: this rank runs `fn` on its contiguous slice of
    `rows` (numpy, (B, ...)) and every per-instance result is all-gathered, so all ranks return
    the same dict as a single process calling fn(rows) would (instances are independent, so the
    entries are bitwise identical to the unsharded run).  `fn` returns a dict of numpy arrays /
    lists with one entry per local instance.  Without an initialised process group (or with one
    rank) this is fn(rows).  `device`: where the collective runs (a CUDA device for NCCL, None =
    CPU for gloo)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return fn(rows)
    total = len(rows)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(total, rank, world)
    local = fn(rows[lo:hi])
    out = {}
    for key in sorted(local):
        val = local[key]
        if isinstance(val, (list, tuple)):                  # e.g. exit messages: gather as objects
            parts = [None] * world
            dist.all_gather_object(parts, list(val), group=group)
            out[key] = [v for part in parts for v in part]
            continue
        arr = np.ascontiguousarray(val)
        t = torch.from_numpy(arr)
        if device is not None:
            t = t.to(device)
        out[key] = gather_rows(t, total, group=group).cpu().numpy()
    return out


def best_instance(result):
    """Index of the converged instance (exit mode 0) with the lowest cost, or of the lowest cost
    overall if none converged -- the multi-start pick."""
    import numpy as np
    fun = np.asarray(result["fun"], dtype=float)
    ok = np.asarray(result["status"]) == 0
    cand = np.where(ok, fun, np.inf) if ok.any() else fun
    return int(np.argmin(cand))
