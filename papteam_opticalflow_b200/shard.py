"""Host-side sharding of independent frame pairs over ranks / devices: pair p belongs
to rank p % world.  No collective touches the data path; the only cross-rank traffic is the
max-reduce of the per-rank timing that bench.py reports."""


def pairs_for_rank(npairs, rank, world):
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    return list(range(rank, npairs, world))


def consecutive_pairs(nframes):
    """Frame t with frame t+1, the pairing rule of Par/InputCreation/TestImagePairGenerator.py:151-171."""
    return [(t, t + 1) for t in range(max(0, nframes - 1))]


def max_over_ranks(dist, value, device=None):
    """Timing of a multi-rank step is the slowest rank's (task contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
