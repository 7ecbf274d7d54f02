"""Frame sharding for multi-GPU runs.  Stereo pairs are independent (the reference keeps no state between
frames), so frame i goes to rank i mod N and there is no data-path collective; the only
communication is the barrier / max-reduction that times a run."""


def frames_for_rank(n_frames, rank, world):
    """Round-robin assignment (north_star): indices of the frames rank `rank` processes."""
    return list(range(rank, n_frames, world))


def reduce_max_and_sum(dist, device, elapsed_ms, units):
    """All ranks: (max over ranks of elapsed_ms, sum over ranks of units).  `dist` is torch.distributed
    (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(elapsed_ms), float(units)
    import torch
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
