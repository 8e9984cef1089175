"""Multi-GPU plumbing: tracks and sequences are independent (the sid loop of run_track_nposes.cpp:193), so ranks share
nothing on the hot path; sequences are split contiguously over the ranks and the per-track results (6 f64 pose
coefficients + per-level iteration counts) are all-gathered ONCE at the end.  torch.distributed only (NCCL on the
GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous [lo, hi) of n_units for `rank`; the first n_units % world ranks get one more."""
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(poses, iters, group=None):
    """poses [T_local, 6] f64, iters [T_local, L] i32 on this rank -> the same for all ranks, rank-major.
    Ranks may hold different numbers of tracks (ragged): sizes are exchanged first."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return poses, iters
    n = torch.tensor([poses.shape[0]], dtype=torch.int64, device=poses.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)

    def pad(x):
        if x.shape[0] == mx:
            return x.contiguous()
        out = x.new_zeros((mx,) + tuple(x.shape[1:]))
        out[:x.shape[0]] = x
        return out

    gp = [poses.new_empty((mx, poses.shape[1])) for _ in range(world)]
    gi = [iters.new_empty((mx, iters.shape[1])) for _ in range(world)]
    dist.all_gather(gp, pad(poses), group=group)
    dist.all_gather(gi, pad(iters), group=group)
    return (torch.cat([g[:s] for g, s in zip(gp, sizes)]), torch.cat([g[:s] for g, s in zip(gi, sizes)]))
