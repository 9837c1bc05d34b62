"""Multi-GPU layer: independent meshes shard one-per-rank, no collective on the solve path.

A single mesh is never split across GPUs: CG needs two global reductions and a halo exchange per ~25 us iteration,
which NVLink latency would dominate.  What does shard is the *batch of meshes* (BASELINE config 4: 8 x 250K verts):
mesh i -> rank i mod world.  The only communication is the trivial gather of results / timings at the end, done with
torch.distributed (NCCL on the B200s, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def assign(n_items, rank, world):
    """Round-robin ownership: items rank, rank+world, ... (mesh i -> GPU i mod N)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return list(range(rank, n_items, world))


def is_dist():
    return dist.is_available() and dist.is_initialized()


def world():
    return dist.get_world_size() if is_dist() else 1


def rank():
    return dist.get_rank() if is_dist() else 0


def barrier():
    if is_dist():
        dist.barrier()


def max_over_ranks(value, device="cpu"):
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-measured time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_solutions(x):
    """All-gather equally shaped per-rank results (V,k) -> (world, V, k).  ~3 MB per 250K-vertex mesh."""
    if not is_dist():
        return x.unsqueeze(0)
    x = x.contiguous()
    out = torch.empty((world() * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x)      # concatenation along dim 0 (the layout gloo and NCCL both accept)
    return out.view((world(),) + tuple(x.shape))
