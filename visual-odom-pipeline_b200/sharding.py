"""Multi-GPU = independent sequences / frame-pair batches per GPU.

A single frame pair does not shard (2000 points are tens of microseconds of work), so there is no
collective on the data path: rank r of W owns a contiguous block of the B sequences, builds its own
pyramids and tracks its own points.  torch.distributed is used only for the launch plumbing, the
timing barrier and the (tiny, 13 B/point) result gather.
"""
import os


def world_from_env():
    """(rank, local_rank, world_size) from torchrun's environment (1 process if unset)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items, world_size, rank):
    """Block partition: item i belongs to rank floor(i * W / n).  -> (start, stop)"""
    if world_size <= 0 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError("bad shard arguments")
    start = (rank * n_items + world_size - 1) // world_size
    stop = ((rank + 1) * n_items + world_size - 1) // world_size
    return start, stop


def shard_sizes(n_items, world_size):
    return [shard_range(n_items, world_size, r)[1] - shard_range(n_items, world_size, r)[0] for r in range(world_size)]


def init_process_group(backend=None):
    """Initialise torch.distributed from the environment when WORLD_SIZE > 1.  -> (rank, local_rank, world)"""
    rank, local_rank, world = world_from_env()
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29500")
            if backend == "nccl":
                torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def bind_rank_to_cores(local_rank, local_world):
    """Give every rank of one box its own contiguous slice of the host cores (the ranks' Python threads, the helper
    threads of the pageable path and the driver's threads then stop migrating over each other: with 8 processes on one
    host the end-to-end path lost 19 % to that in round 1).  No-op for a single rank or where affinity is unsupported.
    -> the cores this process may use now"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(cores) < 2 * local_world:
            return cores
        per = len(cores) // local_world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except (AttributeError, OSError):
        return []


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value):
    """MAX-reduce a python float over all ranks (device timing: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_shards(local, n_items, dst=0):
    """Gather per-sequence results (a tensor whose dim 0 is this rank's shard) on rank `dst` in global
    sequence order.  Returns the full tensor on dst, None elsewhere.  Not on the timed path."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_items, world)
    pad = max(sizes)
    dev = local.device
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    buf[: local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    if rank != dst:
        return None
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)
