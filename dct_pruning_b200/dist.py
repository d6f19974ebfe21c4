"""One process per GPU, torch.distributed for the plumbing (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  The scoring path has exactly one collective: the all-reduce of the
flat per-layer score sums at the end of a run (hooks.ScoreSession.finalize_device)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*; no-op single process.
    Returns (rank, local_rank, world)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def shutdown():
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_sums(flat, group=None):
    """Sum the flat per-layer score buffer (last entry = image count) over all ranks, in place.
    The only collective of the scoring path; a no-op in a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
