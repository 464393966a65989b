"""Data-parallel plumbing.  The reference is single-process; DP is new.

One process per GPU (torchrun), full parameter replica per rank, the minibatch sharded over
its batch axis, ONE allreduce per optimizer step over a single flat buffer
``[gradients of the un-normalised cost sum(cost*mask) || sum(mask)]``; every rank then divides by
``sum(mask) + 1e-5`` (model.py:784 is a masked mean over the GLOBAL batch), clips and applies Adam,
so replicas stay bit-identical.  ``torch.distributed`` is plumbing only (NCCL on GPU, gloo in the
CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_rows(x, rank, world, axis):
    """Rows [r*B/N, (r+1)*B/N) of the batch axis (SURVEY 8e 'Partitioning')."""
    if x is None or world == 1:
        return x
    B = x.shape[axis]
    assert B % world == 0, 'global batch %d not divisible by %d ranks' % (B, world)
    per = B // world
    idx = [slice(None)] * x.ndim
    idx[axis] = slice(rank * per, (rank + 1) * per)
    return x[tuple(idx)]


def shard_batch(batch, rank, world):
    """batch: dict with time-major frame tensors and batch-major text tensors (reference layout)."""
    out = dict(batch)
    for k in ('features', 'features_mask', 'feedback_noise', 'gmm_unis', 'gmm_normals'):
        if batch.get(k) is not None:
            out[k] = shard_rows(batch[k], rank, world, 1)
    for k in ('labels', 'labels_mask', 'speaker'):
        if batch.get(k) is not None:
            out[k] = shard_rows(batch[k], rank, world, 0)
    return out


def allreduce_flat(flat, group=None):
    """SUM-allreduce of the flat [grads || sum(mask)] buffer, in place.  Identity for one rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_scale_from(flat):
    """1 / (sum(mask) + 1e-5) from the last element of the (all-reduced) flat buffer."""
    return 1.0 / (float(flat[-1]) + 1e-5)
