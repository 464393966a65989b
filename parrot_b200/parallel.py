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


class Comm(object):
    """The C-ABI collective of the data-parallel step (include/parrot_b200.h ``parrot_comm_*``): NCCL bound at run
    time, ONE in-place fp32 SUM allreduce per optimizer step on the caller's CUDA stream.  ``torch.distributed`` is used
    for the rendezvous only (broadcast of the 128-byte NCCL unique id)."""

    def __init__(self, nranks, rank, unique_id=None):
        import ctypes as C
        from . import _lib
        self._C, self._lib = C, _lib.load()
        self.nranks, self.rank = int(nranks), int(rank)
        self.ptr = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        _lib.check(self._lib.parrot_comm_init(self.nranks, self.rank, buf, C.byref(self.ptr)))

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().parrot_comm_unique_id(buf))
        return buf.raw

    def allreduce(self, flat):
        """In-place SUM of a contiguous float32 tensor on the current CUDA stream (identity for one rank)."""
        from . import _lib
        C = self._C
        assert flat.dtype == torch.float32 and flat.is_contiguous()
        stream = torch.cuda.current_stream(flat.device).cuda_stream if flat.is_cuda else 0
        _lib.check(self._lib.parrot_comm_allreduce(self.ptr, C.c_void_p(flat.data_ptr()), flat.numel(),
                                                   C.c_void_p(stream)))
        return flat

    def info(self):
        C = self._C
        n, r, v = C.c_int32(), C.c_int32(), C.c_int32()
        self._lib.parrot_comm_info(self.ptr, C.byref(n), C.byref(r), C.byref(v))
        return n.value, r.value, v.value

    def __del__(self):
        try:
            if self.ptr:
                self._lib.parrot_comm_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


_COMM = None


def get_comm():
    """The process-wide NCCL communicator of the C ABI, created on first use from the torch.distributed rendezvous
    (rank 0 draws the unique id, everyone receives it).  None for a single process or a CPU-only (gloo) run."""
    global _COMM
    if _COMM is not None:
        return _COMM
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if not torch.cuda.is_available():
        return None
    box = [Comm.unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    _COMM = Comm(dist.get_world_size(), dist.get_rank(), box[0])
    return _COMM


def allreduce_flat(flat, group=None):
    """SUM-allreduce of the flat [grads || sum(mask)] buffer, in place.  Identity for one rank.  CUDA tensors go
    through the C-ABI NCCL communicator on the current stream; CPU tensors (the gloo tests of the host logic) through
    torch.distributed."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        comm = get_comm() if (flat.is_cuda and group is None) else None
        if comm is not None:
            comm.allreduce(flat)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def grad_scale_from(flat):
    """1 / (sum(mask) + 1e-5) from the last element of the (all-reduced) flat buffer."""
    return 1.0 / (float(flat[-1]) + 1e-5)
