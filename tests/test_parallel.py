"""CPU tests of the data-parallel plumbing with world_size 2 over gloo.

The compute engine here is the ORACLE (tests may use it; the GPU product path is exercised by the -m gpu
tests): what is under test is the host logic every rank runs -- shard the batch rows, backward of the
UN-normalised cost, one allreduce of the flat [grads || sum(mask)] buffer, divide by the global mask count --
and the claim that it reproduces the single-process global-batch gradient (model.py:784 is a masked mean
over the GLOBAL batch)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import util


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


CFG = dict(util.TINY, weak_feedback=True, attention_alignment=0.4)
B, T, U = 8, 6, 10


def _flat(grads, msum):
    return torch.from_numpy(np.concatenate([g.ravel() for g in grads.values()] + [np.array([msum], np.float32)]))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from parrot_b200 import parallel
    r, w, _ = parallel.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    orc = util.make_oracle(CFG, gain=0.5, encoder_time_axis=1)      # identical replicas
    bt = util.make_batch(CFG, B, T, U, seed=3)
    sh = parallel.shard_batch(bt, rank, world)
    assert sh['features'].shape[1] == B // world and sh['labels'].shape[0] == B // world
    orc.compute_cost(sh['features'], sh['features_mask'], sh['labels'], sh['labels_mask'], None, 1.0, B // world)
    g = orc.backward(unnormalised=True)
    flat = _flat(g, sh['features_mask'][1:].sum())
    parallel.allreduce_flat(flat)
    scale = parallel.grad_scale_from(flat)
    if rank == 0:
        out.put((flat[:-1].numpy() * scale, float(flat[-1])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_global_batch():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, msum = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = util.make_oracle(CFG, gain=0.5, encoder_time_axis=1)
    bt = util.make_batch(CFG, B, T, U, seed=3)
    orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
    ref = np.concatenate([g.ravel() for g in orc.backward().values()])
    assert abs(msum - bt['features_mask'][1:].sum()) < 1e-4
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5


def test_shard_rows_and_single_rank_identity():
    from parrot_b200 import parallel
    x = np.arange(24).reshape(4, 6)
    assert (parallel.shard_rows(x, 1, 2, 0) == x[2:]).all()
    assert (parallel.shard_rows(x, 2, 3, 1) == x[:, 4:]).all()
    assert parallel.shard_rows(x, 0, 1, 0) is x
    t = torch.arange(5.0)
    assert parallel.allreduce_flat(t.clone()).equal(t)       # no process group: identity
    with pytest.raises(AssertionError):
        parallel.shard_rows(x, 0, 5, 0)


def test_segment_sequence_matches_reference_semantics():
    """datasets.py:41-138 with seq_size+1, share_value=1, return_last=False: windows overlap by one frame, the
    first carries start_flag=1, and the tail shorter than min_size + seq_size is dropped."""
    from parrot_b200.datasets import segment_sequence, parrot_stream
    feats = np.arange(130, dtype=np.float32)[:, None, None] * np.ones((1, 2, 3), np.float32)
    mask = np.ones((130, 2), np.float32)
    segs = list(segment_sequence(feats, mask, 51, share_value=1, return_last=False))
    assert [s[2] for s in segs] == [1, 0]
    assert segs[0][0][0, 0, 0] == 0 and segs[0][0][-1, 0, 0] == 50
    assert segs[1][0][0, 0, 0] == 50 and segs[1][0].shape[0] == 51        # one-frame overlap
    st = parrot_stream('vctk', use_speaker=True, batch_size=4, seq_size=20, noise_level=0.1)
    tup = next(iter(st.get_epoch_iterator()))
    d = dict(zip(st.sources, tup))
    assert d['features'].shape[0] == 21 and d['features'].shape[1] == 4 and d['features'].shape[2] == 63
    assert d['labels'].shape[0] == 4 and d['speaker_index'].shape == (4, 1) and d['start_flag'] == 1
    assert d['feedback_noise_level'] == 0.1 and d['features_mask'].shape == (21, 4)
