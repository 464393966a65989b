"""GPU tests of the grouped persistent scans (kernels.cuh scan_fwd_grouped / scan_bwd_grouped): every variant of the
schedule computes the same training step -- other layer-group partitions, no TMEM-resident weight tiles, the attention
stand-alone GRU pre-pass instead of the fused one, another chunk length -- and all of them agree with the one-launch-per-phase path that shares nothing of the
group scheduling (no group barriers, counter-based split-K exchange, generic epilogues)."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

CFG = dict(util.TINY, rnn_h_dim=256, readouts_dim=256, encoder_dim=64, weak_feedback=True, which_cost='GMM',
           attention_alignment=0.3)
B, T, U = 16, 40, 24


def _run(monkeypatch, env, per_phase=False, seed=3):
    from parrot_b200 import _lib
    from parrot_b200.model import Parrot
    for k in ('PARROT_GROUPS_F', 'PARROT_GROUPS_B', 'PARROT_TC', 'PARROT_NO_RESIDENT', 'PARROT_NO_FUSED_PRE'):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    bt = util.make_batch(CFG, B, T, U, seed=11)
    m = Parrot(**CFG)
    m.initialize(seed=seed, gain=0.5)
    kw = dict(gmm_noise=(bt['gmm_unis'], bt['gmm_normals']))
    if per_phase:
        # profiling level 2 = one launch per phase with events around every launch (use_persistent() is false)
        m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B, **kw)
        _lib.load().parrot_set_profiling(m._last.ptr, 2)
    outs = []
    for flag in (1.0, 0.0):                      # two TBPTT segments: learned initial state, then carried state
        cost, _, av, _ = m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None,
                                        flag, B, **kw)
        g = m.backward()
        torch.cuda.synchronize()
        outs.append((cost.item(), [a.cpu().numpy().copy() for a in av], m.flat_grads.cpu().numpy().copy()))
    if per_phase:
        _lib.load().parrot_set_profiling(m._last.ptr, 0)
    del m
    torch.cuda.empty_cache()
    return outs


def _close(a, b, what):
    for seg, (x, y) in enumerate(zip(a, b)):
        assert abs(x[0] - y[0]) <= 2e-6 * abs(y[0]), (what, seg, x[0], y[0])
        for i, (p, q) in enumerate(zip(x[1], y[1])):
            assert util.rel_err(p, q) < 2e-5, (what, seg, i)
        assert (x[1][4].argmax(-1) == y[1][4].argmax(-1)).all(), (what, seg)
        assert util.rel_err(x[2], y[2]) < 2e-4, (what, seg)


VARIANTS = {
    'partition_100_24_24': {'PARROT_GROUPS_F': '100,24,24', 'PARROT_GROUPS_B': '100,24,24'},
    'partition_64_36_48': {'PARROT_GROUPS_F': '64,36,48', 'PARROT_GROUPS_B': '64,36,48'},
    'partition_small_grid': {'PARROT_GROUPS_F': '32,16,16', 'PARROT_GROUPS_B': '32,16,16'},
    'no_resident_tiles': {'PARROT_NO_RESIDENT': '1'},
    'standalone_pre_pass': {'PARROT_NO_FUSED_PRE': '1'},
    'chunk_length_4': {'PARROT_TC': '4'},
    'chunk_length_16': {'PARROT_TC': '16'},
}


def test_grouped_scan_matches_per_phase_launches(monkeypatch):
    ref = _run(monkeypatch, {}, per_phase=True)
    got = _run(monkeypatch, {})
    _close(got, ref, 'default grouped vs per-phase')


@pytest.mark.parametrize('name', sorted(VARIANTS))
def test_schedule_variants_agree(monkeypatch, name):
    ref = _run(monkeypatch, {})
    got = _run(monkeypatch, VARIANTS[name])
    _close(got, ref, name)


def test_grouped_scan_is_deterministic(monkeypatch):
    a = _run(monkeypatch, {})
    b = _run(monkeypatch, {})
    for x, y in zip(a, b):
        assert x[0] == y[0] and np.array_equal(x[2], y[2]) and np.array_equal(x[1][4], y[1][4])
