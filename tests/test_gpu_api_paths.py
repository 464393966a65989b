"""Device paths of the Parrot mirror class that the other GPU tests do not reach (VERDICT round 1, weak item 4 and the
advisor's carried-state finding): encoder_type=None, initial_states on the device model, sample_using_input,
GradientDescent.global_cost, and carried TBPTT state across handles of different lengths (T1, T2, T1 with flag 0)."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def test_encoder_type_none_matches_oracle():
    """model.py:235-236 / 276-277: without an encoder the 'labels' ARE the (B, U, input_dim) context features."""
    cfg = dict(util.TINY, encoder_type=None, input_dim=24, num_characters=24, weak_feedback=True,
               attention_alignment=0.4)
    B, T, U = 8, 10, 12
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    bt = util.make_batch(cfg, B, T, U, seed=5)
    ctx = np.random.default_rng(0).standard_normal((B, U, cfg['input_dim'])).astype(np.float32)
    c_o, _, av_o, _ = orc.compute_cost(bt['features'], bt['features_mask'], ctx, bt['labels_mask'], None, 1.0, B)
    g_o = orc.backward()
    c_d, _, av_d, _ = dev.compute_cost(bt['features'], bt['features_mask'], ctx, bt['labels_mask'], None, 1.0, B)
    g_d = dev.backward()
    torch.cuda.synchronize()
    assert abs(c_d.item() - c_o) / abs(c_o) < 1e-3
    assert util.rel_err(av_d[0].cpu().numpy(), av_o[0]) < 1e-3
    assert (av_d[4].cpu().numpy().argmax(-1) == av_o[4].argmax(-1)).all()
    for n in g_o:
        assert util.rel_err(g_d[n].cpu().numpy(), g_o[n]) < 2e-3, n


def test_initial_states_and_state_carry_across_handles():
    """model.py:529-549 order; the carried state belongs to the MODEL: segments of lengths T1, T2, T1 with
    start_flag = 0 reuse a cached (B, T1, U) handle and must continue from the T2 segment, not from stale state."""
    cfg = dict(util.TINY, weak_feedback=True, attention_alignment=0.4)
    B, U = 4, 10
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    st_d = dev.initial_states(B)
    st_o = orc.initial_states(B)
    assert len(st_d) == 10
    for a, b in zip(st_d, st_o):
        assert np.allclose(a.cpu().numpy(), b, atol=1e-6)
    for seg, (T, flag) in enumerate([(12, 1.0), (7, 0.0), (12, 0.0)]):
        bt = util.make_batch(cfg, B, T, U, seed=30 + seg)
        lab, lm = util.make_batch(cfg, B, 12, U, seed=30)['labels'], util.make_batch(cfg, B, 12, U, seed=30)['labels_mask']
        c_o, up_o, av_o, _ = orc.compute_cost(bt['features'], bt['features_mask'], lab, lm, None, flag, B)
        c_d, up_d, av_d, _ = dev.compute_cost(bt['features'], bt['features_mask'], lab, lm, None, flag, B)
        torch.cuda.synchronize()
        assert abs(c_d.item() - c_o) / abs(c_o) < 1e-3, seg
        assert util.rel_err(av_d[0].cpu().numpy(), av_o[0]) < 1e-3, seg
        for (n1, v1), (n2, v2) in zip(up_d, up_o):
            assert n1 == n2 and util.rel_err(v1.cpu().numpy(), v2) < 1e-3, (seg, n1)
    # the third segment reused the first handle: its last_* must equal the model's state, also through initial_states
    last = dev.initial_states(B)
    assert util.rel_err(last[1].cpu().numpy(), dict(up_o)['last_h1']) < 1e-3


def test_sample_using_input_and_global_cost():
    from parrot_b200.algorithms import Adam, CompositeRule, GradientDescent, StepClipping
    cfg = dict(util.TINY, weak_feedback=True, attention_alignment=0.4)
    B, T, U = 6, 9, 11
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    bt = util.make_batch(cfg, B, T, U, seed=8)
    data = dict(features=bt['features'], features_mask=bt['features_mask'], labels=bt['labels'],
                labels_mask=bt['labels_mask'], start_flag=1.0)
    outs = dev.sample_using_input(data, B)                      # model.py:1085-1111 (fixed arity)
    c_o, _, av_o, _ = orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
    assert len(outs) == 6 and util.rel_err(outs[0], av_o[0]) < 1e-3
    algo = GradientDescent(model=dev, step_rule=CompositeRule([StepClipping(9.0), Adam(1e-4)]))
    algo.process_batch(data, B)
    assert abs(algo.global_cost() - c_o) / abs(c_o) < 1e-3      # one rank: the masked mean of model.py:784 itself
