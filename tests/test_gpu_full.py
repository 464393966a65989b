"""GPU tests at BASELINE sizes and through the entry points: size-independent properties where the CPU oracle is
too slow (tensor-core engine vs its SIMT twin on the same job tables, run-to-run determinism), a short oracle
comparison at the base hidden size, one optimizer step against the oracle, and the train / sample scripts."""
import os

import numpy as np
import pytest
import torch

from oracle.parrot_oracle import OracleAdamClip
from tests import util

pytestmark = pytest.mark.gpu

BASE = dict(input_dim=420, output_dim=63, rnn_h_dim=1024, readouts_dim=1024, weak_feedback=True, which_cost='MSE',
            num_characters=43, attention_type='graves', attention_size=10, attention_alignment=0.15,
            encoder_type='bidirectional', encoder_dim=128)


def _device_model(cfg, impl, seed=0, gain=0.5):
    from parrot_b200.model import Parrot
    m = Parrot(gemm_impl=impl, **cfg)
    m.initialize(seed=seed, gain=gain)
    return m


def test_base_config_tensor_core_matches_simt_twin():
    """BASELINE configs[1] shapes (B=64, H=1024, U=128), T=48: persistent tcgen05 scan vs the SIMT twin that
    executes the same job tables with fp32 FMAs -- forward outputs, alignment argmax and every gradient."""
    B, T, U = 64, 48, 128
    bt = util.make_batch(BASE, B, T, U, seed=7)
    res = {}
    for impl in ('tcgen05', 'simt'):
        m = _device_model(BASE, impl)
        cost, _, av, _ = m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                        None, 1.0, B)
        g = m.backward()
        torch.cuda.synchronize()
        res[impl] = (cost.item(), [a.cpu().numpy().copy() for a in av],
                     {n: v.cpu().numpy().copy() for n, v in g.items()})
        del m
        torch.cuda.empty_cache()
    a, b = res['tcgen05'], res['simt']
    assert abs(a[0] - b[0]) / abs(b[0]) < 1e-5
    for nm, x, y in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], a[1], b[1]):
        assert util.rel_err(x, y) < 1e-4, nm
    assert (a[1][4].argmax(-1) == b[1][4].argmax(-1)).mean() > 0.999
    for n in a[2]:
        assert util.rel_err(a[2][n], b[2][n]) < 5e-4, n


def test_base_hidden_short_segment_matches_oracle():
    """H = R = 1024, E = 128 (25 M parameters) on a short segment the numpy oracle finishes in seconds."""
    B, T, U = 16, 6, 24
    orc = util.make_oracle(BASE, gain=0.5)
    dev = util.make_device_model(BASE, orc)
    bt = util.make_batch(BASE, B, T, U, seed=2)
    c_o, _, av_o, _ = orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                       None, 1.0, B)
    g_o = orc.backward()
    c_d, _, av_d, _ = dev.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                       None, 1.0, B)
    g_d = dev.backward()
    torch.cuda.synchronize()
    assert abs(c_d.item() - c_o) / abs(c_o) < 1e-3
    assert util.rel_err(av_d[0].cpu().numpy(), av_o[0]) < 1e-3
    assert (av_d[4].cpu().numpy().argmax(-1) == av_o[4].argmax(-1)).all()
    for n in g_o:
        assert util.rel_err(g_d[n].cpu().numpy(), g_o[n]) < 2e-3, n


def test_run_to_run_determinism():
    """Split-K partials are reduced in part order and there are no float atomics: two runs are bit-identical."""
    cfg = dict(util.TINY, rnn_h_dim=256, readouts_dim=256, weak_feedback=True, which_cost='GMM')
    B, T, U = 32, 20, 24
    bt = util.make_batch(cfg, B, T, U, seed=4)
    outs = []
    for _ in range(2):
        m = _device_model(cfg, 'tcgen05', seed=1)
        cost, _, av, _ = m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None,
                                        1.0, B, gmm_noise=(bt['gmm_unis'], bt['gmm_normals']))
        g = m.backward()
        torch.cuda.synchronize()
        outs.append((cost.item(), m.flat_grads.cpu().numpy().copy(), av[4].cpu().numpy().copy()))
    assert outs[0][0] == outs[1][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][2], outs[1][2])


def test_training_step_matches_oracle_optimizer():
    """compute_cost + backward + StepClipping(9) + Adam(1e-4) (train.py:100-108) vs the oracle, two steps."""
    from parrot_b200.algorithms import Adam, CompositeRule, GradientDescent, StepClipping
    cfg = dict(util.TINY, weak_feedback=True, attention_alignment=0.4)
    B, T, U = 8, 10, 12
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    algo = GradientDescent(model=dev, step_rule=CompositeRule([StepClipping(0.05), Adam(1e-3)]))
    opt = OracleAdamClip(orc.shapes, learning_rate=1e-3, threshold=0.05)
    for step in range(2):
        bt = util.make_batch(cfg, B, T, U, seed=20 + step)
        batch = dict(features=bt['features'], features_mask=bt['features_mask'], labels=bt['labels'],
                     labels_mask=bt['labels_mask'], start_flag=1.0)
        algo.process_batch(batch, B)
        orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
        norm = opt.step(orc.params, orc.backward())
        torch.cuda.synchronize()
        assert abs(algo.stats[0].item() - norm) / norm < 2e-3
        assert algo.stats[1].item() < 1.0            # the clip is active in this test
        vals = dev.get_parameter_values()
        for n in orc.params:
            assert np.abs(vals[n] - orc.params[n]).max() < 2e-5, n


def test_train_and_sample_scripts(tmp_path, monkeypatch):
    """train.py / sample.py entry points with the reference's flags on synthetic data."""
    monkeypatch.setenv('RESULTS_DIR', str(tmp_path))
    import importlib
    import train
    import sample
    importlib.reload(train)
    args = ['--experiment_name', 'smoke', '--rnn_h_dim', '64', '--readouts_dim', '64', '--encoder_dim', '32',
            '--input_dim', '24', '--batch_size', '4', '--seq_size', '20', '--save_every', '3', '--steps', '4',
            '--weak_feedback', 'True', '--save_dir', str(tmp_path)]
    train.main(args)
    d = os.path.join(str(tmp_path), 'vctk')
    assert os.path.exists(os.path.join(d, 'pkl', 'best_smoke.npz'))
    assert os.path.exists(os.path.join(d, 'config', 'smoke.pkl'))
    x, lengths = sample.main(['--experiment_name', 'smoke', '--num_samples', '3', '--num_steps', '30',
                              '--save_dir', str(tmp_path)])
    assert x.shape == (3, 30, 63) and len(lengths) == 3 and np.isfinite(x).all()
    assert os.path.exists(os.path.join(d, 'samples', 'best_sample_0.mgc'))
