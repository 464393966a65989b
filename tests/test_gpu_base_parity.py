"""Parity of the CUDA path with the ORACLE at the benchmarked configuration (BASELINE.json configs[1]: B=64,
H=R=1024, E=128, U=128, weak feedback) over long segments -- T=800 for the bench line's MSE / init-scale case,
T=200 for GMM k=20 and for "trained-like" gain-0.5 weights.  The oracle's outputs are frozen in
tests/golden/base_*.npz (tests/golden/make_base_fixtures.py, float32 numpy oracle, minutes of CPU per case);
inputs and parameters are regenerated from seeds and guarded by checksums.

Gates (north_star: 1e-3 relative on the emitted frames, bit-exact argmax of the alignment):
  * emitted frames ``next_x``: max-norm relative error <= 1e-3 on the stored time steps, AND element-wise relative
    error <= 5e-3 on every entry above 1e-2 * max (tests/util.rel_err alone is a max-norm gate); per-step sum / l2
    checksums of ALL steps within 1e-3 of the step's l2;
  * ``argmax_u phi``: equal to the float32 oracle on every frame whose oracle top-2 gap exceeds 1e-3 relative, with
    the covered fraction printed and >= 0.9; frames with a smaller non-zero gap are counted and reported (a tie
    within ~1e-5 of the GEMM precision cannot be reproduced by any arithmetic that is not bit-identical to numpy's)
    and must stay below 0.1 % of all frames;
  * cost within 1e-3; every gradient tensor: l2 norm within 2e-3, 2048 sampled entries within 2e-3 of the tensor's
    max.
"""
import os

import numpy as np
import pytest
import torch

from tests import util
from tests.golden import make_base_fixtures as mb

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', list(mb.CASES))
def test_base_config_matches_oracle_fixture(name):
    path = os.path.join(GOLD, name + '.npz')
    assert os.path.exists(path), 'run tests/golden/make_base_fixtures.py'
    fx = np.load(path, allow_pickle=False)
    cfg, gain, T = mb.case_setup(name)
    B, U = mb.B, mb.U
    # parameters and inputs from the same seeds as the generator (initialisation only: the oracle does not compute)
    orc = util.make_oracle(cfg, gain=gain, bias_std=0.1 if gain else 0)
    bt = util.make_batch(cfg, B, T, U, seed=11)
    assert abs(mb.checksum(bt['features']) - float(fx['check_features'])) < 1e-6 * abs(float(fx['check_features']))
    chk = sum(mb.checksum(v) for v in orc.params.values())
    assert abs(chk - float(fx['check_params'])) < 1e-6 * abs(float(fx['check_params']))
    dev = util.make_device_model(cfg, orc)
    del orc
    cost, _, av, _ = dev.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0,
                                      B, gmm_noise=(bt['gmm_unis'], bt['gmm_normals']))
    g = dev.backward()
    torch.cuda.synchronize()
    c_ref = float(fx['cost'])
    assert abs(cost.item() - c_ref) / abs(c_ref) < 1e-3, (cost.item(), c_ref)

    # ---- emitted frames
    nx = av[0].cpu().numpy().astype(np.float64)
    steps = fx['sample_steps']
    ref = fx['next_x_samples'].astype(np.float64)
    got = nx[steps]
    amax = float(fx['next_x_absmax'])
    err = np.abs(got - ref)
    frame_bad = np.zeros(err.shape[:2], bool)
    if cfg['which_cost'] == 'GMM':
        # a sampled frame follows the multinomial draw: a uniform within ~1e-6 of a cumulative mixture weight may pick
        # the neighbouring component; such frames are counted, not compared
        frame_bad = err.max(-1) > 0.05 * amax
        assert frame_bad.mean() < 1e-3, frame_bad.mean()
    ok = ~frame_bad
    max_norm = err[ok].max() / amax
    big = (np.abs(ref) > 1e-2 * amax) & ok[..., None]
    elem = (err[big] / np.abs(ref[big])).max()
    l2_ref, sum_ref = fx['next_x_step_l2'], fx['next_x_step_sum']
    if cfg['which_cost'] == 'MSE':
        l2_dev = np.sqrt((nx ** 2).sum(axis=(1, 2)))
        sum_dev = nx.sum(axis=(1, 2))
        assert (np.abs(l2_dev - l2_ref) <= 1e-3 * l2_ref).all()
        assert (np.abs(sum_dev - sum_ref) <= 1e-3 * l2_ref).all()
    print('%s: next_x max-norm rel err %.2e, element-wise (|ref| > 1e-2 max) %.2e over %d stored steps'
          % (name, max_norm, elem, len(steps)))
    assert max_norm < 1e-3, max_norm
    assert elem < 5e-3, elem

    # ---- alignment argmax
    am = av[4].argmax(-1).cpu().numpy()
    am_ref = fx['argmax_phi'].astype(np.int64)
    gap = fx['phi_top2_gap'].astype(np.float64)
    covered = gap > 1e-3
    near = (gap > 0) & ~covered
    mism_cov = int((am != am_ref)[covered].sum())
    mism_near = int((am != am_ref)[near].sum())
    print('%s: argmax phi: %d frames, covered (top-2 gap > 1e-3) %.4f with %d mismatches; near ties (0 < gap <= 1e-3) '
          '%d frames with %d mismatches' % (name, am.size, covered.mean(), mism_cov, int(near.sum()), mism_near))
    assert covered.mean() >= 0.9
    assert mism_cov == 0
    assert mism_near <= 1e-3 * am.size

    # ---- carried state and gradients
    assert util.rel_err(av[1][-1].cpu().numpy(), fx['k_last']) < 1e-3
    assert util.rel_err(av[2][-1].cpu().numpy(), fx['w_last']) < 1e-3
    names = [str(n) for n in fx['grad_names']]
    sig = fx['grad_sig']
    worst = 0.0
    for i, n in enumerate(names):
        gd = g[n].cpu().numpy().astype(np.float64)
        l2 = np.sqrt((gd ** 2).sum())
        assert abs(l2 - sig[i, 1]) <= 2e-3 * sig[i, 1] + 1e-12, (n, l2, sig[i, 1])
        idx = mb.grad_sample_index(n, gd.size)
        e = np.abs(gd.ravel()[idx] - fx['gsamp:' + n].astype(np.float64)).max() / (sig[i, 2] + 1e-30)
        worst = max(worst, e)
        assert e < 2e-3, (n, e)
    print('%s: worst sampled-gradient error relative to the tensor max %.2e' % (name, worst))
