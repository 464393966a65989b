"""GPU parity tests: parrot_b200.Parrot (CUDA, through the C ABI) vs the numpy oracle on seeded inputs.

Gates (BASELINE.json north_star): emitted frames within 1e-3 relative of the oracle, argmax of the
alignment phi EQUAL to the oracle's on every frame whose top-2 gap exceeds 1e-4 relative, coverage >= 0.9 printed
(tests/util.argmax_parity); element-wise relative error of the frames above 1e-2 * max also gated (5e-3).
Gradients are held to 2e-3 of each tensor's max-abs (bf16x3 operands, different summation order).
"""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-3
GRAD_TOL = 2e-3

CASES = {
    'mse_weak': dict(weak_feedback=True),
    'gmm_full_spk': dict(which_cost='GMM', full_feedback=True, use_speaker=True),
    'softmax_att_noise': dict(attention_type='softmax', weak_feedback=True, feedback_noise_level=0.3),
    'no_feedback': dict(),
    'layer_norm_weak': dict(weak_feedback=True, layer_norm=True),
    'layer_norm_gmm_full_spk': dict(which_cost='GMM', full_feedback=True, use_speaker=True, layer_norm=True),
}


def _run_pair(cfg, B, T, U, gain, impl, start_flags=(1.0,), axis=0, align=None, check_grads=True):
    cfg = dict(cfg)
    if align is not None:
        cfg['attention_alignment'] = align
    orc = util.make_oracle(cfg, gain=gain, encoder_time_axis=axis)
    o64 = util.make_oracle(cfg, gain=gain, dtype=np.float64, encoder_time_axis=axis)
    dev = util.make_device_model(cfg, orc, gemm_impl=impl, encoder_time_axis=axis)
    out = []
    for si, sf in enumerate(start_flags):
        bt = util.make_batch(cfg, B, T, U, seed=10 + si)
        spk = bt['speaker'] if cfg.get('use_speaker') else None
        kw = dict(feedback_noise=bt['feedback_noise'], noise_level=cfg.get('feedback_noise_level'))
        c_o, up_o, av_o, _ = orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                               spk, sf, B, gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'], **kw)
        c64, _, av64, _ = o64.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                           spk, sf, B, gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'], **kw)
        c_d, up_d, av_d, _ = dev.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                               spk, sf, B, feedback_noise=bt['feedback_noise'],
                                               noise_level=cfg.get('feedback_noise_level'),
                                               gmm_noise=(bt['gmm_unis'], bt['gmm_normals']))
        torch.cuda.synchronize()
        assert abs(c_d.item() - c_o) / abs(c_o) < FWD_TOL, (c_d.item(), c_o)
        names = ['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att']
        for nm, a, b in zip(names, av_d, av_o):
            if b is None:
                continue
            err = util.rel_err(a.cpu().numpy(), b)
            assert err < FWD_TOL, (nm, err)
        util.argmax_parity(av_d[4].cpu().numpy(), av_o[4], 'vs float32 oracle')
        assert util.rel_err_elementwise(av_d[0].cpu().numpy(), av_o[0]) < 5e-3
        for (n1, v1), (n2, v2) in zip(up_d, up_o):
            assert n1 == n2 and util.rel_err(v1.cpu().numpy(), v2) < FWD_TOL, n1
        if check_grads:
            g_o = orc.backward()
            g_d = dev.backward()
            torch.cuda.synchronize()
            worst = ('', 0.0)
            for n in g_o:
                e = util.rel_err(g_d[n].cpu().numpy(), g_o[n])
                if np.abs(g_o[n]).max() < 1e-12:
                    e = float(np.abs(g_d[n].cpu().numpy()).max())
                if e > worst[1]:
                    worst = (n, e)
            assert worst[1] < GRAD_TOL, worst
            assert abs(dev.flat_grads[-1].item() - bt['features_mask'][1:].sum()) < 1e-3
        out.append((c_d.item(), c_o))
    return out


@pytest.mark.parametrize('impl', ['simt', 'tcgen05'])
@pytest.mark.parametrize('case', sorted(CASES))
def test_compute_cost_and_grads_tiny(case, impl):
    """Config 1 scale (H=64): every option, forward + backward, against the oracle."""
    cfg = dict(util.TINY, **CASES[case])
    _run_pair(cfg, B=8, T=12, U=16, gain=0.5, impl=impl, align=0.4)


@pytest.mark.parametrize('impl', ['simt', 'tcgen05'])
def test_state_carry_two_segments(impl):
    """TBPTT: start_flag=1 then start_flag=0 must carry last_h*, last_k, last_w (model.py:633-643, 786-791)."""
    cfg = dict(util.TINY, weak_feedback=True)
    _run_pair(cfg, B=8, T=10, U=16, gain=0.5, impl=impl, start_flags=(1.0, 0.0), align=0.4)


def test_init_scale_and_intended_encoder_axis():
    """Reference init (N(0, 0.01)) and encoder_time_axis=1."""
    cfg = dict(util.TINY, weak_feedback=True, which_cost='GMM')
    orc_cfg = dict(cfg)
    _run_pair(orc_cfg, B=5, T=9, U=11, gain=None, impl='tcgen05', axis=1)


def test_layer_norm_medium_two_segments():
    """layer_norm=True (model.py:24-34, 571-603, 692-722, 743-746) at H=256, odd batch, carried state."""
    cfg = dict(util.TINY, rnn_h_dim=256, readouts_dim=192, encoder_dim=32, weak_feedback=True, layer_norm=True)
    _run_pair(cfg, B=20, T=12, U=24, gain=0.5, impl='tcgen05', start_flags=(1.0, 0.0), align=0.5)


def test_medium_hidden_odd_batch():
    """H=256, B=20 (padded to 32 rows), ragged masks, C != H."""
    cfg = dict(util.TINY, rnn_h_dim=256, readouts_dim=192, encoder_dim=32, weak_feedback=True)
    _run_pair(cfg, B=20, T=16, U=24, gain=0.5, impl='tcgen05', align=0.5)


@pytest.mark.parametrize('case', ['mse_weak', 'gmm_full_spk', 'layer_norm_weak', 'layer_norm_gmm_full_spk'])
def test_sample_model_matches_oracle(case):
    """Free-running generation (model.py:827-1059) with injected GMM noise."""
    cfg = dict(util.TINY, **CASES[case])
    cfg.update(sampling_bias=0.5, sharpening_coeff=1.3, timing_coeff=1.2, attention_alignment=0.5)
    B, T, U = 6, 20, 16
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    bt = util.make_batch(cfg, B, T, U, seed=5)
    spk = bt['speaker'] if cfg.get('use_speaker') else None
    ref = orc.sample_model(bt['labels'], bt['labels_mask'], None, spk, B, T,
                           gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'])
    out = dev.sample_model(bt['labels'], bt['labels_mask'], None, spk, B, T,
                           gmm_noise=(bt['gmm_unis'], bt['gmm_normals']))
    for nm, a, b in zip(['x', 'k', 'w', 'pi', 'phi', 'pi_att'], out, ref):
        assert util.rel_err(a, b) < FWD_TOL, nm


def test_sample_model_philox_runs():
    cfg = dict(util.TINY, which_cost='GMM', weak_feedback=True)
    orc = util.make_oracle(cfg, gain=0.5)
    dev = util.make_device_model(cfg, orc)
    bt = util.make_batch(cfg, 4, 8, 12, seed=5)
    a = dev.sample_model(bt['labels'], bt['labels_mask'], None, None, 4, 8, seed=1)
    b = dev.sample_model(bt['labels'], bt['labels_mask'], None, None, 4, 8, seed=1)
    c = dev.sample_model(bt['labels'], bt['labels_mask'], None, None, 4, 8, seed=2)
    assert np.isfinite(a[0]).all() and (a[0] == b[0]).all() and not (a[0] == c[0]).all()


@pytest.mark.parametrize('name', ['tiny_mse_graves', 'tiny_gmm_softmax_spk'])
def test_device_matches_golden_fixtures(name):
    """CUDA path vs the committed fixtures tests/golden/*.npz (no oracle execution: parameters, inputs and the
    expected outputs all come from the file).  Two consecutive TBPTT segments: frames, alignment, carried state,
    argmax(phi) bit-exact where the fixture's own top-2 gap is unambiguous, gradient signatures."""
    import os
    from tests.golden.make_golden import CASES, B
    from parrot_b200.model import Parrot
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', name + '.npz'))
    cfg = CASES[name]['cfg']
    dev = Parrot(**cfg)
    dev.initialize()
    names = [str(n) for n in z['grad_names']]
    dev.set_parameter_values({n: z['param:' + n] for n in names})
    for seg, sf in enumerate((1.0, 0.0)):
        p = 'seg%d:' % seg
        spk = z[p + 'in:speaker'] if cfg.get('use_speaker') else None
        cost, updates, av, _ = dev.compute_cost(
            z[p + 'in:features'], z[p + 'in:features_mask'], z[p + 'in:labels'], z[p + 'in:labels_mask'], spk,
            sf, B, gmm_noise=(z[p + 'in:gmm_unis'], z[p + 'in:gmm_normals']))
        assert abs(cost.item() - float(z[p + 'cost'])) / abs(float(z[p + 'cost'])) < FWD_TOL
        for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
            key = p + 'out:' + nm
            if key in z.files and v is not None:
                assert util.rel_err(v.cpu().numpy(), z[key]) < FWD_TOL, nm
        phi = z[p + 'out:phi']
        util.argmax_parity(av[4].cpu().numpy(), phi, 'vs frozen oracle vectors')
        for nm, v in updates:
            assert util.rel_err(v.cpu().numpy(), z[p + 'update:' + nm]) < FWD_TOL, nm
        g = dev.backward()
        torch.cuda.synchronize()
        sig = z[p + 'grad_sig']
        for i, n in enumerate(names):
            l2 = float(np.sqrt((g[n].double() ** 2).sum().item()))
            assert abs(l2 - sig[i, 1]) <= 5e-3 * sig[i, 1] + 1e-7, (n, l2, sig[i, 1])
        for key in z.files:
            if key.startswith(p + 'grad:'):
                n = key[len(p + 'grad:'):]
                assert util.rel_err(g[n].cpu().numpy(), z[key]) < GRAD_TOL, n


@pytest.mark.parametrize('name', ['mse_weak', 'gmm_full_spk_softmax', 'layer_norm_noise', 'layer_norm_gmm_full_spk'])
def test_device_matches_reference_source_fixtures(name):
    """CUDA path vs tests/golden/ref_model_*.npz: the outputs of the reference's OWN Parrot.compute_cost /
    sample_model_fun code executed in float64 on the eager Theano/Blocks stand-in (tests/golden/ref_shim.py,
    make_ref_model_fixtures.py).  No oracle in the loop: parameters, inputs and expectations come from the file."""
    import os
    from tests.golden.make_ref_model_fixtures import CASES as RCASES, B, T, SAMP
    from parrot_b200.model import Parrot
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_model_%s.npz' % name))
    cfg = dict(util.TINY, **RCASES[name])
    params = {k[len('param:'):]: z[k].astype(np.float32) for k in z.files if k.startswith('param:')}
    dev = Parrot(**cfg)
    dev.initialize()
    dev.set_parameter_values(params)
    f32 = lambda k: z[k].astype(np.float32)
    for seg, sf in enumerate((1.0, 0.0)):
        p = 'seg%d:' % seg
        spk = z[p + 'in:speaker'] if cfg.get('use_speaker') else None
        cost, updates, av, _ = dev.compute_cost(
            f32(p + 'in:features'), f32(p + 'in:features_mask'), z[p + 'in:labels'], f32(p + 'in:labels_mask'), spk,
            sf, B, feedback_noise=f32(p + 'in:feedback_noise'), noise_level=cfg.get('feedback_noise_level'),
            gmm_noise=(f32(p + 'in:gmm_unis'), f32(p + 'in:gmm_normals')))
        assert abs(cost.item() - float(z[p + 'cost'])) / abs(float(z[p + 'cost'])) < FWD_TOL
        for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
            assert util.rel_err(v.cpu().numpy(), z[p + 'out:' + nm]) < FWD_TOL, (nm, seg)
        phi = z[p + 'out:phi']
        util.argmax_parity(av[4].cpu().numpy(), phi, 'vs reference-source fixtures')
        for nm, v in updates:
            assert util.rel_err(v.cpu().numpy(), z[p + 'update:' + nm]) < FWD_TOL, nm
    smp = Parrot(**dict(cfg, **SAMP))
    smp.initialize()
    smp.set_parameter_values(params)
    spk = z['samp:in:speaker'] if cfg.get('use_speaker') else None
    out = smp.sample_model(z['samp:in:labels'], f32('samp:in:labels_mask'), None, spk, B, T,
                           gmm_noise=(f32('samp:in:gmm_unis'), f32('samp:in:gmm_normals')))
    for nm, a in zip(['x', 'k', 'w', 'pi', 'phi', 'pi_att'], out):
        assert util.rel_err(a, z['samp:out:' + nm]) < FWD_TOL, nm
