"""CPU tests that pin the oracle (the reference ships no tests or fixtures):
finite differences, float32 vs float64, sampler vs teacher forcing, frozen golden vectors,
reference known-answers that exist off the hot path."""
import os

import numpy as np
import pytest

from oracle import parrot_oracle as O
from tests import util

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
SMALL = dict(input_dim=12, output_dim=5, rnn_h_dim=8, readouts_dim=7, num_characters=11,
             attention_size=3, encoder_dim=4, k_gmm=3, num_speakers=4, speaker_dim=6,
             encoder_type='bidirectional')


def _fd_check(cfg, axis=0, start_flag=1.0, n_probe=2):
    m = util.make_oracle(cfg, gain=1.0, dtype=np.float64, encoder_time_axis=axis, bias_std=0.3)
    B, T, U = 3, 4, 5
    bt = util.make_batch(cfg, B, T, U, seed=3, dtype=np.float64)
    spk = bt['speaker'] % cfg['num_speakers'] if cfg.get('use_speaker') else None

    def run():
        m._state_B = None
        m.initial_states(B)
        r = np.random.default_rng(9)
        for nm in ('last_h1', 'last_h2', 'last_h3', 'last_w'):
            getattr(m, nm)[:] = r.standard_normal(getattr(m, nm).shape) * 0.5
        m.last_k[:] = np.abs(r.standard_normal(m.last_k.shape))
        return m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], spk,
                              start_flag, B, feedback_noise=bt['feedback_noise'], noise_level=0.3)[0]
    run()
    g = m.backward()
    rng = np.random.default_rng(5)
    worst = 0.0
    for n, p in m.params.items():
        for _ in range(n_probe):
            idx = tuple(rng.integers(0, s) for s in p.shape)
            old = p[idx]
            p[idx] = old + 1e-6; cp = run()
            p[idx] = old - 1e-6; cm = run()
            p[idx] = old
            fd = (cp - cm) / 2e-6
            if abs(fd) < 1e-7 and abs(g[n][idx]) < 1e-7:
                continue
            worst = max(worst, abs(fd - g[n][idx]) / (abs(fd) + abs(g[n][idx])))
    return worst


@pytest.mark.parametrize('over,axis,sf', [
    (dict(), 0, 1.0),
    (dict(), 1, 1.0),
    (dict(which_cost='GMM', weak_feedback=True), 0, 1.0),
    (dict(attention_type='softmax', full_feedback=True, feedback_noise_level=0.3), 0, 1.0),
    (dict(layer_norm=True, weak_feedback=True, use_speaker=True), 0, 1.0),
    (dict(layer_norm=True, full_feedback=True, use_speaker=True, which_cost='GMM'), 0, 0.0),
    (dict(use_speaker=True), 0, 0.0),
])
def test_backward_matches_finite_differences(over, axis, sf):
    assert _fd_check(dict(SMALL, **over), axis=axis, start_flag=sf) < 2e-3


def test_float32_matches_float64():
    cfg = dict(util.TINY, weak_feedback=True, attention_alignment=0.4)
    o32 = util.make_oracle(cfg, gain=0.5)
    o64 = util.make_oracle(cfg, gain=0.5, dtype=np.float64)
    bt = util.make_batch(cfg, 4, 10, 12)
    a = o32.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, 4)
    b = o64.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, 4)
    assert abs(a[0] - b[0]) / abs(b[0]) < 1e-5
    assert util.rel_err(a[2][0], b[2][0]) < 1e-4
    ga, gb = o32.backward(), o64.backward()
    for n in ga:
        assert util.rel_err(ga[n], gb[n]) < 1e-3, n


@pytest.mark.parametrize('which', ['MSE', 'GMM'])
def test_sampler_consistent_with_teacher_forcing(which):
    """Feeding the sampler's own output back through compute_cost must reproduce its k, w, phi
    (same step function, model.py:651-724 vs 882-1036) when the sampling knobs are neutral."""
    cfg = dict(util.TINY, which_cost=which, weak_feedback=True, attention_alignment=0.4)
    orc = util.make_oracle(cfg, gain=0.5)
    B, T, U = 3, 7, 9
    bt = util.make_batch(cfg, B, T, U, seed=2)
    x, k, w, pi, phi, pia = orc.sample_model(bt['labels'], bt['labels_mask'], None, None, B, T,
                                              gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'])
    feats = np.concatenate([np.zeros((1, B, 63), np.float32), x], 0)
    _, _, av, _ = orc.compute_cost(feats, np.ones((T + 1, B), np.float32), bt['labels'], bt['labels_mask'],
                                   None, 1.0, B, gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'])
    assert util.rel_err(av[1], k) < 1e-5 and util.rel_err(av[2], w) < 1e-5 and util.rel_err(av[4], phi) < 1e-5
    assert util.rel_err(av[0], x) < 1e-4


@pytest.mark.parametrize('name', ['tiny_mse_graves', 'tiny_gmm_softmax_spk'])
def test_golden_vectors(name):
    """Frozen vectors (tests/golden/make_golden.py): the oracle must keep reproducing them."""
    from tests.golden.make_golden import CASES, B, T, U
    z = np.load(os.path.join(GOLD, name + '.npz'))
    cfg = CASES[name]['cfg']
    orc = O.OracleParrot(**cfg)
    orc.set_params({n: z['param:' + n] for n in orc.shapes})
    for seg, sf in enumerate((1.0, 0.0)):
        p = 'seg%d:' % seg
        spk = z[p + 'in:speaker'] if cfg.get('use_speaker') else None
        cost, updates, av, _ = orc.compute_cost(
            z[p + 'in:features'], z[p + 'in:features_mask'], z[p + 'in:labels'], z[p + 'in:labels_mask'], spk,
            sf, B, gmm_unis=z[p + 'in:gmm_unis'], gmm_normals=z[p + 'in:gmm_normals'])
        assert abs(cost - z[p + 'cost']) / abs(z[p + 'cost']) < 1e-5
        for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
            assert util.rel_err(v, z[p + 'out:' + nm]) < 1e-4, nm
        assert (av[4].argmax(-1) == z[p + 'argmax_phi']).all()
        grads = orc.backward()
        sig = np.array([[g.sum(dtype=np.float64), np.sqrt((g.astype(np.float64) ** 2).sum())]
                        for g in grads.values()])
        assert np.allclose(sig[:, 1], z[p + 'grad_sig'][:, 1], rtol=1e-3, atol=1e-7)


def test_param_inventory_counts():
    """25.2 M (MSE) / 27.7 M (GMM) parameters at the base configuration."""
    base = dict(O.DEFAULTS, encoder_type='bidirectional', encoded_input_dim=256)
    n_mse = sum(int(np.prod(s)) for s in O.param_shapes(base).values())
    n_gmm = sum(int(np.prod(s)) for s in O.param_shapes(dict(base, which_cost='GMM')).values())
    assert abs(n_mse - 25.2e6) < 0.1e6 and abs(n_gmm - 27.7e6) < 0.2e6


def test_logsumexp_cost_gmm_against_direct_formula():
    rng = np.random.default_rng(0)
    N, D, k = 6, 4, 3
    y = rng.standard_normal((N, D)); mu = rng.standard_normal((N, D * k))
    sig = np.exp(rng.standard_normal((N, D * k)) * 0.3); w = O.softmax(rng.standard_normal((N, k)))
    nll = O.cost_gmm(y, mu, sig, w)
    direct = np.zeros(N)
    for n in range(N):
        comp = 0.0
        for j in range(k):
            lp = 0.0
            for d in range(D):
                m, s = mu[n, d * k + j], sig[n, d * k + j]
                lp += -0.5 * ((y[n, d] - m) ** 2 / s ** 2 + 2 * np.log(s) + np.log(2 * np.pi))
            comp += w[n, j] * np.exp(lp)
        direct[n] = -np.log(comp)
    assert np.allclose(nll, direct, rtol=1e-10)


def test_multinomial_from_uniform_semantics():
    p = np.array([[0.2, 0.3, 0.5], [0.2, 0.3, 0.5], [0.2, 0.3, 0.5]], np.float32)
    assert list(O.multinomial_from_uniform(p, np.array([0.1, 0.2, 0.99], np.float32))) == [0, 1, 2]


def test_stop_heuristic_reference_behaviour():
    """sample.py:147-163: first frame where phi[t, len] exceeds every phi[t, :len-1]; +40, clamped."""
    T, U, L = 100, 10, 6
    phi = np.zeros((T, U), np.float32)
    for t in range(T):
        phi[t, min(U - 1, t // 10)] = 1.0
    assert O.stop_heuristic(phi, L, T) == min(T, 60 + 40)
    assert O.stop_heuristic(np.ones((T, U), np.float32), L, T) == T


def test_mma_emulator_ordering():
    """The precision experiment behind the bf16x3 choice: error(bf16) >> error(tf32) >> error(bf16x3)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, 512)).astype(np.float32); W = rng.standard_normal((512, 128)).astype(np.float32)
    ref = x.astype(np.float64) @ W.astype(np.float64)
    e = {m: np.abs(O.emulate_mma(x, W, m) - ref).max() / np.abs(ref).max() for m in ('bf16', 'tf32', 'fp16', 'bf16x3')}
    assert e['bf16'] > 5 * e['tf32'] and e['tf32'] > 5 * e['bf16x3'] and e['bf16x3'] < 2e-5


REF_MODEL_CASES = ['mse_weak', 'gmm_full_spk_softmax', 'layer_norm_noise', 'layer_norm_gmm_full_spk']


def test_free_functions_match_the_reference_source():
    """tests/golden/ref_functions.npz holds outputs of the reference's own _simple_norm / logsumexp / cost_gmm /
    sample_gmm (model.py:24-118), executed unmodified on a numpy stand-in (make_ref_function_fixtures.py)."""
    z = np.load(os.path.join(GOLD, 'ref_functions.npz'))
    for k in ('norm_a', 'norm_b'):
        assert np.abs(O.simple_norm(z[k + ':x']) - z[k + ':y']).max() < 1e-12
    assert np.abs(O.logsumexp(z['lse:x'], axis=-1) - z['lse:y']).max() < 1e-12
    nll = O.cost_gmm(z['gmm:y'], z['gmm:mu'], z['gmm:sig'], z['gmm:weight'])
    assert nll.shape == z['gmm:nll'].shape and np.abs(nll - z['gmm:nll']).max() < 1e-12
    x = O.sample_gmm(z['samp:mu'], z['samp:sigma'], z['samp:weight'], z['samp:unis'], z['samp:normals'])
    assert np.abs(x - z['samp:x']).max() < 1e-12


@pytest.mark.parametrize('name', REF_MODEL_CASES)
def test_oracle_matches_the_reference_parrot_class(name):
    """tests/golden/ref_model_*.npz: outputs of the reference's own Parrot.compute_cost (two TBPTT segments) and
    Parrot.sample_model_fun, executed unmodified on the eager Theano/Blocks stand-in tests/golden/ref_shim.py
    (make_ref_model_fixtures.py).  The oracle must reproduce them to rounding in float64 and to 1e-4 in float32:
    this pins its wiring to the reference's code (only the brick arithmetic of Blocks stays restated)."""
    from tests.golden.make_ref_model_fixtures import CASES, B, T, SAMP
    z = np.load(os.path.join(GOLD, 'ref_model_%s.npz' % name))
    cfg = dict(util.TINY, **CASES[name])
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 2e-4)):
        orc = O.OracleParrot(dtype=dtype, **cfg)
        orc.set_params({n: z['param:' + n].astype(dtype) for n in orc.shapes})
        for seg, sf in enumerate((1.0, 0.0)):
            p = 'seg%d:' % seg
            spk = z[p + 'in:speaker'] if cfg.get('use_speaker') else None
            cost, updates, av, _ = orc.compute_cost(
                z[p + 'in:features'].astype(dtype), z[p + 'in:features_mask'].astype(dtype), z[p + 'in:labels'],
                z[p + 'in:labels_mask'].astype(dtype), spk, sf, B, gmm_unis=z[p + 'in:gmm_unis'],
                gmm_normals=z[p + 'in:gmm_normals'], feedback_noise=z[p + 'in:feedback_noise'].astype(dtype),
                noise_level=cfg.get('feedback_noise_level'))
            assert abs(cost - z[p + 'cost']) / abs(z[p + 'cost']) < tol
            for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
                assert util.rel_err(v, z[p + 'out:' + nm]) < tol, (nm, seg)
            for nm, v in updates:
                assert util.rel_err(v, z[p + 'update:' + nm]) < tol, (nm, seg)
            if dtype == np.float64:
                assert (av[4].argmax(-1) == z[p + 'out:phi'].argmax(-1)).all()
        so = O.OracleParrot(dtype=dtype, **dict(cfg, **SAMP))
        so.set_params({n: z['param:' + n].astype(dtype) for n in so.shapes})
        spk = z['samp:in:speaker'] if cfg.get('use_speaker') else None
        res = so.sample_model(z['samp:in:labels'], z['samp:in:labels_mask'].astype(dtype), None, spk, B, T,
                              gmm_unis=z['samp:in:gmm_unis'], gmm_normals=z['samp:in:gmm_normals'])
        for nm, v in zip(['x', 'k', 'w', 'pi', 'phi', 'pi_att'], res):
            assert util.rel_err(v, z['samp:out:' + nm]) < (tol if dtype == np.float64 else 2e-3), nm


@pytest.mark.parametrize('name', REF_MODEL_CASES)
def test_oracle_backward_matches_finite_differences_of_the_reference_cost(name):
    """ref_model_*.npz also hold central finite differences of the REFERENCE's compute_cost (executed on the
    stand-in, float64) at two entries of every parameter tensor: the oracle's hand-derived BPTT is the derivative of
    the reference's forward code, tensor by tensor."""
    from tests.golden.make_ref_model_fixtures import CASES, B
    z = np.load(os.path.join(GOLD, 'ref_model_%s.npz' % name))
    cfg = dict(util.TINY, **CASES[name])
    orc = O.OracleParrot(dtype=np.float64, **cfg)
    orc.set_params({n: z['param:' + n] for n in orc.shapes})
    p = 'seg0:'
    spk = z[p + 'in:speaker'] if cfg.get('use_speaker') else None
    orc.compute_cost(z[p + 'in:features'], z[p + 'in:features_mask'], z[p + 'in:labels'], z[p + 'in:labels_mask'], spk,
                     1.0, B, gmm_unis=z[p + 'in:gmm_unis'], gmm_normals=z[p + 'in:gmm_normals'],
                     feedback_noise=z[p + 'in:feedback_noise'], noise_level=cfg.get('feedback_noise_level'))
    g = orc.backward()
    checked = 0
    for n in orc.shapes:
        idx, fd = z['fd:idx:' + n], z['fd:val:' + n]
        scale = max(np.abs(g[n]).max(), 1e-12)
        assert np.abs(g[n].reshape(-1)[idx] - fd).max() / scale < 2e-5, n
        checked += int((np.abs(fd) > 0).sum())
    assert checked > 100
