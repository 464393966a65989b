"""Golden vectors from the reference's OWN Parrot class, executed here on an eager numpy stand-in.

/root/reference/model.py (Python 2 + Theano + Blocks) cannot be imported, but the code of RecurrentWithFork,
Encoder and Parrot (model.py:171-1083) is ordinary Python over a small set of Theano / Blocks names.  This script
reads that source from the read-only reference mount, executes it UNMODIFIED (nothing is copied into the
repository) against tests/golden/ref_shim.py, loads seeded parameters, and runs

  * Parrot.compute_cost on two consecutive TBPTT segments (start_flag 1 then 0, carried state handed over the way
    the compiled function's updates do), and
  * Parrot.sample_model_fun (free-running generation),

in float64 with the random draws injected, plus central finite differences of the reference cost (segment 0) at
two entries of every parameter tensor.  Inputs, parameters and every output go to
tests/golden/ref_model_<case>.npz; tests/test_oracle.py holds oracle/parrot_oracle.py to them.  This pins the
oracle's wiring (which Fork feeds what, attention window, masks, readouts, cost, updates, sampler) to the
reference's code; the brick arithmetic inside the stand-in is the published Blocks semantics (see ref_shim.py).

    python tests/golden/make_ref_model_fixtures.py          # build container only
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import util                      # noqa: E402
from tests.golden import ref_shim as S      # noqa: E402

REF = '/root/reference/model.py'
SKIP = ('SampleRnn',)                       # sampleRNN coupling: out of scope

CASES = {
    'mse_weak': dict(weak_feedback=True, attention_alignment=0.4),
    'gmm_full_spk_softmax': dict(which_cost='GMM', full_feedback=True, use_speaker=True, attention_type='softmax',
                                 attention_alignment=0.4),
    'layer_norm_noise': dict(weak_feedback=True, layer_norm=True, feedback_noise_level=0.3, attention_alignment=0.4),
    'layer_norm_gmm_full_spk': dict(which_cost='GMM', full_feedback=True, use_speaker=True, layer_norm=True,
                                    attention_alignment=0.4),
}
B, T, U = 5, 7, 9
SAMP = dict(sampling_bias=0.5, sharpening_coeff=1.3, timing_coeff=1.2)


def load_reference_classes():
    src = open(REF).read().split('\n')
    starts = [i for i, l in enumerate(src) if l.startswith('def ') or l.startswith('class ')]
    ns = S.namespace()
    for a, b in zip(starts, starts[1:] + [len(src)]):
        name = src[a].split()[1].split('(')[0].rstrip(':')
        if name in SKIP:
            continue
        exec(compile('\n' * a + '\n'.join(src[a:b]), REF, 'exec'), ns)     # line numbers as in the reference
    return ns


def build(ns, cfg, params):
    m = ns['Parrot'](name='parrot', **cfg)
    m.allocate_all()
    slots = m.named_parameters()
    assert set(slots) == set(params), (sorted(set(slots) ^ set(params)))
    for n, (brick, key) in slots.items():
        v = S.A(np.asarray(params[n], np.float64))
        if key.startswith('@'):
            setattr(brick, key[1:], v)
        else:
            assert brick.params[key].shape == v.shape, n
            brick.params[key] = v
    return m


def run_case(name, extra, ns):
    cfg = dict(util.TINY, **extra)
    orc = util.make_oracle(cfg, gain=0.5, dtype=np.float64)
    # float32-representable parameters: stored as float32 (half the file), identical on the float32 device path
    orc.params = {n: v.astype(np.float32).astype(np.float64) for n, v in orc.params.items()}
    out = {'param:' + n: v.astype(np.float32) for n, v in orc.params.items()}
    ref_cfg = {k: v for k, v in cfg.items()}
    model = build(ns, ref_cfg, orc.params)
    carried = None
    orig_initial_states = model.initial_states

    def initial_states(batch_size):
        st = list(orig_initial_states(batch_size))
        if carried is not None:                     # shared variables keep the previous call's updates
            for i, v in zip((1, 3, 5, 9, 7), carried):      # last_h1, last_h2, last_h3, last_k, last_w
                st[i] = S.A(v)
        return tuple(st)
    model.initial_states = initial_states

    for seg, sf in enumerate((1.0, 0.0)):
        bt = util.make_batch(cfg, B, T, U, seed=300 + seg, dtype=np.float64)
        p = 'seg%d:' % seg
        for k in ('features', 'features_mask', 'labels', 'labels_mask', 'speaker', 'gmm_unis', 'gmm_normals',
                  'feedback_noise'):
            out[p + 'in:' + k] = bt[k]
        model.theano_rng = S.FakeRng(unis=[bt['gmm_unis'].reshape(-1)], normals=[bt['gmm_normals'].reshape(T * B, -1)],
                                     feedback_noise=bt['feedback_noise'] if cfg.get('feedback_noise_level') else None)
        if cfg.get('feedback_noise_level') is not None:
            model.noise_level_var = cfg['feedback_noise_level']
        spk = S.A(bt['speaker']) if cfg.get('use_speaker') else None
        cost, updates, av, _ = model.compute_cost(
            S.A(bt['features'].copy()), S.A(bt['features_mask']), S.A(bt['labels']), S.A(bt['labels_mask']), spk, sf, B)
        out[p + 'cost'] = np.float64(cost)
        for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
            out[p + 'out:' + nm] = np.asarray(v)
        if seg == 0:
            # derivative of the REFERENCE cost (this very code path) by central differences, two entries per
            # parameter tensor: pins the oracle's hand-derived backward pass to the reference's forward code
            slots = model.named_parameters()
            rng = np.random.default_rng(7)

            def ref_cost():
                model.theano_rng = S.FakeRng(unis=[bt['gmm_unis'].reshape(-1)],
                                             normals=[bt['gmm_normals'].reshape(T * B, -1)],
                                             feedback_noise=bt['feedback_noise'] if cfg.get('feedback_noise_level') else None)
                c, _, _, _ = model.compute_cost(S.A(bt['features'].copy()), S.A(bt['features_mask']), S.A(bt['labels']),
                                                S.A(bt['labels_mask']), spk, sf, B)
                return float(c)
            for n in sorted(slots):
                brick, key = slots[n]
                arr = getattr(brick, key[1:]) if key.startswith('@') else brick.params[key]
                flat = np.asarray(arr).reshape(-1)          # view: perturbing it perturbs the brick's parameter
                idx = rng.choice(flat.size, size=min(2, flat.size), replace=False)
                fd = []
                for i in idx:
                    old = flat[i]
                    h = 1e-6 * max(1.0, abs(old))
                    flat[i] = old + h
                    cp = ref_cost()
                    flat[i] = old - h
                    cm = ref_cost()
                    flat[i] = old
                    fd.append((cp - cm) / (2 * h))
                out['fd:idx:' + n] = idx.astype(np.int64)
                out['fd:val:' + n] = np.asarray(fd)
        carried = [np.asarray(v) for _, v in updates]
        for nm, v in zip(['last_h1', 'last_h2', 'last_h3', 'last_k', 'last_w'], carried):
            out[p + 'update:' + nm] = v
    # free-running sampler with the sampling-time coefficients
    scfg = dict(ref_cfg, **SAMP)
    sampler = build(ns, scfg, orc.params)
    bt = util.make_batch(cfg, B, T, U, seed=400, dtype=np.float64)
    sampler.theano_rng = S.FakeRng(unis=list(bt['gmm_unis']), normals=list(bt['gmm_normals']))
    spk = S.A(bt['speaker']) if cfg.get('use_speaker') else None
    res = sampler.sample_model_fun(S.A(bt['labels']), S.A(bt['labels_mask']), spk, B, T)
    for k in ('labels', 'labels_mask', 'speaker', 'gmm_unis', 'gmm_normals'):
        out['samp:in:' + k] = bt[k]
    for nm, v in zip(['x', 'k', 'w', 'pi', 'phi', 'pi_att'], res[:6]):
        out['samp:out:' + nm] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'ref_model_%s.npz' % name), **out)
    return cfg, out


def main():
    ns = load_reference_classes()
    for name, extra in CASES.items():
        cfg, out = run_case(name, extra, ns)
        print(name, 'cost', float(out['seg0:cost']), float(out['seg1:cost']))


if __name__ == '__main__':
    main()
