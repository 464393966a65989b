"""Generates the frozen golden vectors under tests/golden/ from the numpy oracle.

The reference (Python 2 / Theano / Blocks) cannot be imported in this container and ships no
fixtures of its own ("parity unpinned"), so these vectors pin the ORACLE: any later
change to oracle/parrot_oracle.py that moves a number is caught by tests/test_oracle.py, and the
GPU tests compare the CUDA path with the same files without needing the oracle's RNG to be stable
across numpy versions.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    'tiny_mse_graves': dict(cfg=dict(util.TINY, weak_feedback=True, attention_alignment=0.4), gain=0.5),
    'tiny_gmm_softmax_spk': dict(cfg=dict(util.TINY, which_cost='GMM', attention_type='softmax',
                                          full_feedback=True, use_speaker=True, attention_alignment=0.4),
                                 gain=0.5),
}
B, T, U = 8, 12, 16


def run_case(name, spec):
    cfg = spec['cfg']
    orc = util.make_oracle(cfg, gain=spec['gain'])
    out = {}
    for n, v in orc.params.items():
        out['param:' + n] = v.astype(np.float32)
    for seg, sf in enumerate((1.0, 0.0)):          # two consecutive TBPTT segments pin the state carry
        bt = util.make_batch(cfg, B, T, U, seed=100 + seg)
        spk = bt['speaker'] if cfg.get('use_speaker') else None
        cost, updates, av, _ = orc.compute_cost(
            bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], spk, sf, B,
            gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'])
        grads = orc.backward()
        p = 'seg%d:' % seg
        for k in ('features', 'features_mask', 'labels', 'labels_mask', 'speaker', 'gmm_unis', 'gmm_normals'):
            out[p + 'in:' + k] = bt[k]
        out[p + 'cost'] = np.float32(cost)
        for nm, v in zip(['next_x', 'k', 'w', 'coeff', 'phi', 'pi_att'], av):
            out[p + 'out:' + nm] = v.astype(np.float32)
        out[p + 'argmax_phi'] = av[4].argmax(-1).astype(np.int32)
        for nm, v in updates:
            out[p + 'update:' + nm] = v.astype(np.float32)
        # gradients: full tensors for a few, (sum, l2) signatures for all
        sig = np.array([[g.sum(dtype=np.float64), np.sqrt((g.astype(np.float64) ** 2).sum())]
                        for g in grads.values()])
        out[p + 'grad_sig'] = sig
        for n in ('/parrot/rnn1.state_to_gates', '/parrot/h1_to_att/fork_kappa.W', '/parrot.initial_w',
                  '/parrot/encoder/embed_label.W', '/parrot/rnn3.state_to_state'):
            out[p + 'grad:' + n] = grads[n].astype(np.float32)
    out['grad_names'] = np.array(list(orc.shapes.keys()))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'cost', float(out['seg0:cost']), float(out['seg1:cost']))


if __name__ == '__main__':
    for name, spec in CASES.items():
        run_case(name, spec)
