"""Golden vectors of the numpy oracle at the BENCHMARKED configuration (BASELINE.json configs[1]):
B=64, H=R=1024, E=128 (C=256), U=128, weak feedback, T frames per case below -- MSE and GMM k=20,
init-scale (train.py:30-31, W ~ N(0, 0.01^2)) and "trained-like" gain-0.5 weights (W ~ N(0, (0.5/sqrt(fan_in))^2)).

The oracle takes minutes per case at these sizes, so its outputs are frozen here once and the GPU tests
(tests/test_gpu_base_parity.py) compare the CUDA path with the files: no oracle run on the GPU box.
Inputs and parameters are NOT stored: they are regenerated from the seeds below by tests/util.py
(numpy Generator streams are stable across numpy versions for these calls; a checksum guards it).

    python tests/golden/make_base_fixtures.py [case ...]      # rewrites tests/golden/base_*.npz

Stored per case (small: < 2 MB each):
    cost f64; next_x at SAMPLE_STEPS time steps (all rows, float32) [MSE: predicted frames; GMM: sampled with the fixed
    noise]; per-step checksums of next_x for ALL steps (sum and l2 over (B, D), float64); argmax_phi (T,B) i16 and
    phi_top2_gap (T,B) f16 = (top1 - top2) / top1 of the float32 oracle; k_last / w_last f32; gradient signatures
    (sum, l2, absmax) of every tensor f64 and GRAD_SAMPLES seeded entries of every tensor f32; input / parameter checksums.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

BASE = dict(input_dim=420, output_dim=63, rnn_h_dim=1024, readouts_dim=1024, weak_feedback=True, which_cost='MSE',
            num_characters=43, attention_type='graves', attention_size=10, attention_alignment=0.15,
            encoder_type='bidirectional', encoder_dim=128)
B, U = 64, 128

CASES = {
    # name: (which_cost, gain, T)
    'base_mse_init_T800': ('MSE', None, 800),
    'base_mse_gain_T200': ('MSE', 0.5, 200),
    'base_gmm_init_T200': ('GMM', None, 200),
    'base_gmm_gain_T200': ('GMM', 0.5, 200),
}
GRAD_SAMPLES = 2048


def sample_steps(T):
    """Time steps whose frames are stored in full: dense at the start, geometric in the middle, dense at the end."""
    st = set(range(0, min(T, 4))) | set(range(max(0, T - 8), T))
    t = 4
    while t < T:
        st.add(t)
        t = int(t * 1.5) + 1
    return np.array(sorted(st), np.int32)


def stable_seed(name):
    return sum((i + 1) * ord(c) for i, c in enumerate(name)) % (2 ** 31)


def grad_sample_index(name, size):
    rng = np.random.default_rng(stable_seed(name))   # (hash() is salted per process)
    return rng.integers(0, size, min(size, GRAD_SAMPLES))


def case_setup(name):
    which, gain, T = CASES[name]
    cfg = dict(BASE, which_cost=which)
    return cfg, gain, T


def checksum(a):
    a = np.ascontiguousarray(a)
    return float(np.asarray(a, np.float64).sum()) + float(np.abs(np.asarray(a, np.float64)).sum()) * 1e-3


def run_case(name):
    cfg, gain, T = case_setup(name)
    t0 = time.time()
    orc = util.make_oracle(cfg, gain=gain, bias_std=0.1 if gain else 0)
    bt = util.make_batch(cfg, B, T, U, seed=11)
    cost, updates, av, _ = orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'],
                                            None, 1.0, B, gmm_unis=bt['gmm_unis'], gmm_normals=bt['gmm_normals'])
    t1 = time.time()
    grads = orc.backward()
    t2 = time.time()
    phi = np.asarray(av[4], np.float32)
    srt = np.sort(phi, axis=-1)
    top, second = srt[..., -1], srt[..., -2]
    gap = np.where(top > 0, (top - second) / np.maximum(top, 1e-38), 0.0).astype(np.float32)
    nx = np.asarray(av[0], np.float64)
    steps = sample_steps(T)
    out = {
        'T': np.int32(T), 'cost': np.float64(cost),
        'sample_steps': steps,
        'next_x_samples': av[0][steps].astype(np.float32),
        'next_x_step_sum': nx.sum(axis=(1, 2)),
        'next_x_step_l2': np.sqrt((nx ** 2).sum(axis=(1, 2))),
        'next_x_absmax': np.float64(np.abs(av[0]).max()),
        'argmax_phi': phi.argmax(-1).astype(np.int16),
        'phi_top2_gap': gap.astype(np.float16),
        'k_last': av[1][-1].astype(np.float32),
        'w_last': av[2][-1].astype(np.float32),
        'grad_names': np.array(list(grads.keys())),
        'grad_sig': np.array([[g.sum(dtype=np.float64), np.sqrt((g.astype(np.float64) ** 2).sum()),
                               np.abs(g).max()] for g in grads.values()]),
        'check_features': np.float64(checksum(bt['features'])),
        'check_params': np.float64(sum(checksum(v) for v in orc.params.values())),
    }
    for n, g in grads.items():
        out['gsamp:' + n] = g.ravel()[grad_sample_index(n, g.size)].astype(np.float32)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **out)
    print('%s: cost %.6f  fwd %.0fs bwd %.0fs  -> %s (%.1f MB)' % (name, cost, t1 - t0, t2 - t1, path,
                                                                    os.path.getsize(path) / 1e6), flush=True)


if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
