"""Shared helpers of the parity tests: matching oracle / device models and seeded synthetic batches
(SURVEY.md 8d 'Synthetic inputs')."""
import numpy as np

from oracle.parrot_oracle import OracleParrot

TINY = dict(input_dim=24, output_dim=63, rnn_h_dim=64, readouts_dim=64, num_characters=43,
            attention_size=10, encoder_dim=64, k_gmm=20, num_speakers=5, speaker_dim=16,
            encoder_type='bidirectional')


from parrot_b200.synthetic import make_batch  # noqa: E402,F401  (shared with bench.py; product code)


def make_oracle(cfg, seed=0, gain=None, dtype=np.float32, encoder_time_axis=0, bias_std=0.1):
    m = OracleParrot(dtype=dtype, encoder_time_axis=encoder_time_axis, **cfg)
    m.initialize(np.random.default_rng(seed), gain=gain)
    if bias_std:
        rng = np.random.default_rng(seed + 1000)
        for n in m.params:
            if n.endswith('.b') or n.endswith('initial_state') or n.endswith('initial_w'):
                m.params[n] = (rng.standard_normal(m.params[n].shape) * bias_std).astype(dtype)
    return m


def make_device_model(cfg, oracle, gemm_impl='tcgen05', encoder_time_axis=0):
    from parrot_b200.model import Parrot
    m = Parrot(gemm_impl=gemm_impl, encoder_time_axis=encoder_time_axis, **cfg)
    m.initialize()
    m.set_parameter_values({n: np.asarray(v, np.float32) for n, v in oracle.params.items()})
    return m


def argmax_parity(phi_dev, phi_ref, label='', covered_gap=1e-4):
    """The alignment-argmax gate: the device argmax must EQUAL the reference argmax on
    every frame whose reference top-2 gap exceeds ``covered_gap`` relative (and whose phi has not underflowed); the
    covered fraction must be >= 0.9 (printed); frames with a smaller non-zero gap are counted and may differ on fewer
    than 0.1 % of all frames (a tie finer than the GEMM precision cannot be reproduced by arithmetic that is not
    bit-identical to the reference's)."""
    p = np.asarray(phi_ref, np.float64)
    am_ref = p.argmax(-1)
    am_dev = np.asarray(phi_dev).argmax(-1)
    srt = np.sort(p, axis=-1)
    top, second = srt[..., -1], srt[..., -2]
    gap = np.where(top > 0, (top - second) / np.maximum(top, 1e-300), 0.0)
    covered = (gap > covered_gap) & (top > 1e-30)
    near = (gap > 0) & ~covered
    mism_cov = int((am_dev != am_ref)[covered].sum())
    mism_near = int((am_dev != am_ref)[near].sum())
    print('argmax phi %s: %d frames, covered %.4f (%d mismatches), near-ties %d (%d mismatches)'
          % (label, am_ref.size, covered.mean(), mism_cov, int(near.sum()), mism_near))
    assert covered.mean() >= 0.9, covered.mean()
    assert mism_cov == 0, mism_cov
    assert mism_near <= max(1, int(1e-3 * am_ref.size)), mism_near


def rel_err_elementwise(a, b, floor=1e-2):
    """Largest element-wise relative error over the entries of ``b`` above ``floor`` * max|b| (``rel_err`` below is a
    max-norm gate: max|a-b| / max|b|)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    big = np.abs(b) > floor * np.abs(b).max()
    return float((np.abs(a - b)[big] / np.abs(b)[big]).max()) if big.any() else 0.0


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def stable_argmax_mask(phi32, phi64, margin=1e-4):
    """Frames where the oracle's own argmax is unambiguous: the top-2 gap of phi (float64) exceeds
    ``margin`` relative, and phi is not underflowed to ~0.  Bit-exact argmax is demanded there."""
    p = np.asarray(phi64, np.float64)
    srt = np.sort(p, axis=-1)
    top, second = srt[..., -1], srt[..., -2]
    ok = (top > 1e-30) & ((top - second) > margin * top)
    ok &= (np.asarray(phi32).argmax(-1) == p.argmax(-1))
    return ok
