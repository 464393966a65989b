"""Seeded synthetic batches of the Parrot hot path (SURVEY.md 8d "Synthetic inputs"): the same bytes feed the
CUDA path (bench.py, tests) and the CPU oracle.  Product code: no test or oracle import."""
import numpy as np


def make_batch(cfg, B, T, U, seed=0, ragged=True, dtype=np.float32):
    """features (T+1,B,D) ~ N(0,1); masks with lengths ~ U{0.6..1}; labels ~ U{0..num_characters}."""
    rng = np.random.default_rng(seed)
    D = cfg.get('output_dim', 63)
    feats = rng.standard_normal((T + 1, B, D)).astype(dtype)
    fm = np.ones((T + 1, B), dtype)
    lm = np.ones((B, U), dtype)
    if ragged:
        for b in range(B):
            fl = int(rng.integers(int(0.6 * (T + 1)), T + 2))
            fm[fl:, b] = 0
            ul = int(rng.integers(max(2, int(0.6 * U)), U + 1))
            lm[b, ul:] = 0
    labels = rng.integers(0, cfg.get('num_characters', 43), (B, U)).astype(np.int32)
    spk = rng.integers(0, cfg.get('num_speakers', 21), (B, 1)).astype(np.int32)
    return dict(features=feats, features_mask=fm, labels=labels, labels_mask=lm, speaker=spk,
                feedback_noise=rng.standard_normal((T, B, D)).astype(dtype),
                gmm_unis=rng.random((T, B)).astype(dtype),
                gmm_normals=rng.standard_normal((T, B, D)).astype(dtype))
