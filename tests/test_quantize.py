"""The only numbers the reference tree itself pins (quantize.py:55-63 docstring): mu-law encode spans 0..255 int16,
decode spans -1 .. 0.9574371 float32."""
import numpy as np

from parrot_b200 import quantize


def test_mu_law_known_answers_from_reference_docstring():
    rng = np.random.default_rng(0)
    samples = rng.standard_normal((1, 16000))
    norm = quantize.normalize(samples)
    enc = quantize.linear2mu(2. * norm - 1.)
    assert enc.min() == 0 and enc.max() == 255 and enc.dtype == np.int16
    dec = quantize.mu2linear(enc)
    assert dec.dtype == np.float32 and dec.min() == -1.0
    assert abs(float(dec.max()) - 0.9574371) < 1e-6


def test_batch_quantize_ranges_and_roundtrip_error():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 4000))
    lin = quantize.batch_quantize(x, 256, 'linear')
    assert lin.min() == 0 and lin.max() == 255 and lin.dtype == np.int32
    mu = quantize.batch_quantize(x, 256, 'mu-law')
    assert mu.min() == 0 and mu.max() == 255
    back = quantize.mu2linear(mu)
    ref = 2. * quantize.normalize(x) - 1.
    assert np.abs(back - ref).max() < 0.09          # 8-bit companding, truncating encoder + decode offset (max 0.957)
    assert (np.diff(quantize.mu2linear(np.arange(256, dtype=np.int16))) > 0).all()   # monotone decode
