"""Raw-audio quantisation of the reference (quantize.py:1-98): per-row min/max normalisation, linear and
mu-law (mu = 255) quantisation and its inverse.  Host-side numpy (data preparation for the sampleRNN
vocoder); kept because its docstring holds the only known-answers pinned anywhere in the
reference tree (quantize.py:55-63), which tests/test_quantize.py checks."""
import numpy as np


def normalize(data):
    """quantize.py:14-18: each row to [0, 1] (returns a float64 copy instead of mutating in place)."""
    data = np.array(data, dtype=np.float64)
    data -= data.min(axis=1)[:, None]
    data /= data.max(axis=1)[:, None]
    return data


def linear_quantize(data, q_levels):
    """quantize.py:20-37: floats in (0, 1) -> ints in [0, q_levels - 1]."""
    eps = np.float64(1e-5)
    data = np.array(data, dtype=np.float64) * (q_levels - eps)
    data += eps / 2
    return data.astype('int32')


def linear2mu(x, mu=255):
    """quantize.py:44-66: x in [-1, 1] -> int16 in [0, mu]."""
    x = np.asarray(x)
    x_mu = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    return ((x_mu + 1) / 2 * mu).astype('int16')


def mu2linear(x, mu=255):
    """quantize.py:68-78: inverse of linear2mu (not exact: 8-bit compression)."""
    mu = float(mu)
    x = np.asarray(x).astype('float32')
    y = 2. * (x - (mu + 1.) / 2.) / (mu + 1.)
    return np.sign(y) * (1. / mu) * ((1. + mu) ** np.abs(y) - 1.)


def batch_quantize(data, q_levels, q_type):
    """quantize.py:83-98."""
    data = normalize(np.asarray(data, dtype=np.float64))
    if q_type == 'linear':
        return linear_quantize(data, q_levels)
    if q_type == 'mu-law':
        return linear2mu(2. * data - 1.)
    if q_type == 'a-law':
        raise NotImplementedError      # quantize.py:39-43
    raise ValueError(q_type)
