"""Golden vectors from the reference's OWN free functions (model.py:24-118), executed here.

model.py cannot be imported (Python 2, Theano, Blocks), but `_simple_norm`, `logsumexp`, `predict`, `cost_gmm` and
`sample_gmm` are pure tensor algebra.  This script reads their source text from the read-only reference mount,
executes it UNMODIFIED (nothing is copied into the repository) against a small numpy stand-in for the handful of
`theano.tensor` operations they use, and stores seeded inputs and the outputs in tests/golden/ref_functions.npz.
tests/test_oracle.py then holds the oracle's restatements to these vectors: that pins the D x K memory layout of
the GMM head, the sign / constant conventions of the NLL and the normalisation formula to the reference's code.

Stand-in semantics: `x.std(-1)` population std (Theano's default, like numpy's); `theano_rng.multinomial(pvals)`
one-hot of the first index whose cumulative probability exceeds an injected uniform (MultinomialFromUniform);
`theano_rng.normal` returns injected draws; Python 2 integer division in `dim = mu.shape[-1] / k`.

    python tests/golden/make_ref_function_fixtures.py        # in the build container only
"""
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/model.py'
WANTED = ('_simple_norm', 'logsumexp', 'predict', 'cost_gmm', 'sample_gmm')


class ShapeElem(int):
    """One entry of a shape vector that remembers where it came from (for set_subtensor)."""
    def __new__(cls, v, parent=None, index=None):
        o = int.__new__(cls, v)
        o.parent, o.index = parent, index
        return o

    def __truediv__(self, other):        # Python 2 semantics of `mu.shape[-1] / k`
        return int(self) // int(other)


class ShapeVec(tuple):
    def __getitem__(self, i):
        v = tuple.__getitem__(self, i)
        if isinstance(i, slice):
            return ShapeVec(v)
        return ShapeElem(v, self, i % len(self))


class Arr(numpy.ndarray):
    """ndarray whose .shape / .reshape behave like a Theano variable's for the calls in model.py:24-118."""
    @property
    def shape(self):
        return ShapeVec(numpy.ndarray.shape.__get__(self))

    def reshape(self, shape, ndim=None):
        out = numpy.asarray(self).reshape(tuple(int(s) for s in shape)).view(Arr)
        assert ndim is None or out.ndim == ndim
        return out


def A(x):
    return numpy.asarray(x).view(Arr)


class tensor(object):
    """The theano.tensor operations used by the five functions."""
    @staticmethod
    def shape_padright(x, n=1):
        return A(numpy.asarray(x).reshape(numpy.asarray(x).shape + (1,) * n))

    @staticmethod
    def shape_padleft(x, n=1):
        return A(numpy.asarray(x).reshape((1,) * n + numpy.asarray(x).shape))

    max = staticmethod(lambda x, axis=None, keepdims=False: A(numpy.max(numpy.asarray(x), axis=axis, keepdims=keepdims)))
    sum = staticmethod(lambda x, axis=None, keepdims=False: A(numpy.sum(numpy.asarray(x), axis=axis, keepdims=keepdims)))
    log = staticmethod(lambda x: A(numpy.log(numpy.asarray(x))))
    exp = staticmethod(lambda x: A(numpy.exp(numpy.asarray(x))))
    sqr = staticmethod(lambda x: A(numpy.square(numpy.asarray(x))))
    argmax = staticmethod(lambda x, axis=-1: numpy.argmax(numpy.asarray(x), axis=axis))
    arange = staticmethod(lambda n: numpy.arange(int(n)))
    eq = staticmethod(lambda a, b: A(numpy.equal(numpy.asarray(a), numpy.asarray(b))))

    @staticmethod
    def set_subtensor(elem, value):
        v = list(elem.parent)
        v[elem.index] = int(value)
        return tuple(v)


class FakeRng(object):
    """Injected randomness in place of MRG_RandomStreams."""
    def __init__(self, unis, normals):
        self.unis, self.normals = unis, normals

    def multinomial(self, pvals, dtype=None):
        p = numpy.asarray(pvals)
        cdf = numpy.cumsum(p, axis=-1)
        idx = numpy.minimum((cdf <= self.unis[:, None]).sum(-1), p.shape[-1] - 1)
        return A(numpy.eye(p.shape[-1], dtype=p.dtype)[idx])

    def normal(self, size, avg=0., std=1., dtype=None):
        assert tuple(int(s) for s in size) == self.normals.shape
        return A(self.normals)


def load_reference_functions():
    src = open(REF).read().split('\n')
    starts = [i for i, l in enumerate(src) if l and not l[0].isspace() and not l.startswith('#')]
    ns = {'tensor': tensor, 'numpy': numpy}
    for a, b in zip(starts, starts[1:] + [len(src)]):
        if src[a].startswith('def ') and src[a][4:src[a].index('(')] in WANTED:
            exec(compile('\n'.join(src[a:b]), REF, 'exec'), ns)
    assert all(w in ns for w in WANTED)
    return ns


def main():
    f = load_reference_functions()
    rng = numpy.random.default_rng(2024)
    out = {}
    # _simple_norm on the shapes the model uses: (T, B, H) and (1, B, 2H)
    for name, shape in (('norm_a', (3, 4, 16)), ('norm_b', (1, 5, 32))):
        x = (rng.standard_normal(shape) * 2 + 0.5).astype(numpy.float64)
        out[name + ':x'] = x
        out[name + ':y'] = numpy.asarray(f['_simple_norm'](A(x)))
    # cost_gmm: y (T, B, D); mu, sig (T, B, D*K) in the Fork's flat layout; weight (T, B, K)
    T, B, D, K = 3, 4, 5, 6
    y = rng.standard_normal((T, B, D))
    mu = rng.standard_normal((T, B, D * K))
    sig = numpy.exp(rng.standard_normal((T, B, D * K)) * 0.3) + 1e-5
    w = rng.random((T, B, K)) + 0.05
    w = w / w.sum(-1, keepdims=True) + 1e-5
    out.update({'gmm:y': y, 'gmm:mu': mu, 'gmm:sig': sig, 'gmm:weight': w,
                'gmm:nll': numpy.asarray(f['cost_gmm'](A(y), A(mu), A(sig), A(w)))})
    # sample_gmm on (B, D*K) / (B, K) as sample_model_fun calls it (model.py:1025-1033)
    unis = rng.random(B)
    normals = rng.standard_normal((B, D))
    mu2, sig2, w2 = mu[0], sig[0], w[0] / w[0].sum(-1, keepdims=True)
    out.update({'samp:mu': mu2, 'samp:sigma': sig2, 'samp:weight': w2, 'samp:unis': unis, 'samp:normals': normals,
                'samp:x': numpy.asarray(f['sample_gmm'](A(mu2), A(sig2), A(w2), FakeRng(unis, normals)))})
    # logsumexp
    z = rng.standard_normal((7, 9)) * 5
    out.update({'lse:x': z, 'lse:y': numpy.asarray(f['logsumexp'](A(z), axis=-1))})
    numpy.savez_compressed(os.path.join(HERE, 'ref_functions.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
