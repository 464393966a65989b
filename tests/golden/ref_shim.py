"""Eager numpy stand-ins for the few Theano / Blocks names that /root/reference/model.py uses, so that the
reference's OWN Parrot class (model.py:171-1083: RecurrentWithFork, Encoder, Parrot.compute_cost,
Parrot.sample_model_fun) can be executed here, unmodified, on concrete arrays.  Test infrastructure only: used
by tests/golden/make_ref_model_fixtures.py in the build container to generate golden vectors.

What this pins and what it does not: every line of model.py that is executed (wiring of the Forks, attention
window, masks, readouts, cost normalisation, carried-state updates, sampling loop, GMM head) is the reference's
code.  The bricks themselves (Linear, Fork, LookupTable, GatedRecurrent, Bidirectional) live in mila-iqia/blocks,
which is not in the reference tree; they are restated below from the published Blocks semantics, the same
assumption oracle/parrot_oracle.py makes.
"""
import copy

import numpy


# --------------------------------------------------------------------------- arrays
class ShapeElem(int):
    def __new__(cls, v, parent=None, index=None):
        o = int.__new__(cls, v)
        o.parent, o.index = parent, index
        return o

    def __truediv__(self, other):            # Python 2: `dim = mu.shape[-1] / k` (model.py:97)
        return int(self) // int(other)


class ShapeVec(tuple):
    def __getitem__(self, i):
        v = tuple.__getitem__(self, i)
        if isinstance(i, slice):
            return ShapeVec(v)
        return ShapeElem(v, self, i % len(self))


class Arr(numpy.ndarray):
    """ndarray with Theano-variable manners: `+=` rebinds instead of mutating, reshape(shape, ndim=)."""
    @property
    def shape(self):
        return ShapeVec(numpy.ndarray.shape.__get__(self))

    def reshape(self, *shape, **kw):
        if len(shape) == 1 and not isinstance(shape[0], (int, numpy.integer)):
            shape = shape[0]
        out = numpy.asarray(self).reshape(tuple(int(s) for s in shape)).view(Arr)
        assert kw.get('ndim') in (None, out.ndim)
        return out

    def __iadd__(self, other):               # symbolic `x += y` creates a new variable
        return A(numpy.asarray(self) + numpy.asarray(other))

    def dimshuffle(self, *order):
        return A(numpy.asarray(self).transpose(order))


def A(x):
    return numpy.asarray(x).view(Arr)


# --------------------------------------------------------------------------- theano
class _Config(object):
    floatX = 'float64'


class _NNet(object):
    @staticmethod
    def softmax(x):
        x = numpy.asarray(x)
        assert x.ndim == 2
        e = numpy.exp(x - x.max(axis=1, keepdims=True))
        return A(e / e.sum(axis=1, keepdims=True))


class tensor(object):
    nnet = _NNet

    @staticmethod
    def zeros(shape, dtype=None):
        return A(numpy.zeros(tuple(int(s) for s in shape), dtype or _Config.floatX))

    @staticmethod
    def switch(cond, a, b):
        return a if float(cond) != 0.0 else b

    @staticmethod
    def shape_padright(x, n=1):
        x = numpy.asarray(x)
        return A(x.reshape(x.shape + (1,) * n))

    @staticmethod
    def shape_padleft(x, n=1):
        x = numpy.asarray(x)
        return A(x.reshape((1,) * n + x.shape))

    @staticmethod
    def arange(n, dtype=None):
        return A(numpy.arange(int(n), dtype=dtype))

    @staticmethod
    def repeat(x, n, axis):
        return A(numpy.repeat(numpy.asarray(x), int(n), axis))

    @staticmethod
    def set_subtensor(elem, value):
        v = list(elem.parent)
        v[elem.index] = int(value)
        return tuple(v)

    @staticmethod
    def scalar(name=None):
        return None                          # placeholder; the generator assigns the value

    exp = staticmethod(lambda x: A(numpy.exp(numpy.asarray(x))))
    log = staticmethod(lambda x: A(numpy.log(numpy.asarray(x))))
    sqrt = staticmethod(lambda x: A(numpy.sqrt(numpy.asarray(x))))
    sqr = staticmethod(lambda x: A(numpy.square(numpy.asarray(x))))
    sum = staticmethod(lambda x, axis=None, keepdims=False: A(numpy.sum(numpy.asarray(x), axis=axis, keepdims=keepdims)))
    max = staticmethod(lambda x, axis=None, keepdims=False: A(numpy.max(numpy.asarray(x), axis=axis, keepdims=keepdims)))
    argmax = staticmethod(lambda x, axis=-1: numpy.argmax(numpy.asarray(x), axis=axis))
    eq = staticmethod(lambda a, b: A(numpy.equal(numpy.asarray(a), numpy.asarray(b))))
    concatenate = staticmethod(lambda xs, axis=0: A(numpy.concatenate([numpy.asarray(x) for x in xs], axis=axis)))


class _Updates(list):
    """scan's update dictionary; the reference adds a list to it (model.py:822)."""


class theano(object):
    config = _Config
    tensor = tensor

    @staticmethod
    def scan(fn, sequences=(), non_sequences=(), outputs_info=(), go_backwards=False):
        n = int(numpy.asarray(sequences[0]).shape[0])
        state = list(outputs_info)
        outs = [[] for _ in outputs_info]
        order = range(n - 1, -1, -1) if go_backwards else range(n)
        for t in order:
            args = [A(numpy.asarray(s)[t]) for s in sequences] + [s for s in state if s is not None] + \
                list(non_sequences)
            res = fn(*args)
            if not isinstance(res, (tuple, list)):
                res = [res]
            for i, r in enumerate(res):
                outs[i].append(numpy.asarray(r))
                if outputs_info[i] is not None:
                    state[i] = A(r)
        stacked = [A(numpy.stack(o, 0)) for o in outs]
        return (stacked if len(stacked) > 1 else stacked[0]), _Updates()


def function(*args, **kwargs):
    raise RuntimeError('compile step is not emulated: call compute_cost / sample_model_fun directly')


# --------------------------------------------------------------------------- blocks
def lazy(allocation=None, initialization=None):
    return lambda f: f


class _Bound(object):
    """A brick's application bound to an instance (blocks BoundApplication): callable, knows its brick."""
    def __init__(self, app, brick):
        self.app, self.brick = app, brick

    def __call__(self, *args, **kwargs):
        as_list = kwargs.pop('as_list', False)
        as_dict = kwargs.pop('as_dict', False)
        out = self.app.fn(self.brick, *args, **kwargs)
        if as_dict:
            return out if isinstance(out, dict) else dict(zip(self.outputs, out))
        if isinstance(out, dict):
            out = [out[k] for k in self.outputs]
        if as_list:
            return list(out) if isinstance(out, (list, tuple)) else [out]
        if isinstance(out, (list, tuple)) and len(out) == 1:
            return out[0]
        return out

    def __getattr__(self, name):
        if name in ('app', 'brick') or name.startswith('__'):
            raise AttributeError(name)
        if name in self.app.props:
            return self.app.props[name](self.brick)
        if name in self.app.attrs:
            return self.app.attrs[name]
        raise AttributeError(name)


class application(object):
    """blocks.bricks.base.application, used bare, with keyword arguments, and with .property()."""
    def __init__(self, fn=None, **attrs):
        self.fn, self.attrs, self.props = fn, attrs, {}

    def __call__(self, fn):                  # @application(inputs=[...])
        self.fn = fn
        return self

    def __get__(self, brick, owner):
        return self if brick is None else _Bound(self, brick)

    def property(self, name):
        def deco(f):
            self.props[name] = f
            return f
        return deco


class Brick(object):
    def __init__(self, name=None, **kwargs):
        self.name = name or self.__class__.__name__.lower()
        if not hasattr(self, 'children'):
            self.children = []
        self.params = {}

    # ---- allocation (blocks: push_allocation_config -> allocate, top down)
    def allocate_all(self):
        if hasattr(self, '_push_allocation_config'):
            self._push_allocation_config()
        if hasattr(self, '_allocate'):
            self._allocate()
        for c in self.children:
            c.allocate_all()

    def named_parameters(self, prefix=''):
        path = prefix + '/' + self.name
        out = {}
        for k, v in self.params.items():
            out[path + '.' + k] = (self, k)
        for k in ('initial_w',):              # Parrot._allocate keeps it as an attribute (model.py:506-510)
            if k in self.__dict__ and isinstance(self.__dict__[k], numpy.ndarray):
                out[path + '.' + k] = (self, '@' + k)
        for c in self.children:
            out.update(c.named_parameters(path))
        return out


class Initializable(Brick):
    pass


class Random(Brick):
    theano_rng = None


class Linear(Initializable):
    def __init__(self, input_dim=None, output_dim=None, **kwargs):
        super(Linear, self).__init__(**kwargs)
        self.input_dim, self.output_dim = input_dim, output_dim

    def _allocate(self):
        self.params = {'W': A(numpy.zeros((self.input_dim, self.output_dim))), 'b': A(numpy.zeros(self.output_dim))}

    @application
    def apply(self, x):
        return A(numpy.dot(numpy.asarray(x), self.params['W']) + self.params['b'])


class Fork(Initializable):
    """blocks.bricks.parallel.Fork: one Linear per output, named fork_<output>."""
    def __init__(self, output_names, input_dim=None, output_dims=None, prototype=None, **kwargs):
        super(Fork, self).__init__(**kwargs)
        self.output_names, self.input_dim, self.output_dims = list(output_names), input_dim, output_dims
        self.children = [Linear(name='fork_' + n) for n in self.output_names]

    def _push_allocation_config(self):
        for c, d in zip(self.children, self.output_dims):
            c.input_dim, c.output_dim = self.input_dim, d

    @application
    def apply(self, x):
        return [c.apply(x) for c in self.children]

    @apply.property('outputs')
    def apply_outputs(self):
        return self.output_names


class LookupTable(Initializable):
    def __init__(self, length, dim, **kwargs):
        super(LookupTable, self).__init__(**kwargs)
        self.length, self.dim = length, dim

    def _allocate(self):
        self.params = {'W': A(numpy.zeros((self.length, self.dim)))}

    @application
    def apply(self, indices):
        return A(self.params['W'][numpy.asarray(indices)])


def _sigmoid(x):
    return 1.0 / (1.0 + numpy.exp(-x))


class GatedRecurrent(Initializable):
    """blocks.bricks.recurrent.GatedRecurrent (Blocks >= 0.1): gates packed [update | reset];
    next = tanh((s * reset) . W_state + inputs) * update + s * (1 - update)."""
    def __init__(self, dim, **kwargs):
        super(GatedRecurrent, self).__init__(**kwargs)
        self.dim = dim

    def _allocate(self):
        d = self.dim
        self.params = {'state_to_state': A(numpy.zeros((d, d))), 'state_to_gates': A(numpy.zeros((d, 2 * d))),
                       'initial_state': A(numpy.zeros(d))}

    def get_dim(self, name):
        return {'inputs': self.dim, 'states': self.dim, 'gate_inputs': 2 * self.dim, 'mask': 0}[name]

    def initial_states(self, batch_size, *args, **kwargs):
        return A(numpy.repeat(self.params['initial_state'][None, :], int(batch_size), 0))

    def _step(self, inputs, gate_inputs, states, mask=None):
        d = self.dim
        s = numpy.asarray(states)
        g = _sigmoid(s.dot(self.params['state_to_gates']) + numpy.asarray(gate_inputs))
        update, reset = g[:, :d], g[:, d:]
        nxt = numpy.tanh((s * reset).dot(self.params['state_to_state']) + numpy.asarray(inputs))
        nxt = nxt * update + s * (1 - update)
        if mask is not None:
            m = numpy.asarray(mask)[:, None]
            nxt = m * nxt + (1 - m) * s
        return A(nxt)

    @application(sequences=['mask', 'inputs', 'gate_inputs'], states=['states'], outputs=['states'], contexts=[])
    def apply(self, inputs, gate_inputs, states=None, mask=None, iterate=True, reverse=False):
        if not iterate:
            return self._step(inputs, gate_inputs, states, mask)
        inputs, gate_inputs = numpy.asarray(inputs), numpy.asarray(gate_inputs)
        n = inputs.shape[0]
        s = self.initial_states(inputs.shape[1])
        outs = [None] * n
        for t in (range(n - 1, -1, -1) if reverse else range(n)):
            s = self._step(inputs[t], gate_inputs[t], s, None if mask is None else numpy.asarray(mask)[t])
            outs[t] = numpy.asarray(s)
        # recurrent(reverse=True) returns the outputs in scan order; Bidirectional flips them back (x[::-1])
        seq = outs[::-1] if reverse else outs
        return A(numpy.stack(seq, 0))


class Bidirectional(Initializable):
    """blocks.bricks.recurrent.Bidirectional: two copies (forward, backward), outputs concatenated on axis 2."""
    def __init__(self, prototype, **kwargs):
        super(Bidirectional, self).__init__(**kwargs)
        self.children = [copy.deepcopy(prototype) for _ in range(2)]
        self.children[0].name, self.children[1].name = 'forward', 'backward'

    @application
    def apply(self, *args, **kwargs):
        fwd = self.children[0].apply(*args, as_list=True, **kwargs)
        bwd = [A(numpy.asarray(x)[::-1]) for x in self.children[1].apply(*args, reverse=True, as_list=True, **kwargs)]
        return [tensor.concatenate([f, b], axis=2) for f, b in zip(fwd, bwd)]


INITIAL_STATE = PARAMETER = object()


def add_role(var, role):
    pass


def shared_floatx_zeros(shape, name=None, **kwargs):
    return A(numpy.zeros(tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))))


def dict_union(*dicts, **kwargs):
    out = {}
    for d in dicts:
        out.update(d)
    out.update(kwargs)
    return out


class FakeRng(object):
    """Injected draws in place of MRG_RandomStreams: `multinomial` = MultinomialFromUniform on the next block of
    uniforms, `normal` = the next block of standard normals (one block per call pair)."""
    def __init__(self, unis=None, normals=None, feedback_noise=None):
        self.unis, self.normals, self.feedback_noise = unis, normals, feedback_noise
        self.i_u = self.i_n = 0

    def multinomial(self, pvals, dtype=None):
        p = numpy.asarray(pvals)
        u = numpy.asarray(self.unis[self.i_u]).reshape(-1)
        self.i_u += 1
        cdf = numpy.cumsum(p, axis=-1)
        idx = numpy.minimum((cdf <= u[:, None]).sum(-1), p.shape[-1] - 1)
        return A(numpy.eye(p.shape[-1], dtype=p.dtype)[idx])

    def normal(self, size, avg=0., std=1., dtype=None):
        size = tuple(int(s) for s in size)
        if self.feedback_noise is not None and len(size) == 3:      # model.py:574-577 feedback noise
            assert self.feedback_noise.shape == size
            return A(self.feedback_noise)
        x = numpy.asarray(self.normals[self.i_n]).reshape(size)
        self.i_n += 1
        return A(x)


def namespace():
    """Globals for executing model.py:24-118 and model.py:171-1083."""
    return dict(numpy=numpy, theano=theano, tensor=tensor, function=function, floatX=_Config.floatX,
                Initializable=Initializable, Linear=Linear, Random=Random, lazy=lazy, application=application,
                LookupTable=LookupTable, Fork=Fork, GatedRecurrent=GatedRecurrent, Bidirectional=Bidirectional,
                add_role=add_role, INITIAL_STATE=INITIAL_STATE, PARAMETER=PARAMETER,
                shared_floatx_zeros=shared_floatx_zeros, dict_union=dict_union, Brick=Brick)
