"""CPU oracle for the Parrot attention-RNN hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference algorithm (sotelo/parrot,
``model.py``).  It is the *checker* for the CUDA path: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product (``parrot_b200``) never
does, and fails loudly when its CUDA library is missing.

PARITY: PINNED TO THE REFERENCE'S SOURCE FOR THE MODEL CODE, UNPINNED FOR THE BLOCKS BRICKS.
The reference is Python-2 / Theano / Blocks; none of those can be imported in the
build container and the reference tree holds no tests, golden vectors or fixtures
for this path.  Two things stand in for them:

  * model.py itself is executed.  tests/golden/make_ref_function_fixtures.py and
    tests/golden/make_ref_model_fixtures.py read the source of the free functions
    (model.py:24-118) and of RecurrentWithFork / Encoder / Parrot
    (model.py:171-1083) from the read-only reference mount and run it, unmodified,
    on an eager numpy stand-in for the Theano / Blocks names it uses
    (tests/golden/ref_shim.py).  Parrot.compute_cost (two TBPTT segments) and
    Parrot.sample_model_fun outputs for four configurations are committed under
    tests/golden/ref_*.npz; this oracle reproduces them to 1e-15 relative in
    float64 (tests/test_oracle.py), and the CUDA path is tested against the same
    files (tests/test_gpu_parity.py).  The fixtures also carry central finite
    differences of the reference cost for every parameter tensor, which the
    hand-derived backward pass below matches.  That pins every executed line of model.py:
    Fork wiring, attention window, masks, readouts, GMM head, cost, carried-state
    updates, sampler.
  * What stays UNPINNED: the arithmetic inside ``GatedRecurrent`` / ``Linear`` /
    ``Fork`` / ``LookupTable`` / ``Bidirectional`` and ``Adam`` / ``StepClipping``.
    It lives in the third-party package ``mila-iqia/blocks``, which the reference
    neither vendors nor version-pins; it is restated (here and in ref_shim.py) from
    the published Blocks source (blocks/bricks/recurrent/simple.py
    ``GatedRecurrent.apply``, blocks/bricks/recurrent/misc.py
    ``Bidirectional.apply``, blocks/algorithms ``Adam`` / ``StepClipping``),
    anchored on the reference's call sites, each cited below as ``model.py:LINE``.

Further self-consistency pins (tests/test_oracle.py):
  * the hand-derived backward pass agrees with central finite differences of the
    forward pass in float64 (every parameter tensor, every option);
  * float32 and float64 evaluations agree to float32 round-off;
  * the free-running sampler reproduces the teacher-forced graph when it is fed
    its own samples;
  * frozen golden vectors under tests/golden/ (tests/golden/make_golden.py)
    detect any later drift, gradients included.

Layout conventions (reference): frame tensors are time-major (T, B, .)
(datasets.py:274-275); text tensors are batch-major (B, U) (model.py:645-649).
All arithmetic is ``dtype`` (float32 = reference ``floatX``; float64 is used by
the tests for gradient checking).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

SQRT_1_2PI = 0.3989422917366028  # model.py:681-682 "the weird number"


# --------------------------------------------------------------------------
# free functions (model.py:24-118)
# --------------------------------------------------------------------------
def simple_norm(x, eps=1e-5):
    """model.py:24-27  (x - mean) / (eps + std), population std, last axis."""
    mu = x.mean(-1, keepdims=True)
    sd = x.std(-1, keepdims=True)
    return (x - mu) / (x.dtype.type(eps) + sd)


def simple_norm_bwd(dy, x, eps=1e-5):
    """Gradient of simple_norm wrt x."""
    n = x.shape[-1]
    mu = x.mean(-1, keepdims=True)
    xc = x - mu
    sd = np.sqrt((xc * xc).mean(-1, keepdims=True))
    s = x.dtype.type(eps) + sd
    dyc = dy - dy.mean(-1, keepdims=True)
    dot = (dy * xc).sum(-1, keepdims=True)
    safe_sd = np.where(sd > 0, sd, 1)
    return dyc / s - xc * dot / (n * safe_sd * s * s)


def apply_norm(x, layer_norm):
    """model.py:30-34."""
    return simple_norm(x) if layer_norm else x


def logsumexp(x, axis=None):
    """model.py:37-41."""
    x_max = np.max(x, axis=axis, keepdims=True)
    z = np.log(np.sum(np.exp(x - x_max), axis=axis, keepdims=True)) + x_max
    return z.sum(axis=axis)


def sigmoid(x):
    return 1 / (1 + np.exp(-x))


def softmax(x):
    """theano.tensor.nnet.softmax over the last axis of a matrix."""
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def cost_gmm(y, mu, sig, weight):
    """model.py:65-91.  mu/sig flat index is d*k + j (model.py:79-80)."""
    shape_y = y.shape
    k = weight.shape[-1]
    D = shape_y[-1]
    y2 = y.reshape((-1, D))[:, :, None]
    mu2 = mu.reshape((-1, D, k))
    sig2 = sig.reshape((-1, D, k))
    w2 = weight.reshape((-1, k))
    diff = (y2 - mu2) ** 2
    inner = y.dtype.type(-0.5) * np.sum(
        diff / sig2 ** 2 + 2 * np.log(sig2) + y.dtype.type(np.log(2 * np.pi)),
        axis=-2)
    nll = -logsumexp(np.log(w2) + inner, axis=-1)
    return nll.reshape(shape_y[:-1])


def multinomial_from_uniform(pvals, unis):
    """theano MultinomialFromUniform: first m with unis < cumsum(pvals)[m].

    Returns the index (== argmax of the one-hot sample, model.py:107-108).  If
    no bucket is hit (cannot happen here, rows sum to 1 + k*eps) returns 0 like
    argmax of an all-zero row.
    """
    c = np.cumsum(pvals, axis=-1, dtype=pvals.dtype)
    hit = unis[:, None] < c
    idx = np.argmax(hit, axis=-1)
    return idx.astype(np.int64)


def sample_gmm(mu, sigma, weight, unis, normals):
    """model.py:94-118 with the RNG draws injected (SURVEY hard part 7).

    unis: (N,) uniforms for the multinomial; normals: (N, D) standard normals.
    """
    k = weight.shape[-1]
    D = mu.shape[-1] // k
    lead = weight.shape[:-1]
    mu2 = mu.reshape((-1, D, k))
    sg2 = sigma.reshape((-1, D, k))
    w2 = weight.reshape((-1, k))
    idx = multinomial_from_uniform(w2, unis.reshape(-1))
    ar = np.arange(mu2.shape[0])
    m = mu2[ar, :, idx]
    s = sg2[ar, :, idx]
    out = m + s * normals.reshape((-1, D))
    return out.reshape(lead + (D,))


def _round_mantissa(x, keep_bits):
    """Round-to-nearest-even a float32 array to ``keep_bits`` explicit mantissa bits."""
    x = np.ascontiguousarray(x, np.float32)
    drop = 23 - keep_bits
    i = x.view(np.uint32).astype(np.uint64)
    i = i + ((1 << (drop - 1)) - 1) + ((i >> drop) & 1)
    i = (i >> drop) << drop
    return i.astype(np.uint32).view(np.float32)


def emulate_mma(x, W, mode):
    """Tensor-core operand-precision emulator (fp32 accumulate), SURVEY section 7 step 1.

    'bf16' / 'tf32' / 'fp16': both operands rounded once.  'bf16x3': each operand
    split hi+lo in bf16, products hi*hi + hi*lo + lo*hi (the lo*lo term dropped) --
    what the CUDA gate-GEMM engine computes."""
    x = np.asarray(x, np.float32)
    W = np.asarray(W, np.float32)
    if mode == 'bf16':
        return _round_mantissa(x, 7) @ _round_mantissa(W, 7)
    if mode == 'tf32':
        return _round_mantissa(x, 10) @ _round_mantissa(W, 10)
    if mode == 'fp16':
        return x.astype(np.float16).astype(np.float32) @ W.astype(np.float16).astype(np.float32)
    if mode == 'bf16x3':
        xh = _round_mantissa(x, 7); xl = _round_mantissa(x - xh, 7)
        Wh = _round_mantissa(W, 7); Wl = _round_mantissa(W - Wh, 7)
        return xh @ Wh + (xh @ Wl + xl @ Wh)
    raise ValueError(mode)


# --------------------------------------------------------------------------
# parameter inventory (Blocks brick-path names)
# --------------------------------------------------------------------------
def _fork_names(prefix, outs):
    return [(prefix + '/fork_' + o) for o in outs]


def param_shapes(cfg):
    """OrderedDict name -> shape for a Parrot configuration (model.py:251-506)."""
    H = cfg['rnn_h_dim']
    R = cfg['readouts_dim']
    D = cfg['output_dim']
    A = cfg['attention_size']
    C = cfg['encoded_input_dim']
    P = OrderedDict()

    def lin(name, i, o):
        P[name + '.W'] = (i, o)
        P[name + '.b'] = (o,)

    def fork(name, i, outs, dims):
        for o, d in zip(outs, dims):
            lin(name + '/fork_' + o, i, d)

    def gru(name, dim):
        P[name + '.state_to_state'] = (dim, dim)
        P[name + '.state_to_gates'] = (dim, 2 * dim)
        P[name + '.initial_state'] = (dim,)

    root = '/parrot'
    if cfg['encoder_type'] == 'bidirectional':
        P[root + '/encoder/embed_label.W'] = (cfg['num_characters'], cfg['input_dim'])
        E = cfg['encoder_dim']
        for d in ('forward', 'backward'):
            base = root + '/encoder/encoder/' + d
            gru(base + '/gatedrecurrent', E)
            fork(base + '/fork', cfg['input_dim'], ['inputs', 'gate_inputs'], [E, 2 * E])
    for i in (1, 2, 3):
        gru(root + '/rnn%d' % i, H)
    for i in (1, 2, 3):
        lin(root + '/h%d_to_readout' % i, H, R)
    fork(root + '/h1_to_h2', H, ['rnn2_inputs', 'rnn2_gates'], [H, 2 * H])
    fork(root + '/h1_to_h3', H, ['rnn3_inputs', 'rnn3_gates'], [H, 2 * H])
    fork(root + '/h2_to_h3', H, ['rnn3_inputs', 'rnn3_gates'], [H, 2 * H])
    if cfg['which_cost'] == 'MSE':
        lin(root + '/readout_to_output', R, D)
    else:
        k = cfg['k_gmm']
        fork(root + '/readout_to_output', R, ['gmm_mu', 'gmm_sigma', 'gmm_coeff'],
             [D * k, D * k, k])
    for i in (1, 2, 3):
        fork(root + '/inp_to_h%d' % i, C, ['rnn%d_inputs' % i, 'rnn%d_gates' % i], [H, 2 * H])
    fork(root + '/h1_to_att', H, ['alpha', 'beta', 'kappa'], [A, A, A])
    lin(root + '/att_to_readout', C, R)
    if cfg['use_speaker']:
        S = cfg['speaker_dim']
        P[root + '/lookuptable.W'] = (cfg['num_speakers'], S)
        for i in (1, 2, 3):
            fork(root + '/speaker_to_h%d' % i, S,
                 ['rnn%d_inputs' % i, 'rnn%d_gates' % i], [H, 2 * H])
        lin(root + '/speaker_to_readout', S, R)
        if cfg['which_cost'] == 'MSE':
            lin(root + '/speaker_to_output', S, D)
        else:
            k = cfg['k_gmm']
            fork(root + '/speaker_to_output', S, ['gmm_mu', 'gmm_sigma', 'gmm_coeff'],
                 [D * k, D * k, k])
    if cfg['full_feedback']:
        fork(root + '/out_to_h2', D, ['rnn2_inputs', 'rnn2_gates'], [H, 2 * H])
        fork(root + '/out_to_h3', D, ['rnn3_inputs', 'rnn3_gates'], [H, 2 * H])
    if cfg['weak_feedback']:
        fork(root + '/out_to_h1', D, ['rnn1_inputs', 'rnn1_gates'], [H, 2 * H])
    P[root + '.initial_w'] = (C,)
    return P


DEFAULTS = dict(
    input_dim=420, output_dim=63, rnn_h_dim=1024, readouts_dim=1024,
    weak_feedback=False, full_feedback=False, feedback_noise_level=None,
    layer_norm=False, use_speaker=False, num_speakers=21, speaker_dim=128,
    which_cost='MSE', k_gmm=20, sampling_bias=0, epsilon=1e-5,
    num_characters=43, attention_type='graves', attention_size=10,
    attention_alignment=1., sharpening_coeff=1., timing_coeff=1.,
    encoder_type=None, encoder_dim=128, raw_output=False)


class OracleParrot(object):
    """numpy restatement of model.py:Parrot (model.py:250-1111).

    Extra (non-reference) keyword: ``encoder_time_axis`` -- 0 reproduces the
    reference literally (Blocks' recurrent scan runs over axis 0 of the
    batch-major (B, U, emb) tensor, SURVEY D4 / hazard H1); 1 is the evidently
    intended recurrence over text positions.
    """

    def __init__(self, dtype=np.float32, encoder_time_axis=0, **kwargs):
        cfg = dict(DEFAULTS)
        for k in ('weights_init', 'biases_init', 'name'):
            kwargs.pop(k, None)
        unknown = set(kwargs) - set(cfg)
        assert not unknown, 'unknown Parrot kwargs: %s' % sorted(unknown)
        cfg.update(kwargs)
        assert cfg['encoder_type'] in (None, 'bidirectional')  # model.py:209
        assert cfg['which_cost'] in ('MSE', 'GMM')
        assert cfg['attention_type'] in ('graves', 'softmax')
        assert not cfg['raw_output'], 'sampleRNN coupling is out of scope (SURVEY 8f)'
        if cfg['full_feedback']:
            cfg['weak_feedback'] = True  # model.py:485
        if cfg['encoder_type'] is None:
            assert cfg['num_characters'] == cfg['input_dim']  # model.py:224
        cfg['encoded_input_dim'] = (2 * cfg['encoder_dim']
                                    if cfg['encoder_type'] == 'bidirectional'
                                    else cfg['input_dim'])  # model.py:302-307
        self.cfg = cfg
        self.__dict__.update(cfg)
        self.dtype = np.dtype(dtype)
        self.encoder_time_axis = encoder_time_axis
        self.shapes = param_shapes(cfg)
        self.params = OrderedDict(
            (n, np.zeros(s, self.dtype)) for n, s in self.shapes.items())
        self._state_B = None
        self._cache = None
        self.mma_emulation = None   # None | 'bf16' | 'fp16' | 'tf32' | 'bf16x3'

    # ---------------------------------------------------------------- init
    def initialize(self, rng=None, std=0.01, trained_like=False, gain=None):
        """train.py:30-31: W ~ N(0, 0.01^2), b = 0; GRU initial_state / initial_w = 0.

        gain=g draws W ~ N(0, (g/sqrt(fan_in))^2) instead (SURVEY 8d synthetic
        'trained-like' parameter set) so activations are not all ~0; trained_like=True
        is gain 1 (chaotic over long segments: any perturbation is amplified, so
        parity tests use gain <= 0.5).
        """
        rng = rng or np.random.default_rng(0)
        if trained_like and gain is None:
            gain = 1.0
        for n, s in self.shapes.items():
            if n.endswith('.b') or n.endswith('.initial_state') or n.endswith('.initial_w'):
                self.params[n] = np.zeros(s, self.dtype)
            else:
                sd = std if gain is None else gain / np.sqrt(s[0])
                if gain is not None and ('lookuptable' in n or 'embed_label' in n):
                    sd = 1.0
                self.params[n] = (rng.standard_normal(s) * sd).astype(self.dtype)
        return self

    def set_params(self, params):
        for n in self.shapes:
            self.params[n] = np.asarray(params[n], dtype=self.dtype).reshape(self.shapes[n])

    # ------------------------------------------------------------- helpers
    def _p(self, name):
        return self.params['/parrot' + name]

    def _mm(self, x, W):
        """Every forward contraction goes through here so that tests can emulate
        reduced-precision MMA operands (see ``emulate_mma``)."""
        if self.mma_emulation is None:
            return x @ W
        return emulate_mma(x, W, self.mma_emulation)

    def _lin(self, name, x):
        return self._mm(x, self._p(name + '.W')) + self._p(name + '.b')

    def _fork(self, name, outs, x):
        return [self._lin(name + '/fork_' + o, x) for o in outs]

    def _gru(self, name, inputs, gate_inputs, s):
        """blocks GatedRecurrent.apply: gates packed [update, reset]."""
        H = s.shape[-1]
        g = sigmoid(self._mm(s, self._p(name + '.state_to_gates')) + gate_inputs)
        z = g[:, :H]
        r = g[:, H:]
        c = np.tanh(self._mm(s * r, self._p(name + '.state_to_state')) + inputs)
        s_new = c * z + s * (1 - z)
        return s_new, (z, r, c)

    # -------------------------------------------------------------- states
    def initial_states(self, batch_size):
        """model.py:529-549 (10-tuple order kept)."""
        B = batch_size
        f = self.dtype
        if self._state_B != B:
            self.last_h1 = np.zeros((B, self.rnn_h_dim), f)
            self.last_h2 = np.zeros((B, self.rnn_h_dim), f)
            self.last_h3 = np.zeros((B, self.rnn_h_dim), f)
            self.last_k = np.zeros((B, self.attention_size), f)
            self.last_w = np.zeros((B, self.encoded_input_dim), f)
            self._state_B = B
        ih = [np.repeat(self._p('/rnn%d.initial_state' % i)[None, :], B, 0) for i in (1, 2, 3)]
        initial_k = np.zeros((B, self.attention_size), f)
        initial_w = np.repeat(self._p('.initial_w')[None, :], B, 0)
        return (ih[0], self.last_h1, ih[1], self.last_h2, ih[2], self.last_h3,
                initial_w, self.last_w, initial_k, self.last_k)

    # ------------------------------------------------------------- encoder
    def _encoder_fwd(self, labels):
        """model.py:233-247.  Returns (B, U, C) and a cache for the backward."""
        if self.encoder_type is None:
            # labels are already (B, U, input_dim) features (model.py:235-236)
            return np.asarray(labels, self.dtype), None
        emb = self._p('/encoder/embed_label.W')[labels]          # (B, U, emb)
        seq = emb if self.encoder_time_axis == 0 else emb.transpose(1, 0, 2)
        L, N, _ = seq.shape
        E = self.encoder_dim
        outs = []
        caches = []
        for d in ('forward', 'backward'):
            base = '/encoder/encoder/' + d
            xi = self._lin(base + '/fork/fork_inputs', seq)
            xg = self._lin(base + '/fork/fork_gate_inputs', seq)
            s = np.repeat(self._p(base + '/gatedrecurrent.initial_state')[None, :], N, 0)
            out = np.zeros((L, N, E), self.dtype)
            steps = []
            order = range(L) if d == 'forward' else range(L - 1, -1, -1)
            for i in order:
                s_prev = s
                s, (z, r, c) = self._gru(base + '/gatedrecurrent', xi[i], xg[i], s_prev)
                out[i] = s
                steps.append((i, s_prev, z, r, c))
            outs.append(out)
            caches.append(steps)
        enc = np.concatenate(outs, axis=2)
        if self.encoder_time_axis != 0:
            enc = enc.transpose(1, 0, 2)
        return enc, (labels, seq, caches)

    def _gru_bwd(self, name, ds_new, s_prev, z, r, c, grads):
        """Backward of one GRU step.  Returns (d_inputs, d_gate_inputs, d_s_prev)."""
        Ws = self._p(name + '.state_to_state')
        Wg = self._p(name + '.state_to_gates')
        dc = ds_new * z
        dz = ds_new * (c - s_prev)
        ds_prev = ds_new * (1 - z)
        da_c = dc * (1 - c * c)
        rs = s_prev * r
        grads['/parrot' + name + '.state_to_state'] += rs.T @ da_c
        drs = da_c @ Ws.T
        dr = drs * s_prev
        ds_prev = ds_prev + drs * r
        da_g = np.concatenate([dz * z * (1 - z), dr * r * (1 - r)], axis=1)
        grads['/parrot' + name + '.state_to_gates'] += s_prev.T @ da_g
        ds_prev = ds_prev + da_g @ Wg.T
        return da_c, da_g, ds_prev

    def _lin_bwd(self, name, x, dy, grads, need_dx=True):
        x2 = x.reshape(-1, x.shape[-1])
        dy2 = dy.reshape(-1, dy.shape[-1])
        grads['/parrot' + name + '.W'] += x2.T @ dy2
        grads['/parrot' + name + '.b'] += dy2.sum(0)
        if need_dx:
            return dy @ self._p(name + '.W').T
        return None

    def _encoder_bwd(self, denc, cache, grads):
        if cache is None:
            return
        labels, seq, caches = cache
        if self.encoder_time_axis != 0:
            denc = denc.transpose(1, 0, 2)
        L, N, _ = seq.shape
        E = self.encoder_dim
        dseq = np.zeros_like(seq)
        for di, d in enumerate(('forward', 'backward')):
            base = '/encoder/encoder/' + d
            dout = denc[:, :, di * E:(di + 1) * E]
            dxi = np.zeros((L, N, E), self.dtype)
            dxg = np.zeros((L, N, 2 * E), self.dtype)
            ds = np.zeros((N, E), self.dtype)
            for (i, s_prev, z, r, c) in reversed(caches[di]):
                ds = ds + dout[i]
                dxi[i], dxg[i], ds = self._gru_bwd(
                    base + '/gatedrecurrent', ds, s_prev, z, r, c, grads)
            grads['/parrot' + base + '/gatedrecurrent.initial_state'] += ds.sum(0)
            dseq += self._lin_bwd(base + '/fork/fork_inputs', seq, dxi, grads)
            dseq += self._lin_bwd(base + '/fork/fork_gate_inputs', seq, dxg, grads)
        demb = dseq if self.encoder_time_axis == 0 else dseq.transpose(1, 0, 2)
        np.add.at(grads['/parrot/encoder/embed_label.W'], labels, demb)

    # ----------------------------------------------------------- attention
    def _attention(self, h1, k_prev, ctx, u, sharpening=1.0, timing=1.0):
        """model.py:664-690 (train) / model.py:931-958 (sampling)."""
        f = self.dtype.type
        a_hat, b_hat, k_hat = self._fork('/h1_to_att', ['alpha', 'beta', 'kappa'], h1)
        if self.attention_type == 'softmax':
            a = softmax(a_hat) + f(self.epsilon)
        else:
            a = np.exp(a_hat) + f(self.epsilon)
        if sharpening == 1.0:
            b = np.exp(b_hat) + f(self.epsilon)
        else:
            b = np.exp(b_hat) * f(sharpening) + f(self.epsilon)
        if timing == 1.0:
            k = k_prev + f(self.attention_alignment) * np.exp(k_hat)
        else:
            k = k_prev + f(self.attention_alignment) * np.exp(k_hat) / f(timing)
        a_ = a[:, :, None]
        b_ = b[:, :, None]
        k_ = k[:, :, None]
        if self.attention_type == 'softmax':
            phi = f(SQRT_1_2PI) * np.sum(
                a_ * np.sqrt(b_) * np.exp(f(-0.5) * b_ * (k_ - u) ** 2), axis=1)
        else:
            phi = np.sum(a_ * np.exp(-b_ * (k_ - u) ** 2), axis=1)
        w = (phi[:, :, None] * ctx).sum(axis=1)
        return a, b, k, phi, w, (a_hat, b_hat, k_hat)

    def _attention_bwd(self, dw, dk_next, dphi_ext, h1, k_prev, ctx, u, a, b, k, phi,
                       hats, grads, dctx):
        """Backward of _attention (train form).  Returns (dh1, dk_prev).

        dctx is accumulated in place.  dphi_ext / da_ext are direct gradients on
        the returned phi (None in training: the cost does not read them)."""
        f = self.dtype.type
        a_hat, b_hat, k_hat = hats
        dphi = np.einsum('bc,buc->bu', dw, ctx)
        if dphi_ext is not None:
            dphi = dphi + dphi_ext
        dctx += phi[:, :, None] * dw[:, None, :]
        diff = k[:, :, None] - u                              # (B, A, U)
        if self.attention_type == 'softmax':
            e = np.exp(f(-0.5) * b[:, :, None] * diff ** 2)
            sb = np.sqrt(b)[:, :, None]
            g = f(SQRT_1_2PI) * dphi[:, None, :]              # (B,1,U)
            da = (g * sb * e).sum(-1)
            db = (g * a[:, :, None] * e * (f(0.5) / sb - sb * f(0.5) * diff ** 2)).sum(-1)
            dk = (g * a[:, :, None] * sb * e * (-b[:, :, None] * diff)).sum(-1)
        else:
            e = np.exp(-b[:, :, None] * diff ** 2)
            g = dphi[:, None, :]
            da = (g * e).sum(-1)
            db = (g * a[:, :, None] * e * (-(diff ** 2))).sum(-1)
            dk = (g * a[:, :, None] * e * (-2 * b[:, :, None] * diff)).sum(-1)
        dk = dk + dk_next
        dk_prev = dk
        dk_hat = dk * f(self.attention_alignment) * np.exp(k_hat)
        db_hat = db * np.exp(b_hat)
        if self.attention_type == 'softmax':
            sm = a - f(self.epsilon)
            da_hat = sm * (da - (da * sm).sum(-1, keepdims=True))
        else:
            da_hat = da * np.exp(a_hat)
        dh1 = self._lin_bwd('/h1_to_att/fork_alpha', h1, da_hat, grads)
        dh1 = dh1 + self._lin_bwd('/h1_to_att/fork_beta', h1, db_hat, grads)
        dh1 = dh1 + self._lin_bwd('/h1_to_att/fork_kappa', h1, dk_hat, grads)
        return dh1, dk_prev

    # --------------------------------------------------------- compute_cost
    def compute_cost(self, features, features_mask, labels, labels_mask,
                     speaker, start_flag, batch_size, raw_audio=None,
                     feedback_noise=None, noise_level=None,
                     gmm_unis=None, gmm_normals=None, keep_cache=True):
        """model.py:552-824.  Eager: returns (cost, updates, attention_vars, cost_raw).

        ``updates`` is the list [(name, value)] of carried-state updates
        (model.py:786-791); they are also applied to this object, like the
        compiled Theano function would.  RNG draws are injected:
        ``feedback_noise`` ~ N(0,1) of features[:-1].shape (model.py:575-578),
        ``gmm_unis`` (T,B) / ``gmm_normals`` (T,B,D) for next_x in GMM mode
        (model.py:782).
        """
        f = self.dtype
        ft = f.type
        if speaker is None:
            assert not self.use_speaker            # model.py:556-557
        features = np.asarray(features, f)
        features_mask = np.asarray(features_mask, f)
        labels_mask = np.asarray(labels_mask, f)
        B = batch_size
        H = self.rnn_h_dim
        target = features[1:]                      # model.py:559
        mask = features_mask[1:]                   # model.py:560
        T = mask.shape[0]
        LN = self.layer_norm
        cell = [np.zeros((T, B, H), f) for _ in range(3)]      # model.py:562-569
        gat = [np.zeros((T, B, 2 * H), f) for _ in range(3)]
        cache = dict(T=T, B=B)

        x_in = None
        if self.weak_feedback:                     # model.py:571-588
            x_in = features[:-1]
            if self.feedback_noise_level:          # truthiness gate, hazard H6
                x_in = x_in + ft(noise_level) * np.asarray(feedback_noise, f)
            oc, og = self._fork('/out_to_h1', ['rnn1_inputs', 'rnn1_gates'], x_in)
            cache['fb1_pre'] = (oc, og)
            cell[0] = cell[0] + apply_norm(oc, LN)
            gat[0] = gat[0] + apply_norm(og, LN)
        if self.full_feedback:                     # model.py:590-603
            for li, nm in ((1, '/out_to_h2'), (2, '/out_to_h3')):
                oc, og = self._fork(nm, ['rnn%d_inputs' % (li + 1), 'rnn%d_gates' % (li + 1)], x_in)
                cache['fb%d_pre' % (li + 1)] = (oc, og)
                cell[li] = cell[li] + apply_norm(oc, LN)
                gat[li] = gat[li] + apply_norm(og, LN)
        cache['x_in'] = x_in

        emb_spk = None
        if self.use_speaker:                       # model.py:605-627
            spk_idx = np.asarray(speaker)[:, 0]
            emb_spk = self._p('/lookuptable.W')[spk_idx][None, :, :]   # (1,B,S)
            cache['spk_idx'] = spk_idx
            cache['spk_pre'] = []
            for li in range(3):
                sc, sg = self._fork('/speaker_to_h%d' % (li + 1),
                                    ['rnn%d_inputs' % (li + 1), 'rnn%d_gates' % (li + 1)], emb_spk)
                cache['spk_pre'].append((sc, sg))
                cell[li] = apply_norm(sc, LN) + cell[li]
                gat[li] = apply_norm(sg, LN) + gat[li]
        cache['emb_spk'] = emb_spk

        (initial_h1, last_h1, initial_h2, last_h2, initial_h3, last_h3,
         initial_w, last_w, initial_k, last_k) = self.initial_states(B)
        use_init = bool(start_flag != 0)           # tensor.switch on a float scalar, H4
        h1 = initial_h1 if use_init else last_h1   # model.py:633-643
        h2 = initial_h2 if use_init else last_h2
        h3 = initial_h3 if use_init else last_h3
        w = initial_w if use_init else last_w
        k = initial_k if use_init else last_k
        cache['use_init'] = use_init
        cache['in_states'] = (h1, h2, h3, k, w)

        enc, enc_cache = self._encoder_fwd(labels)
        ctx = enc * labels_mask[:, :, None]        # model.py:645-646
        U = labels_mask.shape[1]
        u = np.arange(U, dtype=f)[None, None, :]   # model.py:648-649
        cache.update(enc_cache=enc_cache, ctx=ctx, labels_mask=labels_mask, u=u)

        C = self.encoded_input_dim
        A = self.attention_size
        h1s = np.zeros((T, B, H), f); h2s = np.zeros((T, B, H), f); h3s = np.zeros((T, B, H), f)
        ks = np.zeros((T, B, A), f); ws = np.zeros((T, B, C), f)
        phis = np.zeros((T, B, U), f); pis = np.zeros((T, B, A), f)
        steps = []
        for t in range(T):                          # model.py:651-724
            st = {}
            st['prev'] = (h1, h2, h3, k, w)
            ai, ag = self._fork('/inp_to_h1', ['rnn1_inputs', 'rnn1_gates'], w)
            h1n, g1 = self._gru('/rnn1', cell[0][t] + ai, gat[0][t] + ag, h1)
            a, b, kn, phi, wn, hats = self._attention(h1n, k, ctx, u)
            i2, g2i = self._fork('/inp_to_h2', ['rnn2_inputs', 'rnn2_gates'], wn)
            i3, g3i = self._fork('/inp_to_h3', ['rnn3_inputs', 'rnn3_gates'], wn)
            q12 = self._fork('/h1_to_h2', ['rnn2_inputs', 'rnn2_gates'], h1n)
            q13 = self._fork('/h1_to_h3', ['rnn3_inputs', 'rnn3_gates'], h1n)
            p12 = [apply_norm(x, LN) for x in q12]
            p13 = [apply_norm(x, LN) for x in q13]
            h2n, g2 = self._gru('/rnn2', cell[1][t] + i2 + p12[0], gat[1][t] + g2i + p12[1], h2)
            q23 = self._fork('/h2_to_h3', ['rnn3_inputs', 'rnn3_gates'], h2n)
            p23 = [apply_norm(x, LN) for x in q23]
            h3n, g3 = self._gru('/rnn3', cell[2][t] + i3 + p13[0] + p23[0],
                                gat[2][t] + g3i + p13[1] + p23[1], h3)
            if keep_cache:
                st.update(g1=g1, g2=g2, g3=g3, a=a, b=b, hats=hats,
                          q12=q12 if LN else None, q13=q13 if LN else None,
                          q23=q23 if LN else None)
                steps.append(st)
            h1, h2, h3, k, w = h1n, h2n, h3n, kn, wn
            h1s[t] = h1; h2s[t] = h2; h3s[t] = h3; ks[t] = k; ws[t] = w
            phis[t] = phi; pis[t] = a
        cache['steps'] = steps
        cache.update(h1s=h1s, h2s=h2s, h3s=h3s, ks=ks, ws=ws, phis=phis)

        ro_pre = [self._lin('/h%d_to_readout' % (i + 1), hs)     # model.py:739-741
                  for i, hs in enumerate((h1s, h2s, h3s))]
        ro = [apply_norm(x, LN) for x in ro_pre]                 # model.py:743-746
        readouts = ro[0] + ro[1] + ro[2]                         # model.py:748
        if self.use_speaker:
            readouts = readouts + self._lin('/speaker_to_readout', emb_spk)   # model.py:751
        readouts = readouts + self._lin('/att_to_readout', ws)   # model.py:753
        cache.update(ro_pre=ro_pre if LN else None, readouts=readouts)

        if self.which_cost == 'MSE':                             # model.py:757-764
            predicted = self._lin('/readout_to_output', readouts)
            if self.use_speaker:
                predicted = predicted + self._lin('/speaker_to_output', emb_spk)
            cost_tb = np.sum((predicted - target) ** 2, axis=-1)
            next_x = predicted
            coeff = predicted
            cache.update(predicted=predicted)
        else:                                                    # model.py:765-782
            mu, sg_hat, co_hat = self._fork(
                '/readout_to_output', ['gmm_mu', 'gmm_sigma', 'gmm_coeff'], readouts)
            if self.use_speaker:
                s0, s1, s2 = self._fork(
                    '/speaker_to_output', ['gmm_mu', 'gmm_sigma', 'gmm_coeff'], emb_spk)
                mu = mu + s0; sg_hat = sg_hat + s1; co_hat = co_hat + s2
            sigma = np.exp(sg_hat) + ft(self.epsilon)            # model.py:774
            kg = self.k_gmm
            coeff = softmax(co_hat.reshape((-1, kg))).reshape(co_hat.shape) + ft(self.epsilon)
            cost_tb = cost_gmm(target, mu, sigma, coeff)         # model.py:781
            if gmm_unis is not None:
                next_x = sample_gmm(mu, sigma, coeff, np.asarray(gmm_unis, f),
                                    np.asarray(gmm_normals, f))  # model.py:782
            else:
                next_x = None
            cache.update(mu=mu, sg_hat=sg_hat, co_hat=co_hat, sigma=sigma, coeff=coeff)
        msum = mask.sum() + ft(1e-5)
        cost = (cost_tb * mask).sum() / msum + ft(0.) * ft(start_flag)   # model.py:784
        cache.update(target=target, mask=mask, msum=msum)

        updates = [('last_h1', h1s[-1]), ('last_h2', h2s[-1]), ('last_h3', h3s[-1]),
                   ('last_k', ks[-1]), ('last_w', ws[-1])]       # model.py:786-791
        self.last_h1, self.last_h2, self.last_h3 = h1s[-1].copy(), h2s[-1].copy(), h3s[-1].copy()
        self.last_k, self.last_w = ks[-1].copy(), ws[-1].copy()
        self._cache = cache if keep_cache else None
        attention_vars = [next_x, ks, ws, coeff, phis, pis]     # model.py:822
        return cost, updates, attention_vars, None

    # -------------------------------------------------------------- backward
    def backward(self, unnormalised=False):
        """Gradient of the last compute_cost wrt every parameter (what
        ``theano.grad`` inside blocks GradientDescent computes, train.py:103-107).

        Carried states entering the segment are constants (shared variables),
        except that with start_flag != 0 the initial states are parameters.
        unnormalised=True returns d(sum cost*mask) instead of d(cost)
        (the quantity allreduced in data-parallel training).
        """
        c = self._cache
        assert c is not None, 'call compute_cost(keep_cache=True) first'
        f = self.dtype
        ft = f.type
        T, B = c['T'], c['B']
        H = self.rnn_h_dim
        LN = self.layer_norm
        grads = OrderedDict((n, np.zeros(s, f)) for n, s in self.shapes.items())
        scale = c['mask'] if unnormalised else c['mask'] / c['msum']     # (T,B)

        # ---- emitter
        readouts = c['readouts']
        emb_spk = c['emb_spk']
        if self.which_cost == 'MSE':
            dpred = 2 * (c['predicted'] - c['target']) * scale[:, :, None]
            dread = self._lin_bwd('/readout_to_output', readouts, dpred, grads)
            if self.use_speaker:
                demb = self._lin_bwd('/speaker_to_output',
                                     emb_spk, dpred.sum(0, keepdims=True), grads)
        else:
            kg = self.k_gmm
            D = self.output_dim
            y = c['target'].reshape(-1, D)[:, :, None]
            mu = c['mu'].reshape(-1, D, kg)
            sig = c['sigma'].reshape(-1, D, kg)
            pi = c['coeff'].reshape(-1, kg)
            diff = y - mu
            inner = ft(-0.5) * np.sum(diff ** 2 / sig ** 2 + 2 * np.log(sig)
                                      + ft(np.log(2 * np.pi)), axis=1)
            lw = np.log(pi) + inner
            rho = softmax(lw)                                   # responsibilities (N,k)
            sc = scale.reshape(-1)[:, None]
            dpi = -(rho / pi) * sc
            dmu = -(rho[:, None, :] * diff / sig ** 2) * sc[:, :, None]
            dsig = -(rho[:, None, :] * (diff ** 2 / sig ** 3 - 1 / sig)) * sc[:, :, None]
            sm = pi - ft(self.epsilon)
            dco_hat = sm * (dpi - (dpi * sm).sum(-1, keepdims=True))
            dsg_hat = dsig * np.exp(c['sg_hat'].reshape(-1, D, kg))
            dmu = dmu.reshape(T, B, D * kg)
            dsg_hat = dsg_hat.reshape(T, B, D * kg)
            dco_hat = dco_hat.reshape(T, B, kg)
            dread = self._lin_bwd('/readout_to_output/fork_gmm_mu', readouts, dmu, grads)
            dread = dread + self._lin_bwd('/readout_to_output/fork_gmm_sigma', readouts, dsg_hat, grads)
            dread = dread + self._lin_bwd('/readout_to_output/fork_gmm_coeff', readouts, dco_hat, grads)
            if self.use_speaker:
                demb = self._lin_bwd('/speaker_to_output/fork_gmm_mu', emb_spk,
                                     dmu.sum(0, keepdims=True), grads)
                demb = demb + self._lin_bwd('/speaker_to_output/fork_gmm_sigma', emb_spk,
                                            dsg_hat.sum(0, keepdims=True), grads)
                demb = demb + self._lin_bwd('/speaker_to_output/fork_gmm_coeff', emb_spk,
                                            dco_hat.sum(0, keepdims=True), grads)
        # ---- readouts
        dws = self._lin_bwd('/att_to_readout', c['ws'], dread, grads)      # (T,B,C)
        if self.use_speaker:
            demb = demb + self._lin_bwd('/speaker_to_readout', emb_spk,
                                        dread.sum(0, keepdims=True), grads)
        dhs = []
        for i, hs in enumerate((c['h1s'], c['h2s'], c['h3s'])):
            d = simple_norm_bwd(dread, c['ro_pre'][i]) if LN else dread
            dhs.append(self._lin_bwd('/h%d_to_readout' % (i + 1), hs, d, grads))
        dh1s, dh2s, dh3s = dhs

        # ---- reverse scan
        ctx = c['ctx']; u = c['u']
        dctx = np.zeros_like(ctx)
        dcell = [np.zeros((T, B, H), f) for _ in range(3)]
        dgat = [np.zeros((T, B, 2 * H), f) for _ in range(3)]
        dh1_n = np.zeros((B, H), f); dh2_n = np.zeros((B, H), f); dh3_n = np.zeros((B, H), f)
        dk_n = np.zeros((B, self.attention_size), f)
        dw_n = np.zeros((B, self.encoded_input_dim), f)
        for t in range(T - 1, -1, -1):
            st = c['steps'][t]
            h1p, h2p, h3p, kp, wp = st['prev']
            h1, h2, h3 = c['h1s'][t], c['h2s'][t], c['h3s'][t]
            k, w, phi = c['ks'][t], c['ws'][t], c['phis'][t]
            dh1 = dh1s[t] + dh1_n
            dh2 = dh2s[t] + dh2_n
            dh3 = dh3s[t] + dh3_n
            dw = dws[t] + dw_n
            # layer 3
            z, r, cc = st['g3']
            dc3, dg3, dh3_n = self._gru_bwd('/rnn3', dh3, h3p, z, r, cc, grads)
            dcell[2][t] = dc3; dgat[2][t] = dg3
            dw = dw + self._lin_bwd('/inp_to_h3/fork_rnn3_inputs', w, dc3, grads)
            dw = dw + self._lin_bwd('/inp_to_h3/fork_rnn3_gates', w, dg3, grads)
            d23c = simple_norm_bwd(dc3, st['q23'][0]) if LN else dc3
            d23g = simple_norm_bwd(dg3, st['q23'][1]) if LN else dg3
            dh2 = dh2 + self._lin_bwd('/h2_to_h3/fork_rnn3_inputs', h2, d23c, grads)
            dh2 = dh2 + self._lin_bwd('/h2_to_h3/fork_rnn3_gates', h2, d23g, grads)
            d13c = simple_norm_bwd(dc3, st['q13'][0]) if LN else dc3
            d13g = simple_norm_bwd(dg3, st['q13'][1]) if LN else dg3
            dh1 = dh1 + self._lin_bwd('/h1_to_h3/fork_rnn3_inputs', h1, d13c, grads)
            dh1 = dh1 + self._lin_bwd('/h1_to_h3/fork_rnn3_gates', h1, d13g, grads)
            # layer 2
            z, r, cc = st['g2']
            dc2, dg2, dh2_n = self._gru_bwd('/rnn2', dh2, h2p, z, r, cc, grads)
            dcell[1][t] = dc2; dgat[1][t] = dg2
            dw = dw + self._lin_bwd('/inp_to_h2/fork_rnn2_inputs', w, dc2, grads)
            dw = dw + self._lin_bwd('/inp_to_h2/fork_rnn2_gates', w, dg2, grads)
            d12c = simple_norm_bwd(dc2, st['q12'][0]) if LN else dc2
            d12g = simple_norm_bwd(dg2, st['q12'][1]) if LN else dg2
            dh1 = dh1 + self._lin_bwd('/h1_to_h2/fork_rnn2_inputs', h1, d12c, grads)
            dh1 = dh1 + self._lin_bwd('/h1_to_h2/fork_rnn2_gates', h1, d12g, grads)
            # attention
            dh1_att, dk_n = self._attention_bwd(
                dw, dk_n, None, h1, kp, ctx, u, st['a'], st['b'], k, phi, st['hats'],
                grads, dctx)
            dh1 = dh1 + dh1_att
            # layer 1
            z, r, cc = st['g1']
            dc1, dg1, dh1_n = self._gru_bwd('/rnn1', dh1, h1p, z, r, cc, grads)
            dcell[0][t] = dc1; dgat[0][t] = dg1
            dw_n = self._lin_bwd('/inp_to_h1/fork_rnn1_inputs', wp, dc1, grads)
            dw_n = dw_n + self._lin_bwd('/inp_to_h1/fork_rnn1_gates', wp, dg1, grads)

        # ---- initial states (parameters only when start_flag selected them)
        if c['use_init']:
            grads['/parrot/rnn1.initial_state'] += dh1_n.sum(0)
            grads['/parrot/rnn2.initial_state'] += dh2_n.sum(0)
            grads['/parrot/rnn3.initial_state'] += dh3_n.sum(0)
            grads['/parrot.initial_w'] += dw_n.sum(0)

        # ---- context / encoder
        denc = dctx * c['labels_mask'][:, :, None]
        self._encoder_bwd(denc, c['enc_cache'], grads)

        # ---- speaker conditioning of the recurrent inputs
        if self.use_speaker:
            for li in range(3):
                sc, sg = c['spk_pre'][li]
                dc = dcell[li].sum(0, keepdims=True)
                dg = dgat[li].sum(0, keepdims=True)
                if LN:
                    dc = simple_norm_bwd(dc, sc); dg = simple_norm_bwd(dg, sg)
                demb = demb + self._lin_bwd(
                    '/speaker_to_h%d/fork_rnn%d_inputs' % (li + 1, li + 1), emb_spk, dc, grads)
                demb = demb + self._lin_bwd(
                    '/speaker_to_h%d/fork_rnn%d_gates' % (li + 1, li + 1), emb_spk, dg, grads)
            np.add.at(grads['/parrot/lookuptable.W'], c['spk_idx'], demb[0])

        # ---- teacher-forcing feedback
        x_in = c['x_in']
        fb = []
        if self.weak_feedback:
            fb.append((0, '/out_to_h1', 'fb1_pre'))
        if self.full_feedback:
            fb.append((1, '/out_to_h2', 'fb2_pre'))
            fb.append((2, '/out_to_h3', 'fb3_pre'))
        for li, nm, key in fb:
            oc, og = c[key]
            dc = simple_norm_bwd(dcell[li], oc) if LN else dcell[li]
            dg = simple_norm_bwd(dgat[li], og) if LN else dgat[li]
            self._lin_bwd(nm + '/fork_rnn%d_inputs' % (li + 1), x_in, dc, grads, need_dx=False)
            self._lin_bwd(nm + '/fork_rnn%d_gates' % (li + 1), x_in, dg, grads, need_dx=False)
        self._bwd_aux = dict(dcell=dcell, dgat=dgat, dctx=dctx, dh1s=dh1s, dws=dws)
        return grads

    # --------------------------------------------------------------- sampling
    def sample_model(self, labels_tr, labels_mask_tr, features_mask_tr, speaker_tr,
                     num_samples, num_steps, gmm_unis=None, gmm_normals=None):
        """model.py:1061-1083 -> sample_model_fun (model.py:827-1059).

        Returns [x, k, w, pi, phi, pi_att], all time-major numpy arrays.
        ``features_mask_tr`` is unused, as in the reference.
        RNG draws injected: gmm_unis (T,B), gmm_normals (T,B,D).
        """
        f = self.dtype
        ft = f.type
        B = num_samples
        T = num_steps
        H = self.rnn_h_dim
        LN = self.layer_norm
        labels_mask = np.asarray(labels_mask_tr, f)
        # model.py:830-832: always the *initial* states
        ih = [np.repeat(self._p('/rnn%d.initial_state' % i)[None, :], B, 0) for i in (1, 2, 3)]
        h1, h2, h3 = ih
        k = np.zeros((B, self.attention_size), f)
        w = np.repeat(self._p('.initial_w')[None, :], B, 0)
        x = np.zeros((B, self.output_dim), f)                  # model.py:834-835
        cellb = [np.zeros((B, H), f) for _ in range(3)]
        gatb = [np.zeros((B, 2 * H), f) for _ in range(3)]
        spk_readout = spk_output = None
        if self.use_speaker:                                    # model.py:846-874
            emb = self._p('/lookuptable.W')[np.asarray(speaker_tr)[:, 0]]
            spk_readout = self._lin('/speaker_to_readout', emb)
            if self.which_cost == 'MSE':
                spk_output = self._lin('/speaker_to_output', emb)
            else:
                spk_output = self._fork('/speaker_to_output',
                                        ['gmm_mu', 'gmm_sigma', 'gmm_coeff'], emb)
            for li in range(3):
                sc, sg = self._fork('/speaker_to_h%d' % (li + 1),
                                    ['rnn%d_inputs' % (li + 1), 'rnn%d_gates' % (li + 1)],
                                    emb[None])
                cellb[li] = cellb[li] + apply_norm(sc, LN)[0]
                gatb[li] = gatb[li] + apply_norm(sg, LN)[0]
        enc, _ = self._encoder_fwd(labels_tr)
        ctx = enc * labels_mask[:, :, None]                     # model.py:876-877
        U = labels_mask.shape[1]
        u = np.arange(U, dtype=f)[None, None, :]
        D = self.output_dim
        out_x = np.zeros((T, B, D), f)
        out_k = np.zeros((T, B, self.attention_size), f)
        out_w = np.zeros((T, B, self.encoded_input_dim), f)
        out_phi = np.zeros((T, B, U), f)
        out_pia = np.zeros((T, B, self.attention_size), f)
        out_pi = (np.zeros((T, B, D), f) if self.which_cost == 'MSE'
                  else np.zeros((T, B, self.k_gmm), f))
        for t in range(T):                                      # model.py:882-1036
            c1, g1 = cellb[0], gatb[0]
            c2, g2 = cellb[1], gatb[1]
            c3, g3 = cellb[2], gatb[2]
            ai, ag = self._fork('/inp_to_h1', ['rnn1_inputs', 'rnn1_gates'], w)
            c1 = c1 + ai; g1 = g1 + ag
            if self.weak_feedback:                              # model.py:899-908
                oc, og = self._fork('/out_to_h1', ['rnn1_inputs', 'rnn1_gates'], x)
                c1 = c1 + apply_norm(oc, LN); g1 = g1 + apply_norm(og, LN)
            if self.full_feedback:                              # model.py:910-924
                oc2, og2 = self._fork('/out_to_h2', ['rnn2_inputs', 'rnn2_gates'], x)
                oc3, og3 = self._fork('/out_to_h3', ['rnn3_inputs', 'rnn3_gates'], x)
                c2 = c2 + apply_norm(oc2, LN); c3 = c3 + apply_norm(oc3, LN)
                g2 = g2 + apply_norm(og2, LN); g3 = g3 + apply_norm(og3, LN)
            h1, _ = self._gru('/rnn1', c1, g1, h1)
            a, b, k, phi, w, _ = self._attention(
                h1, k, ctx, u, sharpening=self.sharpening_coeff, timing=self.timing_coeff)
            i2, g2i = self._fork('/inp_to_h2', ['rnn2_inputs', 'rnn2_gates'], w)
            i3, g3i = self._fork('/inp_to_h3', ['rnn3_inputs', 'rnn3_gates'], w)
            c2 = c2 + i2; g2 = g2 + g2i; c3 = c3 + i3; g3 = g3 + g3i
            p12 = [apply_norm(v, LN) for v in
                   self._fork('/h1_to_h2', ['rnn2_inputs', 'rnn2_gates'], h1)]
            p13 = [apply_norm(v, LN) for v in
                   self._fork('/h1_to_h3', ['rnn3_inputs', 'rnn3_gates'], h1)]
            h2, _ = self._gru('/rnn2', c2 + p12[0], g2 + p12[1], h2)
            p23 = [apply_norm(v, LN) for v in
                   self._fork('/h2_to_h3', ['rnn3_inputs', 'rnn3_gates'], h2)]
            h3, _ = self._gru('/rnn3', c3 + p13[0] + p23[0], g3 + p13[1] + p23[1], h3)
            ro = [apply_norm(self._lin('/h%d_to_readout' % (i + 1), hh), LN)
                  for i, hh in enumerate((h1, h2, h3))]          # model.py:992-999
            readout = ro[0] + ro[1] + ro[2]
            readout = readout + self._lin('/att_to_readout', w)  # model.py:1003
            if self.use_speaker:
                readout = readout + spk_readout                  # model.py:1005-1006
            if self.which_cost == 'MSE':                         # model.py:1010-1016
                x = self._lin('/readout_to_output', readout)
                if self.use_speaker:
                    x = x + spk_output
                pi_t = x
            else:                                                # model.py:1017-1033
                mu, sg_hat, co_hat = self._fork(
                    '/readout_to_output', ['gmm_mu', 'gmm_sigma', 'gmm_coeff'], readout)
                if self.use_speaker:
                    mu = mu + spk_output[0]; sg_hat = sg_hat + spk_output[1]
                    co_hat = co_hat + spk_output[2]
                sigma = np.exp(sg_hat - ft(self.sampling_bias)) + ft(self.epsilon)
                pi_t = softmax(co_hat * ft(1. + self.sampling_bias)) + ft(self.epsilon)
                x = sample_gmm(mu, sigma, pi_t, np.asarray(gmm_unis[t], f),
                               np.asarray(gmm_normals[t], f))
            out_x[t] = x; out_k[t] = k; out_w[t] = w
            out_phi[t] = phi; out_pia[t] = a; out_pi[t] = pi_t
        return [out_x, out_k, out_w, out_pi, out_phi, out_pia]


# --------------------------------------------------------------------------
# optimizer (train.py:100-108): StepClipping(10*grad_clip) then Adam(lr)
# --------------------------------------------------------------------------
class OracleAdamClip(object):
    """blocks CompositeRule([StepClipping(threshold), Adam(lr)]) + GradientDescent.

    Blocks >= 0.2 Adam defaults: beta1=0.9, beta2=0.999, epsilon=1e-8,
    decay_factor=1 (PARITY UNPINNED: the reference pins no Blocks version).
    """

    def __init__(self, shapes, learning_rate=1e-4, threshold=9.0, beta1=0.9,
                 beta2=0.999, epsilon=1e-8, dtype=np.float32):
        self.lr = learning_rate
        self.threshold = threshold
        self.b1, self.b2, self.eps = beta1, beta2, epsilon
        self.dtype = np.dtype(dtype)
        self.m = OrderedDict((n, np.zeros(s, self.dtype)) for n, s in shapes.items())
        self.v = OrderedDict((n, np.zeros(s, self.dtype)) for n, s in shapes.items())
        self.time = 0

    def step(self, params, grads):
        f = self.dtype.type
        norm = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
        mult = 1.0 if norm < self.threshold else self.threshold / norm
        self.time += 1
        t1 = self.time
        lr_t = self.lr * np.sqrt(1. - self.b2 ** t1) / (1. - self.b1 ** t1)
        for n in params:
            g = grads[n] * f(mult)
            self.m[n] = f(self.b1) * self.m[n] + f(1. - self.b1) * g
            self.v[n] = f(self.b2) * self.v[n] + f(1. - self.b2) * g * g
            params[n] = params[n] - f(lr_t) * self.m[n] / (np.sqrt(self.v[n]) + f(self.eps))
        return norm


# --------------------------------------------------------------------------
# host-side stop heuristic (sample.py:147-163)
# --------------------------------------------------------------------------
def stop_heuristic(phi, labels_length, num_steps):
    """phi: (T, U) for one utterance.  First t where phi[t, len] beats every
    phi[t, :len-1]; +40 frames, clamped to num_steps; falls back to num_steps."""
    try:
        cond = (phi[:, labels_length, None] > phi[:, :labels_length - 1]).all(axis=1)
        t = np.where(cond)[0][0]
        return int(min(num_steps, t + 40))
    except Exception:
        return int(num_steps)
