"""Host-side mirror of the reference ``model.py:Parrot`` API (model.py:250-1111).

Same constructor keywords, ``initialize``, ``symbolic_input_variables``,
``initial_states``, ``compute_cost`` (4-tuple), ``sample_model`` (6 time-major
arrays) -- but eager: there is no symbolic graph, every call runs hand-written
sm_100a kernels through the C ABI of ``libparrot_b200.so``.  PyTorch tensors are
storage only (flat parameter / gradient buffers, the device workspace, staging
of host inputs); no torch op is on the compute path.

Differences a reference user has to know:
  * ``compute_cost`` takes arrays (numpy or torch), not symbolic variables, and
    returns concrete values; ``updates`` is the list ``[(name, tensor)]`` of
    carried-state updates, already applied to the model (the reference's compiled
    function applies them too, model.py:786-791 / train.py:108);
  * RNG draws can be injected (``feedback_noise``, ``gmm_noise``) for parity
    tests; otherwise Philox streams are used (Theano's MRG31k3p streams are not
    reproduced, SURVEY hard part 7);
  * ``encoder_time_axis`` (0 = literal reference behaviour);
  * ``raw_output=True`` (sampleRNN coupling) raises NotImplementedError.
"""
import ctypes as C
from collections import OrderedDict, namedtuple

import numpy as np
import torch

from . import _lib
from ._lib import ParrotConfig

_DEFAULTS = OrderedDict([
    ('input_dim', 420), ('output_dim', 63), ('rnn_h_dim', 1024), ('readouts_dim', 1024),
    ('weak_feedback', False), ('full_feedback', False), ('feedback_noise_level', None),
    ('layer_norm', False), ('use_speaker', False), ('num_speakers', 21), ('speaker_dim', 128),
    ('which_cost', 'MSE'), ('k_gmm', 20), ('sampling_bias', 0), ('epsilon', 1e-5),
    ('num_characters', 43), ('attention_type', 'graves'), ('attention_size', 10),
    ('attention_alignment', 1.), ('sharpening_coeff', 1.), ('timing_coeff', 1.),
    ('encoder_type', None), ('encoder_dim', 128), ('raw_output', False)])

SymbolicInput = namedtuple('SymbolicInput', 'name dtype ndim')


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _Handle(object):
    """One ``parrot_model`` (fixed batch, frames, text length) + its workspace."""

    def __init__(self, model, B, T, U, sampling):
        lib = _lib.load()
        self.cfg = model._make_cfg(B, T, U, sampling)
        nbytes = C.c_size_t()
        _lib.check(lib.parrot_workspace_bytes(C.byref(self.cfg), C.byref(nbytes)))
        # 1 KiB-aligned device workspace, owned here (the library never allocates device memory)
        self.ws = torch.empty(nbytes.value + 1024, dtype=torch.uint8, device=model.device)
        self._shift = (-self.ws.data_ptr()) % 1024
        self.ws_ptr = self.ws.data_ptr() + self._shift
        self.ws_bytes = nbytes.value
        self.ptr = C.c_void_p()
        stream = torch.cuda.current_stream(model.device).cuda_stream
        _lib.check(lib.parrot_create(C.byref(self.cfg), _ptr(model.flat_params), _ptr(model.flat_grads),
                                     C.c_void_p(self.ws_ptr), C.c_size_t(self.ws_bytes),
                                     C.c_void_p(stream), C.byref(self.ptr)))
        self.B, self.T, self.U, self.sampling = B, T, U, sampling
        self.lib = lib

    def buffer(self, name, shape, dtype=torch.float32):
        off = C.c_int64(); n = C.c_int64()
        _lib.check(self.lib.parrot_buffer_info(self.ptr, name.encode(), C.byref(off), C.byref(n)))
        numel = int(np.prod(shape))
        assert numel <= n.value, (name, shape, n.value)
        start = self._shift + off.value
        return self.ws[start:start + numel * 4].view(dtype).view(*shape)

    def __del__(self):
        try:
            if self.ptr:
                self.lib.parrot_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Parrot(object):
    """Drop-in for ``model.Parrot`` on the attention-RNN hot path (model.py:250)."""

    def __init__(self, device=None, gemm_impl='tcgen05', encoder_time_axis=0, **kwargs):
        cfg = OrderedDict(_DEFAULTS)
        # Blocks-only keywords the reference passes (train.py:74-77)
        self.weights_init = kwargs.pop('weights_init', None)
        self.biases_init = kwargs.pop('biases_init', None)
        self.name = kwargs.pop('name', 'parrot')
        unknown = set(kwargs) - set(cfg)
        if unknown:
            raise TypeError('unknown Parrot arguments: %s' % sorted(unknown))
        cfg.update(kwargs)
        assert cfg['encoder_type'] in (None, 'bidirectional')          # model.py:209
        assert cfg['which_cost'] in ('MSE', 'GMM')
        assert cfg['attention_type'] in ('graves', 'softmax')
        if cfg['raw_output']:
            raise NotImplementedError('raw_output (sampleRNN coupling) is outside the hot path, SURVEY 8f')
        if cfg['full_feedback']:
            cfg['weak_feedback'] = True                                # model.py:485
        if cfg['encoder_type'] is None:
            assert cfg['num_characters'] == cfg['input_dim']           # model.py:224
        self.__dict__.update(cfg)
        self.encoded_input_dim = (2 * self.encoder_dim if self.encoder_type == 'bidirectional'
                                  else self.input_dim)                 # model.py:302-307
        self.encoder_time_axis = encoder_time_axis
        self.gemm_impl = {'tcgen05': 0, 'simt': 1}[gemm_impl]
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('parrot_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        self._handles = OrderedDict()          # (B, T, U, sampling) -> _Handle, least recently used first
        self.max_handles = 4                   # each handle owns a full workspace (GBs at base size): LRU eviction
        self._last = None
        self._gen = None                       # torch.Generator of the default noise draws, see seed_noise()
        self.flat_params = None
        layout, total = _lib.param_layout(self._make_cfg(1, 1, 1, 0))
        self._layout, self.num_floats = layout, total

    # ------------------------------------------------------------------ config
    def _make_cfg(self, B, T, U, sampling):
        return ParrotConfig(
            input_dim=self.input_dim, output_dim=self.output_dim, rnn_h_dim=self.rnn_h_dim,
            readouts_dim=self.readouts_dim, weak_feedback=int(bool(self.weak_feedback)),
            full_feedback=int(bool(self.full_feedback)), layer_norm=int(bool(self.layer_norm)),
            use_speaker=int(bool(self.use_speaker)), num_speakers=self.num_speakers,
            speaker_dim=self.speaker_dim, which_cost=0 if self.which_cost == 'MSE' else 1,
            k_gmm=self.k_gmm, num_characters=self.num_characters,
            attention_type=0 if self.attention_type == 'graves' else 1,
            attention_size=self.attention_size, encoder_type=0 if self.encoder_type is None else 1,
            encoder_dim=self.encoder_dim, encoder_time_axis=self.encoder_time_axis,
            sampling_bias=float(self.sampling_bias), epsilon=float(self.epsilon),
            attention_alignment=float(self.attention_alignment),
            sharpening_coeff=float(self.sharpening_coeff), timing_coeff=float(self.timing_coeff),
            batch_size=B, seq_len=T, text_len=U, gemm_impl=self.gemm_impl, sampling=int(sampling))

    # -------------------------------------------------------------- parameters
    def _allocate(self):
        if self.flat_params is not None:
            return
        n = self.num_floats
        self.flat_params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.flat_grads = torch.zeros(n + 1, dtype=torch.float32, device=self.device)  # + sum(mask)
        self.parameters = OrderedDict()
        self.gradients = OrderedDict()
        for name, off, shape in self._layout:
            k = int(np.prod(shape))
            self.parameters[name] = self.flat_params[off:off + k].view(*shape)
            self.gradients[name] = self.flat_grads[off:off + k].view(*shape)

    def initialize(self, seed=0, std=0.01, gain=None):
        """train.py:30-31: W ~ IsotropicGaussian(0.01), b = 0; GRU initial_state and initial_w = 0.

        ``gain`` (not in the reference) draws W ~ N(0, (gain/sqrt(fan_in))^2) instead: the
        synthetic 'trained-like' parameter set of SURVEY 8d."""
        self._allocate()
        rng = np.random.default_rng(seed)
        host = np.zeros(self.num_floats, np.float32)
        for name, off, shape in self._layout:
            if name.endswith('.b') or name.endswith('.initial_state') or name.endswith('.initial_w'):
                continue
            sd = std if gain is None else gain / np.sqrt(shape[0])
            if gain is not None and ('lookuptable' in name or 'embed_label' in name):
                sd = 1.0
            k = int(np.prod(shape))
            host[off:off + k] = (rng.standard_normal(shape) * sd).astype(np.float32).ravel()
        self.flat_params.copy_(torch.from_numpy(host))
        self.mark_dirty()
        self.seed_noise(seed)
        return self

    def set_parameter_values(self, values):
        """``blocks.model.Model.set_parameter_values`` equivalent: dict brick-path -> array."""
        self._allocate()
        for name, arr in values.items():
            if name.startswith('__'):        # optimizer state riding along in a train.py checkpoint
                continue
            p = self.parameters[name]
            p.copy_(torch.as_tensor(np.asarray(arr, np.float32)).view_as(p))
        self.mark_dirty()

    def get_parameter_values(self):
        return OrderedDict((n, p.detach().cpu().numpy().copy()) for n, p in self.parameters.items())

    def seed_noise(self, seed, rank=None):
        """Seed the generator of the default feedback-noise / GMM draws from (seed, rank): runs are reproducible from
        ``--seed`` and data-parallel ranks draw DIFFERENT noise for their shards (a global default generator would give
        every rank the same stream)."""
        import os
        if rank is None:
            rank = int(os.environ.get('RANK', '0'))
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed((int(seed) * 1000003 + 7919 * int(rank) + 17) % (2 ** 63 - 1))
        return self

    def _generator(self):
        if self._gen is None:
            self.seed_noise(0)
        return self._gen

    def mark_dirty(self):
        """Tell the device handles that the fp32 parameters changed (operand planes are re-derived)."""
        for h in self._handles.values():
            h.lib.parrot_mark_params_dirty(h.ptr)

    # ---------------------------------------------------------------- handles
    def _handle(self, B, T, U, sampling=False):
        self._allocate()
        key = (B, T, U, bool(sampling))
        h = self._handles.get(key)
        if h is None:
            # least-recently-used eviction: U (padded text length) changes from batch to batch on real data and every
            # handle owns packed weight planes + all activation stashes, so an unbounded cache runs out of memory
            while len(self._handles) >= max(1, self.max_handles):
                old_key = next(k for k in self._handles if self._handles[k] is not self._last or len(self._handles) == 1)
                old = self._handles.pop(old_key)
                if old is self._last:
                    self._last = None
                old.lib.parrot_destroy(old.ptr)
                old.ptr = None
                del old
            h = _Handle(self, B, T, U, sampling)
            self._handles[key] = h
        else:
            self._handles.move_to_end(key)
        # carried state follows the MODEL, not the handle (model.py:534-546: one set of last_* shared variables):
        # whenever the handle changes, the state of the previous training call moves with it
        prev = self._last
        if (not sampling and prev is not None and prev is not h and not prev.sampling and prev.B == B
                and prev.ptr):
            for nm, shp in self._state_shapes(B):
                h.buffer(nm, shp).copy_(prev.buffer(nm, shp))
        return h

    def _state_shapes(self, B):
        H, A, Cc = self.rnn_h_dim, self.attention_size, self.encoded_input_dim
        return [('last_h1', (B, H)), ('last_h2', (B, H)), ('last_h3', (B, H)),
                ('last_k', (B, A)), ('last_w', (B, Cc))]

    def get_state(self):
        """Copies of the carried TBPTT state (last_h1..3, last_k, last_w; model.py:534-546) or None."""
        h = self._last
        if h is None or h.sampling:
            return None
        return (h, [h.buffer(nm, shp).clone() for nm, shp in self._state_shapes(h.B)])

    def set_state(self, state):
        if state is None:
            return
        h, vals = state
        for (nm, shp), v in zip(self._state_shapes(h.B), vals):
            h.buffer(nm, shp).copy_(v)
        self._last = h

    def _dev(self, x, dtype):
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()
        a = np.ascontiguousarray(np.asarray(x))
        t = torch.from_numpy(a)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True)

    # --------------------------------------------------------- reference API
    def symbolic_input_variables(self):
        """model.py:508-527: kept as a descriptor of names / dtypes / ranks."""
        features = SymbolicInput('features', 'float32', 3)
        features_mask = SymbolicInput('features_mask', 'float32', 2)
        labels = SymbolicInput('labels', 'int32', 2)
        labels_mask = SymbolicInput('labels_mask', 'float32', 2)
        start_flag = SymbolicInput('start_flag', 'float32', 0)
        speaker = SymbolicInput('speaker_index', 'int32', 2) if self.use_speaker else None
        raw_sequence = None
        return features, features_mask, labels, labels_mask, speaker, start_flag, raw_sequence

    def initial_states(self, batch_size):
        """model.py:529-549, same 10-tuple order."""
        self._allocate()
        B = batch_size
        P = self.parameters
        ih = [P['/parrot/rnn%d.initial_state' % i][None, :].expand(B, -1) for i in (1, 2, 3)]
        iw = P['/parrot.initial_w'][None, :].expand(B, -1)
        ik = torch.zeros(B, self.attention_size, device=self.device)
        h = self._last if (self._last is not None and self._last.B == B and not self._last.sampling) else None
        if h is not None:
            last = [h.buffer(nm, shp) for nm, shp in self._state_shapes(B)]
        else:
            last = [torch.zeros(*shp, device=self.device) for _, shp in self._state_shapes(B)]
        return ih[0], last[0], ih[1], last[1], ih[2], last[2], iw, last[4], ik, last[3]

    def compute_cost(self, features, features_mask, labels, labels_mask, speaker, start_flag,
                     batch_size, raw_audio=None, feedback_noise=None, noise_level=None,
                     gmm_noise=None):
        """model.py:552-824.  Returns ``(cost, updates, attention_vars, cost_raw)``.

        features (T+1, B, D), features_mask (T+1, B), labels (B, U) int, labels_mask (B, U),
        speaker (B, 1) int or None, start_flag scalar.  attention_vars =
        ``[next_x, k, w, coeff, phi, pi_att]`` (model.py:822), time-major torch tensors that
        alias the device workspace (valid until the next call).
        """
        if speaker is None:
            assert not self.use_speaker                                  # model.py:556-557
        f = self._dev(features, torch.float32)
        fm = self._dev(features_mask, torch.float32)
        lm = self._dev(labels_mask, torch.float32)
        if self.encoder_type is None:
            lab = self._dev(labels, torch.float32)
        else:
            lab = self._dev(labels, torch.int32)
        spk = self._dev(speaker, torch.int32)
        T = f.shape[0] - 1
        B = batch_size
        assert f.shape[1] == B and lab.shape[0] == B
        U = lm.shape[1]
        h = self._handle(B, T, U, False)
        noise = None
        level = 0.0
        if self.feedback_noise_level:                                    # truthiness gate (hazard H6)
            level = float(self.feedback_noise_level if noise_level is None else noise_level)
            if feedback_noise is None:
                feedback_noise = torch.randn(T, B, self.output_dim, device=self.device, generator=self._generator())
            noise = self._dev(feedback_noise, torch.float32)
        unis = normals = None
        if self.which_cost == 'GMM':
            if gmm_noise is None:
                gmm_noise = (torch.rand(T, B, device=self.device, generator=self._generator()),
                             torch.randn(T, B, self.output_dim, device=self.device, generator=self._generator()))
            unis = self._dev(gmm_noise[0], torch.float32)
            normals = self._dev(gmm_noise[1], torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        cost4 = h.buffer('cost', (4,))
        self._keep = (f, fm, lab, lm, spk, noise, unis, normals)         # keep device inputs alive for backward
        _lib.check(h.lib.parrot_compute_cost(
            h.ptr, _ptr(f), _ptr(fm), _ptr(lab), _ptr(lm), _ptr(spk), C.c_float(float(start_flag)),
            _ptr(noise), C.c_float(level), _ptr(unis), _ptr(normals), None, C.c_void_p(stream)))
        self._last = h
        D, A, Cc, H = self.output_dim, self.attention_size, self.encoded_input_dim, self.rnn_h_dim
        k = h.buffer('kappa', (T + 1, B, A))[1:]
        w = h.buffer('w', (T + 1, B, Cc))[1:]
        phi = h.buffer('phi', (T, B, U))
        pi_att = h.buffer('ab', (T, B, 2 * A))[:, :, :A]
        if self.which_cost == 'MSE':
            predicted = h.buffer('pred', (T, B, D))
            next_x, coeff = predicted, predicted                         # model.py:762-764
        else:
            next_x = h.buffer('next_x', (T, B, D))
            coeff = h.buffer('dpred', (T, B, self.k_gmm))               # parked there until backward()
        updates = [(nm, h.buffer(nm, shp)) for nm, shp in self._state_shapes(B)]
        cost = cost4[0]
        self.cost_terms = cost4   # [cost, sum(cost*mask), sum(mask), 1/(sum(mask)+1e-5)]
        return cost, updates, [next_x, k, w, coeff, phi, pi_att], None

    def backward(self, unnormalised=False):
        """Gradient of the last ``compute_cost`` wrt every parameter into ``flat_grads``
        (what theano.grad computes inside blocks GradientDescent, train.py:103-107).
        ``flat_grads[-1]`` receives sum(mask).  Returns the OrderedDict of gradient views."""
        h = self._last
        assert h is not None and not h.sampling, 'call compute_cost first'
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(h.lib.parrot_backward(h.ptr, int(bool(unnormalised)), C.c_void_p(stream)))
        return self.gradients

    def sample_model(self, labels_tr, labels_mask_tr, features_mask_tr, speaker_tr,
                     num_samples, num_steps, gmm_noise=None, seed=0, as_numpy=True):
        """model.py:1061-1083.  Returns ``[x, k, w, pi, phi, pi_att]``, time-major.
        ``features_mask_tr`` is unused, as in the reference."""
        lm = self._dev(labels_mask_tr, torch.float32)
        lab = self._dev(labels_tr, torch.float32 if self.encoder_type is None else torch.int32)
        spk = self._dev(speaker_tr, torch.int32) if self.use_speaker else None
        B, T, U = num_samples, num_steps, lm.shape[1]
        assert lab.shape[0] == B
        h = self._handle(B, T, U, True)
        unis = normals = None
        if gmm_noise is not None:
            unis = self._dev(gmm_noise[0], torch.float32)
            normals = self._dev(gmm_noise[1], torch.float32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(h.lib.parrot_sample_scan(h.ptr, _ptr(lab), _ptr(lm), _ptr(spk), _ptr(unis), _ptr(normals),
                                            C.c_uint64(seed), C.c_void_p(stream)))
        D, A, Cc = self.output_dim, self.attention_size, self.encoded_input_dim
        out = [h.buffer('samp_x', (T, B, D)),
               h.buffer('kappa', (T + 1, B, A))[1:],
               h.buffer('w', (T + 1, B, Cc))[1:],
               h.buffer('samp_pi', (T, B, self.k_gmm if self.which_cost == 'GMM' else D)),
               h.buffer('phi', (T, B, U)),
               h.buffer('ab', (T, B, 2 * A))[:, :, :A]]
        if as_numpy:
            return [o.detach().cpu().numpy().copy() for o in out]
        return out

    def sample_using_input(self, data_tr, num_samples):
        """model.py:1085-1111 is broken in the reference (unpacks 3 of compute_cost's 4 values,
        SURVEY D8); kept as the evidently intended teacher-forced pass."""
        cost, updates, attention_vars, _ = self.compute_cost(
            data_tr['features'], data_tr['features_mask'], data_tr['labels'], data_tr['labels_mask'],
            data_tr.get('speaker_index'), data_tr.get('start_flag', 1.0), num_samples)
        return [v.detach().cpu().numpy().copy() for v in attention_vars]
