"""CPU tests of the C-ABI boundary: the library builds, loads, exports every symbol the header declares,
and its parameter inventory matches the oracle's restatement of the reference constructor."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import parrot_oracle as O
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from parrot_b200 import _lib
    return _lib


def test_library_builds_and_exports_every_declared_symbol():
    L = _lib()
    lib = L.load()
    header = open(os.path.join(ROOT, 'include', 'parrot_b200.h')).read()
    declared = set(re.findall(r'\b(parrot_[a-z_0-9]+)\s*\(', header))
    declared -= {'parrot_config', 'parrot_model'}
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.parrot_abi_version() == 1


def test_config_struct_matches_header_field_order():
    L = _lib()
    header = open(os.path.join(ROOT, 'include', 'parrot_b200.h')).read()
    body = header.split('typedef struct parrot_config {')[1].split('} parrot_config;')[0]
    fields = re.findall(r'(?:int32_t|float)\s+([a-z_]+);', body)
    assert fields == [f[0] for f in L.ParrotConfig._fields_]


@pytest.mark.parametrize('over', [
    dict(), dict(which_cost='GMM', use_speaker=True, full_feedback=True), dict(weak_feedback=True),
    dict(encoder_type=None, input_dim=43, num_characters=43)])
def test_param_inventory_matches_oracle(over):
    L = _lib()
    from parrot_b200.model import Parrot
    cfg = dict(util.TINY, **over)
    orc = O.OracleParrot(**cfg)
    m = Parrot.__new__(Parrot)
    # build the config without touching CUDA
    from parrot_b200.model import _DEFAULTS
    d = dict(_DEFAULTS); d.update(cfg)
    if d['full_feedback']:
        d['weak_feedback'] = True
    m.__dict__.update(d)
    m.encoder_time_axis = 0; m.gemm_impl = 0
    layout, total = L.param_layout(m._make_cfg(4, 5, 6, 0))
    assert [n for n, _, _ in layout] == list(orc.shapes.keys())
    for n, off, shape in layout:
        assert tuple(shape) == tuple(orc.shapes[n]), n
        assert off % 64 == 0
    assert total >= sum(int(np.prod(s)) for s in orc.shapes.values())


def test_workspace_query_and_config_errors():
    L = _lib()
    lib = L.load()
    cfg = L.ParrotConfig(input_dim=24, output_dim=63, rnn_h_dim=64, readouts_dim=64, num_characters=43,
                         attention_size=10, encoder_type=1, encoder_dim=64, k_gmm=20, num_speakers=5,
                         speaker_dim=16, epsilon=1e-5, attention_alignment=1.0, sharpening_coeff=1.0,
                         timing_coeff=1.0, batch_size=8, seq_len=50, text_len=16)
    n = C.c_size_t()
    assert lib.parrot_workspace_bytes(C.byref(cfg), C.byref(n)) == 0 and n.value > 0
    small = n.value
    cfg.seq_len = 100
    assert lib.parrot_workspace_bytes(C.byref(cfg), C.byref(n)) == 0 and n.value > small
    plain = n.value
    cfg.layer_norm = 1; cfg.weak_feedback = 1
    assert lib.parrot_workspace_bytes(C.byref(cfg), C.byref(n)) == 0 and n.value > plain   # pre-norm stashes
    cfg.sampling = 1
    assert lib.parrot_workspace_bytes(C.byref(cfg), C.byref(n)) == 0
    cfg.layer_norm = 0; cfg.sampling = 0; cfg.rnn_h_dim = 50
    assert lib.parrot_workspace_bytes(C.byref(cfg), C.byref(n)) != 0


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under parrot_b200/ or the entry scripts may touch it."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, 'parrot_b200')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle', src, re.M):
                    bad.append(f)
    for f in ('train.py', 'sample.py', 'generate.py'):
        p = os.path.join(ROOT, f)
        if os.path.exists(p) and re.search(r'^\s*(from|import)\s+oracle', open(p).read(), re.M):
            bad.append(f)
    assert not bad, bad


def test_parrot_without_cuda_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from parrot_b200.model import Parrot
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        Parrot(**util.TINY)


def test_comm_single_rank_is_the_identity_without_nccl_or_gpu():
    """parrot_comm_* (SURVEY 8b C1): a 1-rank communicator needs neither NCCL nor a device; allreduce leaves the buffer
    untouched, info reports (1, 0).  The N-rank path is covered on hardware by tests/test_gpu_multirank.py."""
    import torch
    from parrot_b200 import parallel
    c = parallel.Comm(1, 0)
    x = torch.arange(16, dtype=torch.float32)
    y = c.allreduce(x.clone())
    assert torch.equal(x, y)
    assert c.info()[:2] == (1, 0)
    lib = _lib().load()
    # bad arguments come back as error codes, never as exceptions across the C boundary
    ptr = C.c_void_p()
    assert lib.parrot_comm_init(2, 5, None, C.byref(ptr)) != 0
    assert b'rank' in lib.parrot_last_error()
