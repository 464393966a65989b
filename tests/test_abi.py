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


def _workspace_bytes(cfg, B, T, U, sampling=0):
    L = _lib()
    from parrot_b200.model import Parrot, _DEFAULTS
    m = Parrot.__new__(Parrot)
    d = dict(_DEFAULTS); d.update(cfg)
    if d['full_feedback']:
        d['weak_feedback'] = True
    m.__dict__.update(d)
    m.encoder_time_axis = 0; m.gemm_impl = 0
    c = m._make_cfg(B, T, U, sampling)
    n = C.c_size_t()
    L.check(L.load().parrot_workspace_bytes(C.byref(c), C.byref(n)))
    return n.value


def test_dry_table_build_for_group_partitions(monkeypatch):
    """parrot_workspace_bytes runs the whole host-side build (planes, tensor-map slots, job tables of the grouped scans,
    split-K scratch regions, TMEM residency assignment) without touching a GPU: every partition of the 148 CTAs into
    layer groups must produce tables (a group never gets more split jobs than CTAs), an invalid partition falls back to
    the default, and the per-table scratch regions make the grouped build larger than nothing else does."""
    base = dict(input_dim=420, output_dim=63, rnn_h_dim=1024, readouts_dim=1024, weak_feedback=True,
                which_cost='MSE', num_characters=43, attention_type='graves', attention_size=10,
                attention_alignment=0.15, encoder_type='bidirectional', encoder_dim=128)
    for k in ('PARROT_GROUPS_F', 'PARROT_GROUPS_B', 'PARROT_TC'):
        monkeypatch.delenv(k, raising=False)
    default = _workspace_bytes(base, 64, 800, 128)
    assert default > 10 * 2 ** 30                     # 13 GB of stashes and planes at the benchmarked size
    for part in ('100,24,24', '64,42,42', '64,36,48', '32,16,16', '50,49,49'):
        monkeypatch.setenv('PARROT_GROUPS_F', part)
        monkeypatch.setenv('PARROT_GROUPS_B', part)
        assert _workspace_bytes(base, 64, 800, 128) > 0, part
    monkeypatch.setenv('PARROT_GROUPS_F', '100,100,100')      # more than 148 CTAs: ignored
    monkeypatch.setenv('PARROT_GROUPS_B', 'nonsense')
    assert _workspace_bytes(base, 64, 800, 128) == default
    for k in ('PARROT_GROUPS_F', 'PARROT_GROUPS_B'):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv('PARROT_TC', '8')
    assert _workspace_bytes(base, 64, 800, 128) > 0
    monkeypatch.delenv('PARROT_TC', raising=False)
    # small and odd problems, GMM / speaker / full feedback, the sampling handle
    tiny = dict(util.TINY, which_cost='GMM', use_speaker=True, full_feedback=True)
    assert _workspace_bytes(tiny, 5, 7, 9) > 0
    assert _workspace_bytes(tiny, 5, 7, 9, sampling=1) > 0
    assert _workspace_bytes(dict(util.TINY, layer_norm=True), 4, 12, 6) > 0
