"""GPU unit tests of individual kernels through the C ABI (K6 GEMM engine, K7 attention step, K14 optimizer)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import parrot_oracle as O
from tests import util

pytestmark = pytest.mark.gpu


def _gemm(A, B, impl):
    from parrot_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    dA = torch.from_numpy(A).cuda(); dB = torch.from_numpy(B).cuda()
    dC = torch.zeros(M, N, device='cuda')
    nbytes = lib.parrot_gemm_nt_workspace_bytes(M, N, K)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device='cuda')
    shift = (-ws.data_ptr()) % 1024
    _lib.check(lib.parrot_gemm_nt(C.c_void_p(dA.data_ptr()), C.c_void_p(dB.data_ptr()), C.c_void_p(dC.data_ptr()),
                                  M, N, K, impl, C.c_void_p(ws.data_ptr() + shift), C.c_size_t(nbytes),
                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return dC.cpu().numpy()


@pytest.mark.parametrize('impl', [1, 0], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('shape', [(128, 128, 64), (128, 64, 256), (256, 200, 192), (100, 37, 70),
                                   (384, 512, 1024)])
def test_gemm_nt_matches_float64(impl, shape):
    """bf16x3 engine vs float64: error must be at the 2^-16 operand-split level, far below bf16 (4e-3)."""
    M, N, K = shape
    rng = np.random.default_rng(0)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = _gemm(A, B, impl)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    assert err < 5e-5, err
    # and it must agree with the oracle's emulation of the same operand split
    emu = O.emulate_mma(A, B.T.copy(), 'bf16x3')
    assert np.abs(out - emu).max() / np.abs(ref).max() < 2e-5


def test_gemm_tc_matches_simt_large():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((512, 2304)).astype(np.float32)
    B = rng.standard_normal((300, 2304)).astype(np.float32)
    a = _gemm(A, B, 0); b = _gemm(A, B, 1)
    assert np.abs(a - b).max() / np.abs(b).max() < 5e-5


@pytest.mark.parametrize('impl', [1, 0], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('variant', ['tile16', 'tile64', 'tile256', 'slots3d_16', 'slots3d_64', 'twoseg',
                                     'slots3d_16_twoseg'])
def test_gemm_engine_variants(impl, variant):
    """The scan uses sample tiles of the padded batch (16..256 columns), slot-indexed 3-D TMA maps and
    multi-segment jobs; exercise each against float64."""
    flags = impl
    M, N, K = 256, 200, 320
    if variant.startswith('tile'):
        t = int(variant[4:]); flags |= (255 if t == 256 else t) << 8
    if 'slots3d' in variant:
        t = int(variant.split('_')[1]); flags |= (t << 8) | (1 << 16); N = t - 3
    if 'twoseg' in variant:
        flags |= 1 << 17
    rng = np.random.default_rng(2)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    out = _gemm(A, B, flags)
    assert np.abs(out - ref).max() / np.abs(ref).max() < 5e-5


@pytest.mark.parametrize('att', ['graves', 'softmax'])
def test_attention_step_matches_oracle(att):
    """parrot_attention_step (K7) vs oracle._attention: w within 1e-5 rel, argmax(phi) bit-exact."""
    from parrot_b200 import _lib
    from parrot_b200.model import Parrot
    cfg = dict(util.TINY, attention_type=att, rnn_h_dim=128)
    B, U = 8, 40
    orc = util.make_oracle(cfg, gain=0.5)
    rng = np.random.default_rng(3)
    H, A, Cc = cfg['rnn_h_dim'], cfg['attention_size'], 2 * cfg['encoder_dim']
    h1 = np.tanh(rng.standard_normal((B, H))).astype(np.float32)
    ctx = rng.standard_normal((B, U, Cc)).astype(np.float32)
    k_prev = np.abs(rng.standard_normal((B, A))).astype(np.float32) * 10
    u = np.arange(U, dtype=np.float32)[None, None, :]
    a, b, k, phi, w, hats = orc._attention(h1, k_prev, ctx, u)
    m = Parrot(**cfg)
    ccfg = m._make_cfg(B, 1, U, 0)
    wT = np.concatenate([orc.params['/parrot/h1_to_att/fork_%s.W' % n].T for n in ('alpha', 'beta', 'kappa')], 0)
    bt = np.concatenate([orc.params['/parrot/h1_to_att/fork_%s.b' % n] for n in ('alpha', 'beta', 'kappa')], 0)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()
    dh1, dwT, dbt, dctx, dk = d(h1), d(wT), d(bt), d(ctx), d(k_prev)
    ko = torch.zeros(B, A, device='cuda'); wo = torch.zeros(B, Cc, device='cuda')
    po = torch.zeros(B, U, device='cuda'); ab = torch.zeros(B, 2 * A, device='cuda')
    eo = torch.zeros(B, 3 * A, device='cuda')
    lib = _lib.load()
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.parrot_attention_step(C.byref(ccfg), p(dh1), p(dwT), p(dbt), p(dctx), p(dk), p(ko), p(wo), p(po),
                                         p(ab), p(eo), 1, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert util.rel_err(ko.cpu().numpy(), k) < 1e-5
    assert util.rel_err(po.cpu().numpy(), phi) < 1e-4
    assert util.rel_err(wo.cpu().numpy(), w) < 1e-4
    assert (po.cpu().numpy().argmax(-1) == phi.argmax(-1)).all()


def test_adam_clip_matches_oracle():
    from parrot_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = 100003
    p0 = rng.standard_normal(n).astype(np.float32)
    shapes = {'x': (n,)}
    opt = O.OracleAdamClip(shapes, learning_rate=1e-3, threshold=9.0)
    params = {'x': p0.copy()}
    dp = torch.from_numpy(p0.copy()).cuda()
    dm = torch.zeros(n, device='cuda'); dv = torch.zeros(n, device='cuda')
    stats = torch.zeros(4, device='cuda'); scratch = torch.zeros(1024, dtype=torch.float64, device='cuda')
    for step in range(1, 4):
        g = (rng.standard_normal(n) * (0.1 if step == 2 else 0.01)).astype(np.float32)
        norm = opt.step(params, {'x': g})
        dg = torch.from_numpy(g).cuda()
        _lib.check(lib.parrot_adam_clip_step(
            C.c_void_p(dp.data_ptr()), C.c_void_p(dg.data_ptr()), C.c_void_p(dm.data_ptr()),
            C.c_void_p(dv.data_ptr()), n, 1.0, None, 9.0, 1e-3, 0.9, 0.999, 1e-8, step,
            C.c_void_p(stats.data_ptr()), C.c_void_p(scratch.data_ptr()),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        assert abs(stats[0].item() - norm) / norm < 1e-5
        assert util.rel_err(dp.cpu().numpy(), params['x']) < 1e-5
