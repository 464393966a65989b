"""Micro-timing of the engine tables at the base configuration (debug tool, run under gpurun)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from parrot_b200 import Parrot, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 40
cfg = dict(bench.BASE)
B, U = 64, 128
m = Parrot(**cfg); m.initialize(seed=0)
bt = bench.make_batch(cfg, B, T, U, seed=1)
for _ in range(2):
    m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
    m.backward()
torch.cuda.synchronize()
h = m._last
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def tt(name, tick, rev=0, reps=50, timeline=False):
    ms = C.c_float()
    tl = torch.zeros(148 * 16, dtype=torch.int64, device='cuda') if timeline else None
    _lib.check(lib.parrot_debug_time_table(h.ptr, name.encode(), tick, rev, reps, C.byref(ms),
                                           C.c_void_p(tl.data_ptr()) if timeline else None, st))
    return ms.value * 1e3, (tl.cpu().numpy().reshape(148, 16) if timeline else None)
for name, rev in (('fwdA', 0), ('fwdB', 0), ('bwd1', 1), ('bwd2', 1)):
    us_skip, _ = tt(name, -100, rev)
    us, tl = tt(name, 10, rev, timeline=True)
    print('%-5s all-skipped launch %.2f us ; full launch %.2f us' % (name, us_skip, us))
    act = tl[tl[:, 0] > 0]
    t0 = act[:, 0].min()
    names = ['entry', 'prologue', 'tma_done', 'mma_done', 'acc_ready', 'part_written', 'arrived', 'epi_done', 'exit']
    for i, n in enumerate(names):
        col = act[:, i]
        ok = col > 0
        if ok.any():
            d = (col[ok] - t0) / 1e3
            print('    %-13s n=%3d  min %7.2f  median %7.2f  max %7.2f us' % (n, ok.sum(), d.min(), np.median(d), d.max()))
for name in ('readout', 'wgrad', 'dread', 'gC2', 'gC3', 'chunkF'):
    us, _ = tt(name, 0, 0, reps=5)
    print('%-8s %.1f us' % (name, us))
