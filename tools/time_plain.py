"""Milestones of one launch of a plain (chunk / batched) table (debug tool, run under gpurun)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from parrot_b200 import Parrot, _lib

T = 128
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
for name in sys.argv[1:] or ['gC2', 'gC3', 'readout']:
    ms = C.c_float()
    tl = torch.zeros(148 * 16, dtype=torch.int64, device='cuda')
    _lib.check(lib.parrot_debug_time_table(h.ptr, name.encode(), 0, 0, 3, C.byref(ms), C.c_void_p(tl.data_ptr()), st))
    a = tl.cpu().numpy().reshape(148, 16).astype(np.float64)
    print('%s: %.1f us per launch' % (name, ms.value * 1e3))
    act = a[a[:, 0] > 0]
    t0 = act[:, 0].min()
    names = {0: 'entry', 1: 'prologue', 2: 'tma_done', 3: 'mma_done', 4: 'acc_ready(last job)', 5: 'tmem_read(last job)', 8: 'exit'}
    for i, n in names.items():
        col = act[:, i]; ok = col > 0
        if ok.any():
            d = (col[ok] - t0) / 1e3
            print('    %-22s n=%3d  min %7.2f  median %7.2f  max %7.2f us' % (n, ok.sum(), d.min(), np.median(d), d.max()))
