"""Per-barrier timeline of the persistent backward sweep (debug tool, run under gpurun)."""
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
bars = 3 * T + 8
st = torch.zeros(148 * bars * 2, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(st.data_ptr()), bars + (1 << 20))
m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
m.backward()
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
s = st.cpu().numpy().reshape(148, bars, 2).astype(np.float64)
t0 = s[s > 0].min()
s = np.where(s > 0, (s - t0) / 1e3, np.nan)
names = ['att_bwd+pre', 'bwd1 (d r*h)', 'bwd2 (dgrads)']
for k in range(3):
    idx = [3 * t + k for t in range(5, T - 3)]
    passed = s[:, idx, 0]; arr = s[:, idx, 1]
    work = arr - passed
    print('%-14s work median %.2f us  max-CTA median %.2f us ; phase span %.2f us' % (
        names[k], np.nanmedian(work), np.nanmedian(np.nanmax(work, axis=0)),
        np.nanmedian(np.nanmax(arr, axis=0) - np.nanmin(passed, axis=0))))
print('tick period %.2f us' % np.nanmedian(np.diff(s[0, 0::3, 0])[5:T - 4]))
