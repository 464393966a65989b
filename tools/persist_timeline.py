"""Per-barrier timeline of the persistent forward scan (debug tool, run under gpurun)."""
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
torch.cuda.synchronize()
h = m._last
lib = _lib.load()
bars = 4 * T + 8
st = torch.zeros(148 * bars * 2, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(st.data_ptr()), bars)
m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
s = st.cpu().numpy().reshape(148, bars, 2).astype(np.float64)
t0 = s[s > 0].min()
s = (s - t0) / 1e3
names = ['A(gates)', 'B(cand)', 'att-proj', 'att-window']
# steady-state ticks 5..T-1: barrier index = 4*tick + k
for k in range(4):
    idx = [4 * t + k for t in range(5, T - 2)]
    passed = s[:, idx, 0]; arr = s[:, idx, 1]
    work = arr - passed                       # this CTA's time inside the phase
    last_arr = arr.max(axis=0)                # when the slowest CTA arrived
    nxt = np.array([s[:, i + 1, 0].min() for i in idx])   # earliest CTA past the next barrier
    print('%-10s work median %.2f us  max-CTA median %.2f us ; barrier latency (last arrival -> first pass) %.2f us ; '
          'phase span %.2f us' % (names[k], np.median(work), np.median(work.max(axis=0)),
                                  np.median(nxt - last_arr), np.median(last_arr - passed.min(axis=0))))
tick_span = np.median(np.diff(s[0, 0::4, 0])[5:T - 3])
print('tick period %.2f us' % tick_span)

# ---- intra-phase milestones of one steady-state tick
tl = torch.zeros(3 * 148 * 16, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(tl.data_ptr()), -20)
m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
tl = tl.cpu().numpy().reshape(3, 148, 16).astype(np.float64)[1:]
names = {0: 'enter_phase', 1: 'barrier_passed', 2: 'tma_done', 3: 'mma_done', 4: 'acc_ready', 5: 'part_written',
         6: 'all_arrived', 7: 'epi_done'}
for ph in range(2):
    a = tl[ph]
    t0 = a[a > 0].min()
    print('phase', 'A' if ph == 0 else 'B')
    for i in range(8):
        col = a[:, i]; ok = col > 0
        if ok.any():
            d = (col[ok] - t0) / 1e3
            print('    %-15s n=%3d  min %7.2f  median %7.2f  max %7.2f us' % (names[i], ok.sum(), d.min(), np.median(d), d.max()))
