"""Per-barrier timeline of the chunk-lagged persistent scans (debug tool, run under gpurun).
    python tools/persist_timeline2.py [T] [fwd|bwd]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from parrot_b200 import Parrot, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
which = sys.argv[2] if len(sys.argv) > 2 else 'fwd'
cfg = dict(bench.BASE)
B, U = 64, 128
m = Parrot(**cfg); m.initialize(seed=0)
bt = bench.make_batch(cfg, B, T, U, seed=1)


def step():
    m.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
    if which == 'bwd':
        m.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
h = m._last
lib = _lib.load()
Tc = 16 if T >= 64 else 8
nticks = T + 2 * Tc
seq = []
for tick in range(nticks):
    if which == 'fwd':
        seq += [('A', tick), ('B', tick)]
        if tick < T:
            seq += [('proj', tick), ('window', tick)]
    else:
        seq += [('phase0', tick), ('bwd1', tick), ('bwd2', tick)]
    if (tick + 1) % Tc == 0:
        seq.append(('chunk', tick))
bars = len(seq)
st = torch.zeros(148 * bars * 2, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(st.data_ptr()), bars + ((1 << 20) if which == 'bwd' else 0))
step()
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
s = st.cpu().numpy().reshape(148, bars, 2).astype(np.float64)
t0 = s[s > 0].min()
s = np.where(s > 0, (s - t0) / 1e3, np.nan)
lo, hi = 2 * Tc + 2, T - 2      # steady state: all three layers active
names = ['A', 'B', 'proj', 'window'] if which == 'fwd' else ['phase0', 'bwd1', 'bwd2']
names.append('chunk')
tot = 0.0
for nm in names:
    idx = [i for i, (n, t) in enumerate(seq) if n == nm and lo <= t < hi and i + 1 < bars]
    if not idx:
        continue
    passed = s[:, idx, 0]; arr = s[:, idx, 1]
    first_pass = np.nanmin(passed, axis=0)
    last_arr = np.nanmax(arr, axis=0)
    nxt = np.array([np.nanmin(s[:, i + 1, 0]) for i in idx])
    span = nxt - first_pass              # first CTA into this phase -> first CTA into the next
    work = arr - passed
    print('%-8s n=%4d  span median %.2f us  (work median %.2f, slowest-CTA median %.2f, last arrival -> next first pass %.2f)'
          % (nm, len(idx), np.median(span), np.nanmedian(work), np.median(np.nanmax(work, axis=0)),
             np.median(nxt - last_arr)))
    per_tick = np.median(span) * (1.0 / Tc if nm == 'chunk' else 1.0)
    tot += per_tick
print('sum per tick (chunk amortised): %.2f us' % tot)

# ---- intra-phase milestones of one steady-state tick (both GEMM phases)
tick = 2 * Tc + 8
tl = torch.zeros(3 * 148 * 16, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(tl.data_ptr()), -tick)
step()
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
dbg = tl.cpu().numpy()[2 * 148 * 16:2 * 148 * 16 + 8].astype(np.float64)
if (dbg > 0).any():
    print('attention-bwd row 0 milestones (us):', [round((x - dbg[0]) / 1e3, 2) for x in dbg[:6]])
tl = tl.cpu().numpy()[:2 * 148 * 16].reshape(2, 148, 16).astype(np.float64)
names = {1: 'barrier_passed(epi)', 10: 'barrier_passed(tma)', 9: 'operands_requested', 11: 'first_stage_landed', 2: 'tma_all_issued', 3: 'mma_all_issued', 4: 'acc_ready',
         5: 'part_written', 6: 'all_arrived', 7: 'finish_done', 8: 'phase_done'}
for ph in range(2):
    a = tl[ph]
    if not (a > 0).any():
        continue
    t0 = a[a > 0].min()
    print('phase', ph)
    for i in (1, 10, 9, 11, 2, 3, 4, 5, 6, 7, 8):
        col = a[:, i]; ok = col > 0
        if ok.any():
            d = (col[ok] - t0) / 1e3
            print('    %-20s n=%3d  min %7.2f  median %7.2f  max %7.2f us' % (names[i], ok.sum(), d.min(), np.median(d), d.max()))
