"""Per-group, per-phase timeline of the grouped persistent scans (debug tool, run under gpurun).
    python tools/group_timeline.py [T] [fwd|bwd]
PARROT_GROUPS_F / PARROT_GROUPS_B / PARROT_TC select the partition (defaults 64,40,44 forward, 64,36,48 backward, Tc 16)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from parrot_b200 import Parrot, _lib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
which = sys.argv[2] if len(sys.argv) > 2 else 'fwd'
env = os.environ.get('PARROT_GROUPS_F', '64,40,44') if which == 'fwd' else os.environ.get('PARROT_GROUPS_B', '64,36,48')
grp = [int(x) for x in env.split(',')]
Tc = int(os.environ.get('PARROT_TC', 16 if T >= 64 else 8))
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


def sequence(g):
    seq = []
    for k in range(T):
        if which == 'fwd':
            if g > 0 and k % Tc == 0:
                seq.append(('chunk', k))
            seq += [('G', k), ('C', k)]
            if g == 0:
                seq += [('proj', k), ('window', k)]
        else:
            if g == 0 or k == 0 or os.environ.get('PARROT_NO_FUSED_PRE'):
                seq.append(('pre', k))
            seq += [('bwd1', k), ('bwd2', k)]
            if g > 0 and ((k + 1) % Tc == 0 or k == T - 1):
                seq.append(('chunk', k))
    return seq


seqs = [sequence(g) for g in range(3)]
bars = max(len(s) for s in seqs)
ncta = sum(grp)
st = torch.zeros(148 * bars * 2, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(st.data_ptr()), bars + ((1 << 20) if which == 'bwd' else 0))
step()
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
s = st.cpu().numpy().reshape(148, bars, 2).astype(np.float64)
t0 = s[s > 0].min()
s = np.where(s > 0, (s - t0) / 1e3, np.nan)
cta0 = 0
for g in range(3):
    sg = s[cta0:cta0 + grp[g]]
    cta0 += grp[g]
    seq = seqs[g]
    first, last = np.nanmin(sg[:, 0, 0]), np.nanmax(sg[:, len(seq) - 1, 1])
    print('group %d (%d CTAs): first barrier pass %.1f us, last arrival %.1f us, %.2f us per step overall'
          % (g, grp[g], first, last, (last - first) / T))
    lo, hi = 3 * Tc, T - Tc
    tot = 0.0
    for nm in ['chunk', 'G', 'C', 'proj', 'window', 'pre', 'bwd1', 'bwd2']:
        idx = [i for i, (n, k) in enumerate(seq) if n == nm and lo <= k < hi and i + 1 < len(seq)]
        if not idx:
            continue
        passed = sg[:, idx, 0]; arr = sg[:, idx, 1]
        first_pass = np.nanmin(passed, axis=0)
        nxt = np.array([np.nanmin(sg[:, i + 1, 0]) for i in idx])
        span = nxt - first_pass
        work = arr - passed
        print('    %-7s n=%4d  span median %7.2f us  (work median %.2f, slowest-CTA median %.2f, last arrival -> next pass %.2f)'
              % (nm, len(idx), np.median(span), np.nanmedian(work), np.median(np.nanmax(work, axis=0)),
                 np.median(nxt - np.nanmax(arr, axis=0))))
        tot += np.median(span) * (1.0 / Tc if nm == 'chunk' else 1.0)
    print('    sum per step (chunk amortised): %.2f us' % tot)

# ---- intra-phase milestones of one steady-state step (both scan phases), per group
tick = 3 * Tc + 8
tl = torch.zeros(3 * 148 * 16, dtype=torch.int64, device='cuda')
lib.parrot_debug_set_stamps(h.ptr, C.c_void_p(tl.data_ptr()), -tick)
step()
torch.cuda.synchronize()
lib.parrot_debug_set_stamps(h.ptr, None, 0)
dbg = tl.cpu().numpy()[2 * 148 * 16:2 * 148 * 16 + 8].astype(np.float64)
if which == 'bwd' and (dbg > 0).any():
    print('attention-backward, batch row 0, step %d: milestones (us) entry / operands in smem / dphi done / '
          'reductions + datt done / dh1 + pre-pass done / exit:' % tick, [round((x - dbg[0]) / 1e3, 2) for x in dbg[:6]])
tl = tl.cpu().numpy()[:2 * 148 * 16].reshape(2, 148, 16).astype(np.float64)
names = {1: 'barrier_passed(epi)', 10: 'barrier_passed(tma)', 9: 'operands_requested', 11: 'first_stage_landed',
         2: 'tma_all_issued', 3: 'mma_all_issued', 4: 'acc_ready', 5: 'part_written', 6: 'all_arrived',
         7: 'finish_done', 8: 'phase_done'}
cta0 = 0
for g in range(3):
    for ph in range(2):
        a = tl[ph, cta0:cta0 + grp[g]]
        if not (a > 0).any():
            continue
        t0 = a[:, 1][a[:, 1] > 0].min()
        print('group %d phase %d (us after the first CTA passed the barrier)' % (g, ph))
        for i in (1, 10, 9, 11, 2, 3, 4, 5, 6, 7, 8):
            col = a[:, i]; ok = col > 0
            if ok.any():
                d = (col[ok] - t0) / 1e3
                print('    %-20s n=%3d  min %7.2f  median %7.2f  max %7.2f' % (names[i], ok.sum(), d.min(), np.median(d), d.max()))
    cta0 += grp[g]
