import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from parrot_b200 import Parrot
cfg = dict(bench.BASE)
m = Parrot(**cfg); m.initialize(seed=0)
bt = bench.make_batch(cfg, 10, 8, 128, seed=1)
for T in (256, 2048):
    m.sample_model(bt['labels'], bt['labels_mask'], None, None, 10, T, as_numpy=False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m.sample_model(bt['labels'], bt['labels_mask'], None, None, 10, T, as_numpy=False)
    e1.record(); torch.cuda.synchronize()
    print('T=%d: %.1f us per step' % (T, e0.elapsed_time(e1) * 1e3 / 3 / T))
