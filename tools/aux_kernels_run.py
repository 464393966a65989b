"""Runs one training step + a few sampling steps of the GMM / speaker configuration at base width with the
one-launch-per-phase schedule, so that ncu can capture the kernels that the persistent scan hides:

    PARROT_NO_PERSISTENT=1 ncu --set full --clock-control none --import-source on \
        --kernel-name regex:'emit_|sample_emit|attention_|encoder_|adam_|rownorm_' -c 40 \
        -o gpurun_out/aux python tools/aux_kernels_run.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parrot_b200.model import Parrot            # noqa: E402
from parrot_b200 import algorithms              # noqa: E402


def main():
    B, T, U = 64, int(os.environ.get('AUX_T', 64)), 128
    cfg = dict(input_dim=128, output_dim=63, rnn_h_dim=1024, readouts_dim=1024, num_characters=43, attention_size=10,
               encoder_type='bidirectional', encoder_dim=128, which_cost='GMM', k_gmm=20, weak_feedback=True,
               use_speaker=True, num_speakers=21, speaker_dim=128, attention_alignment=0.15)
    m = Parrot(**cfg)
    m.initialize()
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((T + 1, B, 63)).astype(np.float32)
    fm = np.ones((T + 1, B), np.float32)
    labels = rng.integers(0, 43, (B, U)).astype(np.int32)
    lm = np.ones((B, U), np.float32)
    spk = rng.integers(0, 21, (B, 1)).astype(np.int32)
    algo = algorithms.GradientDescent(model=m, step_rule=algorithms.CompositeRule(
        [algorithms.StepClipping(10.0), algorithms.Adam(1e-4)]))
    for _ in range(2):
        cost = algo.process_batch(dict(features=feats, features_mask=fm, labels=labels, labels_mask=lm,
                                       speaker_index=spk, start_flag=1.0))
    torch.cuda.synchronize()
    print('cost', float(cost))
    out = m.sample_model(labels, lm, None, spk, B, 8, seed=1)
    torch.cuda.synchronize()
    print('sampled', out[0].shape)


if __name__ == '__main__':
    main()
