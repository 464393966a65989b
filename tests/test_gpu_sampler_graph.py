"""GPU test: the CUDA-graph replay of the sampling loop (api.cu sample_scan) produces exactly what the plain
launch-by-launch loop produces, on the first (capturing) call and on later replays with new inputs and a new seed."""
import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('which_cost', ['MSE', 'GMM'])
def test_sampler_graph_replay_matches_plain_loop(which_cost):
    from parrot_b200 import _lib
    from parrot_b200.model import Parrot
    cfg = dict(util.TINY, rnn_h_dim=128, readouts_dim=128, weak_feedback=True, which_cost=which_cost,
               attention_alignment=0.4)
    B, T, U = 6, 24, 16
    res = {}
    for mode in ('graph', 'plain'):
        m = Parrot(**cfg)
        m.initialize(seed=5, gain=0.5)
        outs = []
        for call in range(3):      # call 0 captures the graph, calls 1-2 replay it with other inputs / seeds
            bt = util.make_batch(cfg, B, T, U, seed=30 + call)
            if mode == 'plain':
                h = m._handle(B, T, U, True)
                _lib.load().parrot_set_profiling(h.ptr, 1)      # profiling on: sample_scan runs the plain loop
            xs = m.sample_model(bt['labels'], bt['labels_mask'], None, None, B, T, seed=100 + call)
            torch.cuda.synchronize()
            outs.append([np.array(x) for x in xs])
        res[mode] = outs
        del m
        torch.cuda.empty_cache()
    for a, b in zip(res['graph'], res['plain']):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # different seeds / inputs really produce different samples
    assert not np.array_equal(res['graph'][0][0], res['graph'][1][0])
