"""Host-side contract of the reference entry points, on CPU: flag surface (utils.py:173-327), the sampling stop
heuristic (sample.py:147-163) and the batch / TBPTT layout of the data stream (datasets.py:206-298)."""
import json
import os

import numpy as np

from oracle import parrot_oracle as O
from parrot_b200 import datasets, utils

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def test_flag_surface_matches_the_reference_table():
    """tests/golden/cli_flags.json is extracted from the reference's parsers by tests/golden/make_cli_fixture.py:
    same flag names, same literal defaults; additions are the documented three."""
    table = json.load(open(os.path.join(GOLD, 'cli_flags.json')))
    extra = {'train_parse': {'synthetic', 'steps', 'seed'}, 'sample_parse': {'seed'}}
    for fn, ref in table.items():
        ours = vars(getattr(utils, fn)([]))
        assert set(ref) - set(ours) == set(), fn
        assert set(ours) - set(ref) == extra[fn], fn
        for k, v in ref.items():
            if v['default'] != '<expr>':
                assert ours[k] == v['default'], (fn, k)
            if v['type'] == 'bool':                      # reference: type=bool (any non-empty string is True)
                assert isinstance(ours[k], bool)


def test_boolean_flags_parse_false_strings_as_false():
    a = utils.train_parse(['--layer_norm', 'False', '--weak_feedback', 'true', '--use_speaker', '0'])
    assert a.layer_norm is False and a.weak_feedback is True and a.use_speaker is False
    assert utils.train_parse(['--encoder_type', 'none']).encoder_type is None
    a = utils.train_parse(['--dataset', 'blizzard', '--save_dir', '/tmp/x'])
    assert a.save_dir == os.path.join('/tmp/x', 'blizzard')                  # utils.py:250-251


def test_stop_heuristic_matches_oracle_and_known_case():
    rng = np.random.default_rng(0)
    for _ in range(20):
        T, U = 60, 12
        phi = rng.random((T, U + 2)).astype(np.float32)
        L = int(rng.integers(3, U))
        assert utils.stop_heuristic(phi, L, T) == O.stop_heuristic(phi, L, T)
    phi = np.zeros((30, 8), np.float32)
    phi[:, 1] = 1.0
    phi[7:, 5] = 2.0                                   # position 5 (= labels_length) wins from frame 7 on
    assert utils.stop_heuristic(phi, 5, 100) == 47     # first frame + 40 frames of slack
    assert utils.stop_heuristic(phi, 5, 20) == 20      # clamped to num_steps
    assert utils.stop_heuristic(np.zeros((4, 8), np.float32), 5, 9) == 9   # never ends -> num_steps


def test_stream_layout_sorting_and_segments():
    ds = datasets.SyntheticVoice(num_examples=40, seed=3)
    st = datasets.parrot_stream('synthetic', use_speaker=True, batch_size=4, seq_size=20, sorting_mult=2,
                                noise_level=0.3, dataset=ds, seed=1)
    assert st.sources == ('features', 'features_mask', 'labels', 'labels_mask', 'speaker_index', 'start_flag',
                          'feedback_noise_level')
    batches = list(st.get_epoch_iterator(as_dict=True))
    assert batches
    n_first = 0
    prev_tail = None
    for b in batches:
        f, m, lab, lm = b['features'], b['features_mask'], b['labels'], b['labels_mask']
        assert f.ndim == 3 and f.shape[1] == 4 and f.shape[2] == 63 and f.dtype == np.float32   # time-major
        assert m.shape == f.shape[:2] and lab.shape == lm.shape and lab.shape[0] == 4
        assert f.shape[0] <= 21                                     # seq_size + 1 frames per segment
        assert b['speaker_index'].shape == (4, 1) and b['feedback_noise_level'] == 0.3
        assert ((f != 0).any(-1) <= (m > 0)).all()                  # padding is zero where the mask is zero
        if b['start_flag'] == 1:
            n_first += 1
            # lengths inside a batch come from one sorted chunk: masks are nested (sorted by length)
            lens = m.sum(0) if f.shape[0] < 21 else None
        else:
            assert np.array_equal(prev_tail, f[0])                  # consecutive segments share one frame
        prev_tail = f[-1]
    assert n_first == 40 // 4                                        # every full batch starts exactly once
    # ragged last batch is dropped (datasets.py:259-260): 42 examples -> still 10 batches
    ds2 = datasets.SyntheticVoice(num_examples=42, seed=3)
    st2 = datasets.parrot_stream('synthetic', batch_size=4, seq_size=20, sorting_mult=2, dataset=ds2, seed=1)
    assert sum(1 for t in st2 if t[-1] == 1) == 10


def test_segmentation_matches_the_reference_transformer():
    """tests/golden/segments.json: windows emitted by the reference's own SegmentSequence (datasets.py:41-138,
    executed by tests/golden/make_segment_fixture.py with parrot_stream's arguments) for 61 utterance lengths."""
    table = json.load(open(os.path.join(GOLD, 'segments.json')))
    assert len(table) > 50
    for key, wins in table.items():
        seq, L = map(int, key.split(':'))
        f = np.arange(L)[:, None, None].repeat(2, 1)
        m = np.ones((L, 2))
        ours = [[int(a[0, 0, 0]), int(a[-1, 0, 0]) + 1, int(flag)]
                for a, _, flag in datasets.segment_sequence(f, m, seq + 1, share_value=1, return_last=False)]
        assert ours == wins, key


def test_parrot_stream_matches_the_reference_pipeline():
    """tests/golden/stream.json: the reference's own ``parrot_stream`` (datasets.py:206-298) with its helpers and
    transformer classes, executed on a Fuel stand-in over synthetic utterances whose feature values identify
    (utterance, frame) (make_stream_fixture.py).  parrot_b200.datasets.parrot_stream must emit the same tuples: same
    sources, same utterance in every batch row (sort inside windows of batch_size * sorting_mult, ragged batch
    dropped), same TBPTT windows and start flags, same padding of features and labels."""
    import json
    import os
    import numpy as np
    from parrot_b200 import datasets
    fx = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'stream.json')))
    L, U = fx['lengths'], fx['chars']

    class Voice(object):
        num_examples = len(L)

        def get_example(self, i):
            f = (1000 * i + np.arange(L[i]))[:, None].repeat(2, 1).astype(np.float32)
            return dict(features=f, text=(100 * i + np.arange(U[i])).astype(np.int32), speaker_index=i % 5)

    assert len(fx['streams']) == 3
    for key, rows in fx['streams'].items():
        kv = dict(p.split('=') for p in key.split(','))
        noise = None if kv['noise'] == 'None' else float(kv['noise'])
        use_speaker = kv['speaker'] == '1'
        st = datasets.parrot_stream('vctk', use_speaker=use_speaker, which_sets=('valid',), batch_size=int(kv['bs']),
                                    seq_size=int(kv['seq']), sorting_mult=int(kv['mult']), noise_level=noise,
                                    labels_type='text', dataset=Voice())
        got = list(st.get_epoch_iterator(as_dict=True))
        assert len(got) == len(rows), (key, len(got), len(rows))
        for d, r in zip(got, rows):
            assert list(st.sources) == r['sources']
            f, m = d['features'], d['features_mask']
            assert f.shape[0] == r['T'] and f.shape[:2] == m.shape
            assert [int(x) // 1000 for x in f[0, :, 0]] == r['utt']
            assert [int(x) % 1000 for x in f[0, :, 0]] == r['first']
            assert [int(x) for x in m.sum(0)] == r['valid']
            # every valid frame continues its utterance; padding is zero
            for b in range(f.shape[1]):
                n = r['valid'][b]
                assert (f[:n, b, 0] == 1000 * r['utt'][b] + r['first'][b] + np.arange(n)).all()
                assert (f[n:, b] == 0).all()
            assert list(d['labels'].shape) == r['labels_shape']
            assert [int(x) for x in d['labels'][0]] == r['labels_row0']
            assert [int(x) for x in d['labels_mask'].sum(1)] == r['labels_valid']
            assert int(d['start_flag']) == r['start_flag']
            if use_speaker:
                assert list(np.asarray(d['speaker_index']).shape) == r['speaker_shape']
                assert [int(x) for x in np.asarray(d['speaker_index']).ravel()] == r['speaker']
            if noise is not None:
                assert float(d['feedback_noise_level']) == r['noise']


def test_learning_rate_schedule_follows_the_reference_extension(tmp_path):
    """extensions.py:83-150: patience counter, NaN handling, reload-best + zero buffers + lr cut, stop after num_cuts."""
    import train

    class Adam(object):
        learning_rate = 1e-3

    class Algo(object):
        zeroed = 0

        def zero_buffers(self):
            self.zeroed += 1

    class Model(object):
        loaded = 0

        def set_parameter_values(self, v):
            self.loaded += 1

    path = str(tmp_path / 'best.npz')
    np.savez(path, w=np.zeros(2))
    adam, algo, model = Adam(), Algo(), Model()
    s = train.LearningRateSchedule(adam, algo, model, path, patience=3, num_cuts=2, cut_size=.5)
    assert not s.do(None) and not s.do(5.0) and not s.do(4.0)           # improving
    assert not s.do(4.5) and not s.do(4.5) and s.counter == 2
    assert not s.do(4.5)                                                 # third check without improvement: cut 1
    assert s.count_cuts == 1 and adam.learning_rate == 5e-4 and algo.zeroed == 1 and model.loaded == 1 and s.counter == 0
    assert s.do(float('nan'))                                            # NaN forces a cut at once: cut 2 -> finish
    assert s.count_cuts == 2 and adam.learning_rate == 2.5e-4 and algo.zeroed == 2
