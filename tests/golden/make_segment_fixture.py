"""TBPTT segmentation table from the reference's OWN SegmentSequence transformer (datasets.py:41-138), executed
unmodified on a minimal stand-in for fuel.transformers.Transformer, with the arguments parrot_stream uses
(datasets.py:286-292: seq_size + 1, share_value=1, return_last=False, add_flag=True).  For a range of utterance
lengths the (start, stop, start_flag) of every emitted window goes to tests/golden/segments.json;
tests/test_cli_and_data.py holds parrot_b200.datasets.segment_sequence to it.

    python tests/golden/make_segment_fixture.py        # build container only
"""
import json
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/datasets.py'


class Transformer(object):
    """fuel.transformers.Transformer, as far as SegmentSequence uses it."""
    def __init__(self, data_stream, produces_examples=False, **kwargs):
        self.data_stream = data_stream
        self.produces_examples = produces_examples
        self.child_epoch_iterator = iter(data_stream.batches)


class Upstream(object):
    produces_examples = False
    sources = ('features', 'features_mask', 'labels')

    def __init__(self, batches):
        self.batches = batches


def load_class():
    src = open(REF).read().split('\n')
    a = next(i for i, l in enumerate(src) if l.startswith('class SegmentSequence'))
    b = next(i for i in range(a + 1, len(src)) if src[i] and not src[i][0].isspace())
    ns = {'Transformer': Transformer, 'numpy': numpy}
    exec(compile('\n' * a + '\n'.join(src[a:b]), REF, 'exec'), ns)
    return ns['SegmentSequence']


def main():
    Seg = load_class()
    table = {}
    for seq_size in (20, 50):
        for L in list(range(12, 200, 7)) + [seq_size + 1, seq_size + 11, seq_size + 12, 2 * seq_size + 11]:
            feats = numpy.arange(L, dtype=numpy.int64)[:, None, None].repeat(2, 1)      # (L, B=2, 1): value = frame index
            mask = numpy.ones((L, 2))
            t = Seg(Upstream([(feats, mask, 'labels')]), seq_size=seq_size + 1, share_value=1, return_last=False,
                    add_flag=True, which_sources=('features', 'features_mask'))
            wins = []
            while True:
                try:
                    f, m, lab, flag = t.get_data()
                except StopIteration:
                    break
                assert lab == 'labels' and f.shape[0] == m.shape[0]
                wins.append([int(f[0, 0, 0]), int(f[-1, 0, 0]) + 1, int(flag)])
            table['%d:%d' % (seq_size, L)] = wins
    json.dump(table, open(os.path.join(HERE, 'segments.json'), 'w'), sort_keys=True)
    print(len(table), 'lengths;', table['50:117'])


if __name__ == '__main__':
    main()
