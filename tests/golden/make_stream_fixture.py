"""Batch pipeline table from the reference's OWN ``parrot_stream`` (datasets.py:206-298), executed unmodified --
together with its helpers ``_length`` / ``_transpose`` / ``_check_batch_size`` and its transformer classes
``SegmentSequence`` / ``SourceMapping`` / ``AddConstantSource`` -- on a stand-in for the Fuel names it imports
(``DataStream.default_stream``, ``SequentialExampleScheme``, ``ConstantScheme``, ``Batch``, ``Mapping``,
``SortMapping``, ``Unpack``, ``Filter``, ``Padding``, ``FilterSources``, ``Rename``, ``Transformer``,
``AgnosticSourcewiseTransformer``; semantics restated from Fuel 0.2, which is neither vendored nor pinned by the
reference).  This pins what the reference itself decides: the ORDER of the pipeline (sort inside windows of
``batch_size * sorting_mult`` examples, re-batch, drop the ragged batch, pad, keep sources, time-major, TBPTT
windows, noise-level source), the arguments of every stage and the layout of what comes out.

The dataset is a list of synthetic utterances whose feature VALUE identifies (utterance, frame), so the fixture can
store, for every emitted tuple, which utterance occupies which batch row and which frames each window covers.
tests/test_cli_and_data.py holds parrot_b200.datasets.parrot_stream to it.

    python tests/golden/make_stream_fixture.py        # build container only
"""
import json
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/datasets.py'
floatX = 'float32'


# ------------------------------------------------------------------ Fuel stand-in
class SequentialExampleScheme(object):
    def __init__(self, examples):
        self.examples = examples

    def get_request_iterator(self):
        return iter(range(self.examples))


class ShuffledExampleScheme(SequentialExampleScheme):
    pass   # not used by the fixture (the training order depends on Fuel's RNG)


class ConstantScheme(object):
    def __init__(self, batch_size):
        self.batch_size = batch_size


class DataStream(object):
    produces_examples = True
    axis_labels = None

    def __init__(self, dataset, iteration_scheme):
        self.dataset, self.scheme = dataset, iteration_scheme
        self.sources = dataset.provides_sources

    @classmethod
    def default_stream(cls, dataset, iteration_scheme=None):
        return cls(dataset, iteration_scheme)

    def get_epoch_iterator(self):
        return (self.dataset.get_data(request=i) for i in self.scheme.get_request_iterator())


class Transformer(object):
    """fuel.transformers.Transformer: wraps a stream, exposes sources / produces_examples / get_epoch_iterator."""
    def __init__(self, data_stream, produces_examples=None, axis_labels=None, **kwargs):
        self.data_stream = data_stream
        self.produces_examples = data_stream.produces_examples if produces_examples is None else produces_examples
        self.axis_labels = axis_labels
        self.child_epoch_iterator = data_stream.get_epoch_iterator()

    @property
    def sources(self):
        return getattr(self, '_sources', self.data_stream.sources)

    @sources.setter
    def sources(self, v):
        self._sources = v

    def get_data(self, request=None):
        data = next(self.child_epoch_iterator)
        return self.transform(data)

    def get_epoch_iterator(self):
        def it():
            while True:
                try:
                    yield self.get_data()
                except StopIteration:
                    return
        return it()


class Batch(Transformer):
    """examples -> batches of ``batch_size`` (the last one may be smaller, strictness=0)."""
    def __init__(self, data_stream, iteration_scheme, **kwargs):
        super(Batch, self).__init__(data_stream, produces_examples=False)
        self.n = iteration_scheme.batch_size

    def get_data(self, request=None):
        rows = []
        for _ in range(self.n):
            try:
                rows.append(next(self.child_epoch_iterator))
            except StopIteration:
                break
        if not rows:
            raise StopIteration
        out = []
        for col in zip(*rows):
            try:
                arr = numpy.asarray(col)
                if arr.dtype == object:
                    raise ValueError
            except ValueError:
                arr = numpy.empty(len(col), dtype=object)
                for i, c in enumerate(col):
                    arr[i] = c
            out.append(arr)
        return tuple(out)


class Mapping(Transformer):
    def __init__(self, data_stream, mapping, add_sources=None, **kwargs):
        super(Mapping, self).__init__(data_stream, **kwargs)
        self.mapping, self.add_sources = mapping, add_sources
        if add_sources:
            self.sources = tuple(data_stream.sources) + tuple(add_sources)

    def get_data(self, request=None):
        data = next(self.child_epoch_iterator)
        image = self.mapping(data)
        return image if not self.add_sources else tuple(data) + tuple(image)


class SortMapping(object):
    def __init__(self, key, reverse=False):
        self.key, self.reverse = key, reverse

    def __call__(self, batch):
        output = sorted(zip(*batch), key=self.key, reverse=self.reverse)
        cols = []
        for col in zip(*output):
            arr = numpy.empty(len(col), dtype=object)
            for i, c in enumerate(col):
                arr[i] = c
            cols.append(arr)
        return tuple(cols)


class Unpack(Transformer):
    def __init__(self, data_stream, **kwargs):
        super(Unpack, self).__init__(data_stream, produces_examples=True)
        self.pending = iter(())

    def get_data(self, request=None):
        while True:
            try:
                return next(self.pending)
            except StopIteration:
                self.pending = iter(list(zip(*next(self.child_epoch_iterator))))


class Filter(Transformer):
    def __init__(self, data_stream, predicate, **kwargs):
        super(Filter, self).__init__(data_stream)
        self.predicate = predicate

    def get_data(self, request=None):
        while True:
            data = next(self.child_epoch_iterator)
            if self.predicate(data):
                return data


class Padding(Transformer):
    """zero-pads every source to the longest example of the batch and adds ``<source>_mask`` behind it."""
    def __init__(self, data_stream, mask_sources=None, mask_dtype=None, **kwargs):
        super(Padding, self).__init__(data_stream, produces_examples=False)
        self.mask_sources = data_stream.sources if mask_sources is None else mask_sources
        src = []
        for s in data_stream.sources:
            src.append(s)
            if s in self.mask_sources:
                src.append(s + '_mask')
        self.sources = tuple(src)

    def transform(self, batch):
        out = []
        for s, col in zip(self.data_stream.sources, batch):
            if s not in self.mask_sources:
                out.append(col)
                continue
            shapes = [numpy.asarray(x).shape for x in col]
            lengths = [sh[0] for sh in shapes]
            rest = shapes[0][1:]
            dtype = numpy.asarray(col[0]).dtype
            padded = numpy.zeros((len(col), max(lengths)) + rest, dtype=dtype)
            mask = numpy.zeros((len(col), max(lengths)), dtype=floatX)
            for i, x in enumerate(col):
                padded[i, :lengths[i]] = x
                mask[i, :lengths[i]] = 1
            out += [padded, mask]
        return tuple(out)


class FilterSources(Transformer):
    def __init__(self, data_stream, sources, **kwargs):
        super(FilterSources, self).__init__(data_stream)
        self.sources = tuple(s for s in data_stream.sources if s in sources)

    def transform(self, data):
        return tuple(d for d, s in zip(data, self.data_stream.sources) if s in self.sources)


class Rename(Transformer):
    def __init__(self, data_stream, names, **kwargs):
        super(Rename, self).__init__(data_stream)
        self.sources = tuple(names.get(s, s) for s in data_stream.sources)

    def transform(self, data):
        return data


class AgnosticSourcewiseTransformer(Transformer):
    def __init__(self, data_stream, produces_examples, which_sources=None, **kwargs):
        super(AgnosticSourcewiseTransformer, self).__init__(data_stream, produces_examples, **kwargs)
        self.which_sources = data_stream.sources if which_sources is None else which_sources

    def transform(self, data):
        return tuple(self.transform_any_source(d, s) if s in self.which_sources else d
                     for d, s in zip(data, self.data_stream.sources))


# ------------------------------------------------------------------ synthetic dataset
class VoiceData(object):
    """Stand-in for the HDF5 dataset: utterance i has L_i frames of dimension 2 whose values are 1000 i + frame,
    U_i characters with values 100 i + position, speaker i % 5."""
    provides_sources = ('features', 'text', 'speaker_index')

    def __init__(self, voice=None, which_sets=None):
        rng = numpy.random.RandomState(7)
        self.num_examples = 37
        self.L = rng.randint(25, 160, self.num_examples)
        self.U = rng.randint(4, 12, self.num_examples)

    def get_data(self, request=None):
        i = request
        f = (1000 * i + numpy.arange(self.L[i]))[:, None].repeat(2, 1).astype('float32')
        t = (100 * i + numpy.arange(self.U[i])).astype('int32')
        return (f, t, numpy.array([i % 5], dtype='int32'))


def load_reference():
    src = open(REF).read().split('\n')

    def block(start):
        a = next(i for i, l in enumerate(src) if l.startswith(start))
        b = next((i for i in range(a + 1, len(src)) if src[i] and not src[i][0].isspace() and not src[i].startswith(')')),
                 len(src))
        # a multi-line ``def f(`` header ends at the first line starting with a non-space that is not '):'
        while b < len(src) and (src[b].startswith('        ') or src[b].startswith(')')):
            b += 1
        return '\n' * a + '\n'.join(src[a:b])
    ns = dict(numpy=numpy, Transformer=Transformer, AgnosticSourcewiseTransformer=AgnosticSourcewiseTransformer,
              Mapping=Mapping, Batch=Batch, Filter=Filter, FilterSources=FilterSources, Padding=Padding, Rename=Rename,
              SortMapping=SortMapping, Unpack=Unpack, DataStream=DataStream, ConstantScheme=ConstantScheme,
              ShuffledExampleScheme=ShuffledExampleScheme, SequentialExampleScheme=SequentialExampleScheme,
              VoiceData=VoiceData)
    for start in ('def _length', 'def _transpose', 'def _check_batch_size', 'def _check_ratio',
                  'class SegmentSequence', 'class SourceMapping', 'class AddConstantSource'):
        exec(compile(block(start), REF, 'exec'), ns)
    # parrot_stream: from its ``def`` line to the line before ``if __name__``
    a = next(i for i, l in enumerate(src) if l.startswith('def parrot_stream('))
    b = next(i for i, l in enumerate(src) if l.startswith('if __name__'))
    exec(compile('\n' * a + '\n'.join(src[a:b]), REF, 'exec'), ns)
    return ns


def main():
    ns = load_reference()
    table = {}
    for use_speaker, noise, bs, mult, seq in ((False, None, 4, 2, 20), (True, 0.5, 3, 4, 50), (False, None, 5, 1, 30)):
        stream = ns['parrot_stream']('vctk', use_speaker=use_speaker, which_sets=('valid',), batch_size=bs,
                                     seq_size=seq, sorting_mult=mult, noise_level=noise, labels_type='text',
                                     raw_data=False)
        rows = []
        for data in stream.get_epoch_iterator():
            d = dict(zip(stream.sources, data))
            f, m = d['features'], d['features_mask']
            assert f.shape[:2] == m.shape and f.shape[1] == bs
            rec = {
                'sources': list(stream.sources),
                'utt': [int(x) // 1000 for x in f[0, :, 0]],                 # utterance in every batch row
                'first': [int(x) % 1000 for x in f[0, :, 0]],                # first frame of the window
                'T': int(f.shape[0]),
                'valid': [int(x) for x in m.sum(0)],                          # valid frames per row in this window
                'labels_shape': list(d['labels'].shape),
                'labels_row0': [int(x) for x in d['labels'][0]],
                'labels_valid': [int(x) for x in d['labels_mask'].sum(1)],
                'start_flag': int(d['start_flag']),
            }
            if use_speaker:
                rec['speaker'] = [int(x) for x in numpy.asarray(d['speaker_index']).ravel()]
                rec['speaker_shape'] = list(numpy.asarray(d['speaker_index']).shape)
            if noise is not None:
                rec['noise'] = float(d['feedback_noise_level'])
            rows.append(rec)
        table['speaker=%d,noise=%s,bs=%d,mult=%d,seq=%d' % (use_speaker, noise, bs, mult, seq)] = rows
    json.dump(dict(lengths=[int(x) for x in VoiceData().L], chars=[int(x) for x in VoiceData().U], streams=table),
              open(os.path.join(HERE, 'stream.json'), 'w'), sort_keys=True)
    for k, v in table.items():
        print(k, len(v), 'tuples; first:', {kk: v[0][kk] for kk in ('sources', 'utt', 'first', 'T', 'start_flag')})


if __name__ == '__main__':
    main()
