"""Batch / TBPTT data contract of the reference (datasets.py:41-138, 206-298), restated without Fuel.

``parrot_stream`` yields tuples with the reference's sources, in the reference's layout:
``features`` (T, B, D) time-major, ``features_mask`` (T, B), ``labels`` (B, U), ``labels_mask`` (B, U)
[, ``speaker_index`` (B, 1)], ``start_flag`` [, ``feedback_noise_level``] -- after the same pipeline:
shuffle -> batches of ``batch_size * sorting_mult`` sorted by length -> batches of ``batch_size``
(ragged last batch dropped, datasets.py:259-260) -> Padding (masks) -> transpose to time-major
(datasets.py:274-275) -> SegmentSequence(seq_size + 1, share_value=1, return_last=False,
add_flag=True) (datasets.py:286-292).

The HDF5 ``VoiceData`` source needs Fuel/h5py and the VCTK/Blizzard files, which are not in this
tree; ``SyntheticVoice`` generates utterances of the same shape statistics instead
(datasets.py:322-332 frames-per-character notes).  A real source only has to provide
``get_example(i) -> dict(features (L, D) float32, text (U,) int[, speaker_index int])``.
"""
import numpy as np


class SyntheticVoice(object):
    """Seeded synthetic utterances: features ~ N(0,1) (the reference features are mean-variance
    normalised, sample.py:176-178), text ~ U{0..num_characters-1}, 8-16 frames per character."""

    def __init__(self, num_examples=256, output_dim=63, num_characters=43, min_chars=8, max_chars=24,
                 frames_per_char=(8, 16), num_speakers=21, seed=0):
        self.num_examples = num_examples
        self.output_dim = output_dim
        self.num_characters = num_characters
        self.min_chars, self.max_chars = min_chars, max_chars
        self.frames_per_char = frames_per_char
        self.num_speakers = num_speakers
        self.seed = seed

    def get_example(self, i):
        rng = np.random.default_rng((self.seed, i))
        U = int(rng.integers(self.min_chars, self.max_chars + 1))
        L = int(U * rng.uniform(*self.frames_per_char))
        return dict(features=rng.standard_normal((L, self.output_dim)).astype(np.float32),
                    text=rng.integers(0, self.num_characters, U).astype(np.int32),
                    speaker_index=int(rng.integers(0, self.num_speakers)))


def _pad(seqs, dtype):
    """fuel.transformers.Padding: zero-pad to the longest, return (batch, mask)."""
    n = max(len(s) for s in seqs)
    shape = (len(seqs), n) + tuple(np.asarray(seqs[0]).shape[1:])
    out = np.zeros(shape, dtype)
    mask = np.zeros((len(seqs), n), np.float32)
    for i, s in enumerate(seqs):
        out[i, :len(s)] = s
        mask[i, :len(s)] = 1
    return out, mask


def segment_sequence(features, features_mask, seq_size, share_value=1, min_size=10, return_last=False):
    """datasets.py:41-138 SegmentSequence on time-major arrays.  Yields (features, mask, start_flag)."""
    if not return_last:
        min_size = min_size + seq_size
    step = 0
    flag = 1
    n = features.shape[0]
    while True:
        yield features[step:step + seq_size], features_mask[step:step + seq_size], flag
        flag = 0
        step += seq_size
        step -= share_value
        if step + min_size >= n:
            return


def parrot_stream(voice, use_speaker=False, which_sets=('train',), batch_size=32, seq_size=50,
                  num_examples=None, sorting_mult=4, noise_level=None, labels_type='text',
                  raw_data=False, dataset=None, seed=0, epochs=1):
    """Generator over training tuples; see the module docstring.  ``voice`` is kept for signature
    compatibility (datasets.py:206-209); pass ``dataset=`` (default: SyntheticVoice)."""
    assert labels_type in ('text', 'unaligned_phonemes'), \
        'only sequence-level labels (text / unaligned_phonemes) are on the attention path'
    assert not raw_data, 'raw audio belongs to the sampleRNN path (SURVEY 8f)'
    ds = dataset or SyntheticVoice()
    n = num_examples or ds.num_examples
    sources = ('features', 'features_mask', 'labels', 'labels_mask')
    if use_speaker:
        sources += ('speaker_index',)
    sources += ('start_flag',)
    if noise_level is not None:
        sources += ('feedback_noise_level',)
    rng = np.random.default_rng(seed)

    def gen():
        for _ in range(epochs):
            order = rng.permutation(n) if 'train' in which_sets else np.arange(n)
            sorting_size = batch_size * sorting_mult
            for s0 in range(0, n, sorting_size):
                chunk = [ds.get_example(int(i)) for i in order[s0:s0 + sorting_size]]
                chunk.sort(key=lambda e: len(e['features']))           # SortMapping(_length)
                for b0 in range(0, len(chunk), batch_size):
                    ex = chunk[b0:b0 + batch_size]
                    if len(ex) != batch_size:                           # _check_batch_size filter
                        continue
                    feats, fmask = _pad([e['features'] for e in ex], np.float32)
                    labels, lmask = _pad([e['text'] for e in ex], np.int32)
                    feats = feats.swapaxes(0, 1)                        # _transpose: time-major
                    fmask = fmask.swapaxes(0, 1)
                    spk = np.array([[e['speaker_index']] for e in ex], np.int32)
                    for f, m, flag in segment_sequence(feats, fmask, seq_size + 1, share_value=1,
                                                       return_last=False):
                        out = (np.ascontiguousarray(f), np.ascontiguousarray(m), labels, lmask)
                        if use_speaker:
                            out += (spk,)
                        out += (flag,)
                        if noise_level is not None:
                            out += (noise_level,)
                        yield out
    g = gen()
    return _Stream(g, sources)


class _Stream(object):
    def __init__(self, gen, sources):
        self._gen = gen
        self.sources = sources

    def get_epoch_iterator(self, as_dict=False):
        if as_dict:
            return (dict(zip(self.sources, t)) for t in self._gen)
        return self._gen

    def __iter__(self):
        return self._gen
