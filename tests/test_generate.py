"""generate_wav front half (generate.py:58-115): de-normalisation, stream split, unvoiced rule -- pinned to the
reference's own lines, executed on a stand-in for Merlin's BinaryIOCollection, with the one thing a Python-3
interpreter cannot reproduce by execution (the iteration order of a Python-2 dict literal) computed from CPython 2.7's
string hash and dict probing."""
import collections
import os

import numpy as np

REF = '/root/reference/generate.py'


def py2_hash(s, bits=64):
    """CPython 2.7 string_hash (Objects/stringobject.c), hash randomisation off (the default)."""
    mask = (1 << bits) - 1
    x = (ord(s[0]) << 7) & mask
    for c in s:
        x = ((1000003 * x) ^ ord(c)) & mask
    x = (x ^ len(s)) & mask
    if x >= 1 << (bits - 1):
        x -= 1 << bits
    return -2 if x == -1 else x


def py2_dict_order(keys, bits=64):
    """Iteration order of a CPython 2.7 dict of <= 5 string keys inserted in ``keys`` order (table of 8 slots,
    Objects/dictobject.c lookdict_string: i = (i << 2) + i + perturb + 1, perturb >>= 5)."""
    table = [None] * 8
    for k in keys:
        h = py2_hash(k, bits)
        i, perturb = h & 7, h & ((1 << bits) - 1)
        while table[i & 7] is not None:
            i = (i << 2) + i + perturb + 1
            perturb >>= 5
        table[i & 7] = k
    return [k for k in table if k is not None]


def test_python2_dict_order_simulation_known_answers():
    assert py2_hash('a') == 12416037344                                      # CPython 2.7, 64-bit
    assert py2_dict_order(['a', 'b', 'c']) == ['a', 'c', 'b']               # {'a': 1, 'b': 2, 'c': 3}.keys()
    assert py2_dict_order(['one', 'two', 'three']) == ['three', 'two', 'one']
    assert py2_dict_order(['a', 'b', 'c'], 32) == ['a', 'c', 'b']


def test_stream_order_is_the_reference_interpreters():
    import generate
    src_order = ['bap', 'lf0', 'mgc', 'vuv']          # source order of the literal at generate.py:78
    for bits in (64, 32):
        assert py2_dict_order(src_order, bits) == [n for n, _ in generate.STREAMS] == ['mgc', 'vuv', 'lf0', 'bap']


class _IO(object):
    """Merlin BinaryIOCollection, as far as generate.py uses it: raw float32 files."""
    def array_to_binary_file(self, data, path):
        np.asarray(data, np.float32).tofile(path)

    def load_binary_file_frame(self, path, dim):
        a = np.fromfile(path, dtype=np.float32).reshape((-1, dim))
        return a, a.shape[0]


def _reference_split(data, gen_dir, base, norm_info_file):
    """generate.py:62-115 executed from the reference's source; only the dict literal of line 78 is replaced by the
    same mapping in CPython-2.7 iteration order."""
    if not os.path.exists(REF):
        return None
    src = open(REF).read().split('\n')
    a = next(i for i, l in enumerate(src) if l.strip().startswith('io_funcs = BinaryIOCollection()'))
    b = next(i for i, l in enumerate(src) if l.strip().startswith('pf_coef = 1.4'))
    body = [l[4:] for l in src[a:b]]                  # the lines live inside ``def generate_wav``
    lit = next(i for i, l in enumerate(body) if l.startswith('out_dimension_dict = {'))
    assert "'bap': 1, 'lf0': 1, 'mgc': 60, 'vuv': 1" in body[lit]
    body[lit] = 'out_dimension_dict = _py2_dict'
    order = py2_dict_order(['bap', 'lf0', 'mgc', 'vuv'])
    dims = {'bap': 1, 'lf0': 1, 'mgc': 60, 'vuv': 1}
    ns = dict(BinaryIOCollection=_IO, numpy=np, os=os, xrange=range, data=data, gen_dir=gen_dir, base=base,
              norm_info_file=norm_info_file, _py2_dict=collections.OrderedDict((k, dims[k]) for k in order))
    exec(compile('\n' * a + '\n'.join(body), REF, 'exec'), ns)
    return {ext: np.fromfile(os.path.join(gen_dir, base + '.' + ext), dtype=np.float32) for ext in ('cmp', 'mgc', 'lf0', 'bap')}


def test_generate_wav_front_half_matches_the_reference_lines(tmp_path):
    import generate
    rng = np.random.default_rng(0)
    frames = rng.standard_normal((37, 63)).astype(np.float32)
    norm = np.stack([rng.standard_normal(63), rng.uniform(0.5, 2.0, 63)]).astype(np.float32)
    norm[0, 60] = 0.5                                   # vuv column centred on the 0.5 threshold: both branches occur
    norm_file = str(tmp_path / 'norm.dat')
    norm.tofile(norm_file)
    ours_dir, ref_dir = str(tmp_path / 'ours'), str(tmp_path / 'ref')
    os.makedirs(ref_dir)
    files = generate.generate_wav(frames.copy(), ours_dir, 'utt', None, None, norm_file)
    ours = {ext: np.fromfile(files[ext], dtype=np.float32) for ext in ('cmp', 'mgc', 'lf0', 'bap')}
    # known answers, independent of the reference mount
    den = frames * norm[1] + norm[0]
    assert np.array_equal(ours['cmp'], den.ravel())
    assert np.array_equal(ours['mgc'], den[:, :60].ravel())
    lf0 = den[:, 61].copy()
    lf0[den[:, 60] < 0.5] = -1.0e10
    assert np.array_equal(ours['lf0'], lf0) and np.array_equal(ours['bap'], den[:, 62])
    assert (lf0 == np.float32(-1.0e10)).any() and (lf0 != np.float32(-1.0e10)).any()
    ref = _reference_split(frames.copy(), ref_dir, 'utt', norm_file)
    if ref is not None:                                 # build container only (/root/reference is not on the GPU box)
        for ext in ref:
            assert np.array_equal(ours[ext], ref[ext]), ext
