"""generate_wav (reference generate.py:58-263), signature kept.

The reference de-normalises the 63-dim frames, writes Merlin-style binary feature files and then
shells out to SPTK + WORLD executables.  Neither Merlin's ``io_funcs`` nor the SPTK/WORLD binaries
are in the reference tree or in this image, so this implementation does the front half --
de-normalisation (generate.py:65-72), stream split and the unvoiced-frame rule (generate.py:78-115),
float32 binary ``.cmp/.mgc/.lf0/.bap`` files -- and then either runs the external synthesis when
both tool directories exist, or stops and SAYS SO (no silent success).

Column order: the reference derives the per-stream offsets by iterating a Python-2 ``dict`` literal
``{'bap': 1, 'lf0': 1, 'mgc': 60, 'vuv': 1}`` (generate.py:78-88), i.e. in CPython-2.7 hash order.  That order is
deterministic (no hash randomisation by default) and is reproduced in tests/test_generate.py from the interpreter's
string hash and open-addressing probe sequence (checked against known CPython-2.7 facts): 'mgc' and 'vuv' collide in
slot 0, 'mgc' is inserted first, and the iteration order is **mgc(60) | vuv(1) | lf0(1) | bap(1)** -- Merlin, whose
code this is, composes its ``.cmp`` files by iterating the same kind of dict, so the training features have the same
column order.  (An earlier version of this file assumed mgc | lf0 | vuv | bap from the name of the normalisation
file, ``norm_info_mgc_lf0_vuv_bap_63_MVN.dat``; the executed reference lines say otherwise.)
"""
import os

import numpy

STREAMS = (('mgc', 60), ('vuv', 1), ('lf0', 1), ('bap', 1))
FILE_EXT = {'mgc': '.mgc', 'bap': '.bap', 'lf0': '.lf0', 'cmp': '.cmp'}


def _write_binary(array, path):
    numpy.asarray(array, numpy.float32).tofile(path)      # Merlin array_to_binary_file: raw float32


def generate_wav(data, gen_dir, base, sptk_dir, world_dir, norm_info_file,
                 do_post_filtering=True, mgc_dim=60, fl=1024, sr=16000):
    if not os.path.exists(gen_dir):
        os.makedirs(gen_dir)
    file_name = os.path.join(gen_dir, base + '.cmp')
    cmp_info = numpy.fromfile(norm_info_file, dtype=numpy.float32).reshape((2, -1))   # generate.py:65-69
    cmp_mean, cmp_std = cmp_info[0], cmp_info[1]
    data = numpy.asarray(data, numpy.float32) * cmp_std + cmp_mean                    # generate.py:72
    _write_binary(data, file_name)

    start = {}
    off = 0
    for name, dim in STREAMS:
        start[name] = off
        off += dim if name != 'mgc' else mgc_dim
    features = data.reshape((-1, off))
    files = {'cmp': file_name}
    for name in ('mgc', 'lf0', 'bap'):
        dim = mgc_dim if name == 'mgc' else 1
        cur = features[:, start[name]:start[name] + dim].copy()
        if name == 'lf0':                                                             # generate.py:103-110
            vuv = features[:, start['vuv']]
            cur[vuv < 0.5, 0] = -1.0e+10
        path = os.path.join(gen_dir, base + FILE_EXT[name])
        _write_binary(cur, path)
        files[name] = path

    have_tools = bool(sptk_dir) and bool(world_dir) and os.path.isdir(sptk_dir) and os.path.isdir(world_dir)
    if not have_tools:
        print('generate_wav: wrote %s; SPTK/WORLD binaries not found (sptk_dir=%r, world_dir=%r) -- '
              'no waveform synthesised.' % (sorted(files.values()), sptk_dir, world_dir))
        return files
    raise NotImplementedError(
        'generate_wav: SPTK/WORLD directories exist but the shell pipelines of generate.py:157-262 are '
        'outside the hot path (SURVEY 8f N4); feature files were written: %s' % sorted(files.values()))
