"""bench.py contract on CPU: the reference arm (oracle port timed on host cores) prints ONE JSON line with the
keys the driver reads, rank != 0 of a multi-rank launch prints nothing and exits 0, and the GPU arm refuses to
run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = ['--hidden', '64', '--batch', '4', '--frames', '8', '--text', '8', '--ref_frames', '8',
        '--steps', '1', '--warmup', '0']


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + extra, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(['--impl', 'reference'] + TINY)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and abs(d['value'] - 4 * 8 / (d['ms_per_step'] * 1e-3)) < 1e-6 * d['value']
    for k in ('metric', 'n_gpus', 'steps', 'warmup', 'scaling', 'vs_baseline', 'dtype', 'data', 'config'):
        assert k in d
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--gpus', '2'] + TINY, env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(TINY)
    assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]
