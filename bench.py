#!/usr/bin/env python
"""Benchmark of the Parrot attention-RNN hot path (BASELINE.json metric:
"vocoder acoustic frames/sec (fwd+bwd)"), one JSON line on stdout.

    python bench.py                                   # N=1, configs[1] "Parrot base" on cuda:0
    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3        # weak scaling, 64 rows / GPU
    python bench.py --impl reference --steps 2 --warmup 1   # CPU arm: numpy restatement of the Theano path

A "step" is one full training step over one synthetic batch: compute_cost (encoder, T decoder
steps, readout, emitter, cost) + backward (BPTT, weight gradients) + gradient allreduce (N>1) +
StepClipping/Adam.  frames/s = B_global * T / step time.

  value        inputs already resident in HBM when the timed region starts
  e2e          same metric through the public API with HOST (pinned) inputs: the host->device copies
               of the batch and a device->host read of the cost are inside the timed region
  roofline     the dominant kernel (job_kernel_tc, the tcgen05 gate-GEMM engine) over the forward
               scan: algorithmic FLOPs per launch / average launch duration (CUDA events, separate
               profiled steps after the timed region) against the measured bf16 peak
  cpu_baseline the numpy oracle (a port of the reference's Theano CPU path) timed on the host cores
               on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = 'vocoder acoustic frames/sec (fwd+bwd)'
UNIT = 'frames/s'

# BASELINE.json configs[1] "Parrot base" as mapped by SURVEY.md 8d (config 2)
BASE = dict(input_dim=420, output_dim=63, rnn_h_dim=1024, readouts_dim=1024, weak_feedback=True,
            which_cost='MSE', num_characters=43, attention_type='graves', attention_size=10,
            attention_alignment=0.15, encoder_type='bidirectional', encoder_dim=128)
B_PER_GPU, T_FRAMES, U_TEXT = 64, 800, 128


def workload_config(args, world):
    cfg = dict(BASE)
    if args.which_cost:
        cfg['which_cost'] = args.which_cost
    B, T, U = args.batch or B_PER_GPU, args.frames or T_FRAMES, args.text or U_TEXT
    if args.hidden:
        cfg['rnn_h_dim'] = cfg['readouts_dim'] = args.hidden
    return cfg, B, T, U


def make_batch(cfg, B, T, U, seed):
    from parrot_b200.synthetic import make_batch as mk      # product module: no tests / oracle import on the GPU arm
    return mk(cfg, B, T, U, seed=seed)


def algorithmic_flops(cfg, B, T):
    """gate-GEMM FLOPs per decoder step, forward: 2*B*(18 H^2 + 9 C H) (+ feedback D->3H)."""
    H = cfg['rnn_h_dim']
    C = 2 * cfg['encoder_dim']
    per_step = 2.0 * B * (18 * H * H + 9 * C * H)
    return per_step, per_step * T


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d['bf16_tflops_sustained'],
                    source='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU DURING the timed regions.  NVML in-process
    (nvidia_ml_py) when it is importable: spawning nvidia-smi five times a second takes driver locks that delay
    kernel launches, which the per-step synchronising e2e loop cannot hide.  Falls back to the nvidia-smi query."""

    NAMES = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
    BITS = [0x8, 0x40, 0x20, 0x4]     # nvmlClocksEventReason{HwSlowdown,HwThermalSlowdown,SwThermalSlowdown,SwPowerCap}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.source = 'nvidia-smi'

    def _nvml_open(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None

    def _nvml_sample(self, nv, h):
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        power = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
        get = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or \
            getattr(nv, 'nvmlDeviceGetCurrentClocksThrottleReasons')
        mask = int(get(h))
        return [str(sm), str(sm_max), str(power)] + ['Active' if mask & b else 'Not Active' for b in self.BITS]

    def _smi_sample(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
            'clocks_event_reasons.sw_power_cap'
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                              '--format=csv,noheader,nounits'], capture_output=True, text=True,
                             timeout=5).stdout.strip()
        return [x.strip() for x in out.split(',')] if out else None

    def run(self):
        nv = self._nvml_open()
        if nv:
            self.source = 'nvml'
        while not self.stop_flag:
            row = None
            if nv:
                try:
                    row = self._nvml_sample(*nv)
                except Exception:
                    nv, self.source = None, 'nvidia-smi'
            if row is None:
                try:
                    row = self._smi_sample()
                except Exception:
                    row = None
            if row and len(row) >= 7:
                self.samples.append(row)
            time.sleep(0.1 if nv else 0.5)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for i, n in enumerate(self.NAMES):
            if any(s[3 + i].lower().startswith('active') for s in self.samples):
                reasons.append(n)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(sm), 'power_w_max': max(float(s[2]) for s in self.samples), 'source': self.source}


# --------------------------------------------------------------------------- CPU arms
def set_blas_threads(n):
    """The CPU arms run the numpy oracle on BLAS threads.  torchrun exports OMP_NUM_THREADS=1, which made the same
    arm 6x slower under the launcher than stand-alone: the thread count is set explicitly (all host cores unless
    --blas_threads says otherwise) and the count actually in effect is reported."""
    want = n if n and n > 0 else (os.cpu_count() or 1)
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=want, user_api='blas')
        info = [d for d in threadpoolctl.threadpool_info() if d.get('user_api') == 'blas']
        return int(info[0]['num_threads']) if info else want
    except Exception:
        return int(os.environ.get('OMP_NUM_THREADS', want))


def oracle_step_time(cfg, B, T, U, steps, warmup, with_adam=True):
    """Oracle (numpy, BLAS-threaded) fwd + bwd (+ Adam) on a (B, T) sample; returns seconds per step."""
    from oracle.parrot_oracle import OracleAdamClip
    from tests import util
    orc = util.make_oracle(cfg, gain=None, bias_std=0)
    opt = OracleAdamClip(orc.shapes, learning_rate=1e-4, threshold=9.0)
    bt = util.make_batch(cfg, B, T, U, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.compute_cost(bt['features'], bt['features_mask'], bt['labels'], bt['labels_mask'], None, 1.0, B)
        g = orc.backward()
        if with_adam:
            opt.step(orc.params, g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return float(np.median(times))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  Theano/Blocks (Python 2) cannot
    run in this image, so this is the oracle port (numpy float32, BLAS threads = host cores)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg, B, T, U = workload_config(args, 1)
    Ts = min(T, args.ref_frames)
    cores = set_blas_threads(args.blas_threads)
    sec = oracle_step_time(cfg, B, Ts, U, max(1, args.steps), max(0, min(args.warmup, 1)))
    fps = B * Ts / sec
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'Parrot base (BASELINE configs[1]): B=%d H=%d E=%d U=%d, bounded sample T=%d of %d '
                               'frames (per-step cost is T-independent)' % (B, cfg['rnn_h_dim'], cfg['encoder_dim'], U, Ts, T),
                   'which_cost': cfg['which_cost']},
        'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': 'numpy restatement of the Theano CPU path, fwd+bwd+Adam, B=%d T=%d, %d BLAS threads '
                                   'of %d host cores' % (B, Ts, cores, os.cpu_count() or 0)},
        'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
    ap.add_argument('--which_cost', type=str, default=None)
    ap.add_argument('--batch', type=int, default=None, help='rows per GPU (default 64)')
    ap.add_argument('--frames', type=int, default=None, help='T (default 800)')
    ap.add_argument('--text', type=int, default=None, help='U (default 128)')
    ap.add_argument('--hidden', type=int, default=None)
    ap.add_argument('--ref_frames', type=int, default=40, help='bounded T of the CPU arms')
    ap.add_argument('--profile_steps', type=int, default=1)
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--gain', type=float, default=None,
                    help="'trained-like' parameter set of SURVEY 8d: W ~ N(0, (gain/sqrt(fan_in))^2) instead of the "
                         'init scale 0.01 (same kernels, same work; larger activations)')
    ap.add_argument('--scaling', type=str, default='weak', choices=['weak', 'strong'],
                    help='weak: 64 rows per GPU (default) ; strong: the 64-row batch of configs[1] split over the GPUs')
    ap.add_argument('--blas_threads', type=int, default=0, help='BLAS threads of the CPU arms (0: all host cores)')
    ap.add_argument('--no_sample', action='store_true', help='skip the sampling section (B=10, 2048 steps)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from parrot_b200 import Parrot, _lib, parallel
    from parrot_b200.algorithms import Adam, CompositeRule, GradientDescent, StepClipping
    import ctypes as C

    # NCCL_DEBUG is left as the launcher set it (the driver reads the communicator's INFO lines).  NCCL logs to
    # stdout by default: route them to stderr unless the launcher chose a file, so that stdout carries one JSON line
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    rank, world, local = parallel.init_from_env()
    assert world == args.gpus or world == 1 and args.gpus == 1, 'launch with torchrun for --gpus > 1'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    cfg, B, T, U = workload_config(args, world)
    if args.scaling == 'strong':
        assert B % world == 0, 'strong scaling splits the %d-row batch over the GPUs' % B
        B //= world
    W = max(3, args.warmup)
    K = max(1, args.steps)

    model = Parrot(device=dev, **cfg)
    model.initialize(seed=0, gain=args.gain)       # identical replicas (train.py:30-31 init; --gain: trained-like)
    algo = GradientDescent(model=model, parameters=None,
                           step_rule=CompositeRule([StepClipping(9.0), Adam(1e-4)]))
    bt = make_batch(cfg, B, T, U, seed=100 + rank)
    host = {k: torch.from_numpy(np.ascontiguousarray(bt[k])).pin_memory()
            for k in ('features', 'features_mask', 'labels', 'labels_mask')}
    devb = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    lib = _lib.load()

    def step(src):
        cost, _, _, _ = model.compute_cost(src['features'], src['features_mask'], src['labels'],
                                           src['labels_mask'], None, 1.0, B)
        algo.step()
        return cost

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, n, read_cost):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        l0 = lib.parrot_launch_count()
        e0.record()
        for _ in range(n):
            c = step(src)
            if read_cost:
                c.item()                            # device -> host read of the step's result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), lib.parrot_launch_count() - l0

    for _ in range(W):
        step(devb)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(devb, K, False)
    step(host).item()        # untimed: first use of the host-input path (staging buffers of the caching allocator)
    ms_e2e, _ = timed(host, K, True)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    frames_global = B * world * T
    value = frames_global * K / (ms_dev * 1e-3)
    e2e = frames_global * K / (ms_e2e * 1e-3)

    # ---- profiled steps: per-launch durations of the dominant kernels (rank 0)
    roof = None
    extra = {}
    # Profiled steps run on EVERY rank (they contain the gradient allreduce; a collective executed by rank 0
    # alone would deadlock); only rank 0 reads the profile back.
    if True:
        h = model._last
        P = max(1, args.profile_steps)

        def prof(key):
            tot = C.c_double(); n = C.c_int64()
            _lib.check(lib.parrot_get_profile(h.ptr, key.encode(), C.byref(tot), C.byref(n)))
            return tot.value, n.value
        # level 1: section events around the (persistent) scan kernels -- the numbers the roofline uses
        lib.parrot_set_profiling(h.ptr, 1)
        for _ in range(P):
            step(devb)
        torch.cuda.synchronize()
        sec = {k: prof(k) for k in ('sec_pack_prep_encoder', 'sec_scan_fwd', 'sec_readout_emit_fwd',
                                    'sec_readout_emit_bwd', 'sec_scan_bwd', 'sec_grads_tail')}
        # level 2: one launch per phase with events around every launch (diagnostic breakdown only)
        lib.parrot_set_profiling(h.ptr, 2)
        for _ in range(P):
            step(devb)
        torch.cuda.synchronize()
    if rank == 0:

        pk = peaks()
        per_step_flops, _ = algorithmic_flops(cfg, B, T)
        # the forward scan is ONE persistent launch (scan_fwd_grouped): all T decoder steps, gate GEMMs, hoisted
        # chunk products, attention and the group barriers.  Its CUDA-event duration is the denominator
        # (conservative: attention phases and barriers are inside it).
        scan_ms = sec['sec_scan_fwd'][0] / P
        bwd_ms = sec['sec_scan_bwd'][0] / P
        achieved_tf = per_step_flops * T / (scan_ms * 1e-3) / 1e12
        # dram bytes of the same kernels from the committed ncu --set full capture of the shipped binary
        # (tools/summarize_ncu.py writes profiles/r02_scan_traffic.json); base workload only
        traffic, traffic_bwd, traffic_src = None, None, None
        tj = os.path.join(ROOT, 'profiles', 'r02_scan_traffic.json')
        if (B, T, U, cfg['rnn_h_dim'], cfg['which_cost']) == (64, 800, 128, 1024, 'MSE') and os.path.exists(tj):
            td = json.load(open(tj))
            traffic, traffic_bwd, traffic_src = td.get('scan_fwd_grouped'), td.get('scan_bwd_grouped'), td.get('source')
        roof = {'bound': 'tensor',
                'kernel': 'scan_fwd_grouped (T decoder steps in one launch: three layer groups, tcgen05 gate GEMMs '
                          'with TMEM-resident weights, hoisted chunk products, attention)',
                'achieved': achieved_tf, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved_tf / pk['tf_sustained'],
                'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': pk['source'] + ' bf16 sustained (kernel timed inside a long step)',
                'algorithmic_flops_per_launch': per_step_flops * T,
                'avg_launch_us': scan_ms * 1e3, 'launches_per_step': 1,
                'us_per_decoder_step': scan_ms * 1e3 / T}
        bwd_tf = per_step_flops * T / (bwd_ms * 1e-3) / 1e12     # the dgrads contract the same weights once
        extra['roofline_bwd'] = {
            'bound': 'tensor', 'kernel': 'scan_bwd_grouped (reverse sweep: attention backward, GRU dgrads, hoisted '
                                         'chunk dgrads)',
            'achieved': bwd_tf, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s', 'frac': bwd_tf / pk['tf_sustained'],
            'traffic': traffic_bwd, 'traffic_source': traffic_src,
            'algorithmic_flops_per_launch': per_step_flops * T, 'avg_launch_us': bwd_ms * 1e3,
            'us_per_decoder_step': bwd_ms * 1e3 / T}
        # why the tensor fraction is small: at batch 64 every decoder step re-streams the bf16x3 weight planes
        # (hi+lo, 4 bytes per parameter on the recurrent path) -- report that stream against the HBM peak too
        Hh, Cc2 = cfg['rnn_h_dim'], 2 * cfg['encoder_dim']
        wbytes = 4.0 * (18 * Hh * Hh + 9 * Cc2 * Hh + 3 * cfg['output_dim'] * Hh)
        extra['roofline_weight_stream'] = {
            'bound': 'hbm', 'kernel': roof['kernel'], 'achieved': wbytes * T / (scan_ms * 1e-3) / 1e9,
            'peak': pk['hbm'], 'unit': 'GB/s', 'frac': wbytes * T / (scan_ms * 1e-3) / 1e9 / pk['hbm'],
            'algorithmic_bytes_per_decoder_step': wbytes,
            'note': 'hi+lo operand planes of every in-scan weight counted once per decoder step; since round 2 the '
                    'recurrent part (41 MB) is partly resident in tensor memory and the rest stays in L2'}
        # attention-step latency (second half of the BASELINE metric): one stand-alone parrot_attention_step call
        # (one launch: projection + window) at B x U x C of the workload, CUDA events over 200 back-to-back calls
        H, Cc, A = cfg['rnn_h_dim'], 2 * cfg['encoder_dim'], cfg['attention_size']
        att_bytes = 4.0 * (B * U * Cc + B * U + B * Cc + 3 * B * A)
        g = torch.Generator(device='cpu').manual_seed(0)
        rnd = lambda *shape: torch.randn(*shape, generator=g).to(dev)
        a_h1, a_wT, a_b, a_ctx = rnd(B, H).tanh(), rnd(3 * A, H) * 0.03, rnd(3 * A) * 0.1, rnd(B, U, Cc)
        a_k = rnd(B, A).abs() * 10
        a_ko, a_wo, a_po = torch.zeros(B, A, device=dev), torch.zeros(B, Cc, device=dev), torch.zeros(B, U, device=dev)
        a_ab, a_e = torch.zeros(B, 2 * A, device=dev), torch.zeros(B, 3 * A, device=dev)
        ccfg = model._make_cfg(B, 1, U, 0)
        pp = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def att_call():
            _lib.check(lib.parrot_attention_step(C.byref(ccfg), pp(a_h1), pp(a_wT), pp(a_b), pp(a_ctx), pp(a_k),
                                                 pp(a_ko), pp(a_wo), pp(a_po), pp(a_ab), pp(a_e), 1, stream))
        for _ in range(20):
            att_call()
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(200):
            att_call()
        ev1.record()
        torch.cuda.synchronize()
        att_us = ev0.elapsed_time(ev1) * 1e3 / 200
        extra['attn_step_latency_us'] = att_us
        extra['roofline_attention'] = {
            'bound': 'hbm', 'kernel': 'parrot_attention_step: attention_step_kernel, projection + window in one launch (K7)',
            'achieved': att_bytes / (att_us * 1e-6) / 1e9, 'peak': pk['hbm'], 'unit': 'GB/s',
            'frac': att_bytes / (att_us * 1e-6) / 1e9 / pk['hbm'], 'traffic': None,
            'algorithmic_bytes_per_launch': att_bytes, 'avg_launch_us': att_us,
            'note': 'ctx (8.4 MB) is L2-resident when the step is called back to back'}
        # free-running generation (model.py:827-1083) at the reference's own sampling shape (utils.py:270-273:
        # num_samples 10, num_steps 2048): per-step latency bound -- every step consumes the frame the previous one
        # emitted.  One call = input copies + ONE CUDA-graph launch of the whole loop.
        if not args.no_sample:
            SB, ST = 10, 2048
            lab, lm = bt['labels'][:SB], bt['labels_mask'][:SB]
            model.sample_model(lab, lm, None, None, SB, ST, as_numpy=False)       # builds the handle + the graph
            torch.cuda.synchronize()
            ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(3):
                model.sample_model(lab, lm, None, None, SB, ST, as_numpy=False)
            ev1.record()
            torch.cuda.synchronize()
            s_ms = ev0.elapsed_time(ev1) / 3
            extra['sample'] = {'batch': SB, 'steps': ST, 'frames_per_s': SB * ST / (s_ms * 1e-3),
                               'us_per_step': s_ms * 1e3 / ST, 'ms_per_call': s_ms,
                               'launch': 'CUDA graph of the per-phase launches, one graph launch per call',
                               'workload': 'sample_model, %d rows x %d steps, U=%d, %s emitter'
                                           % (SB, ST, U, cfg['which_cost'])}
            if not args.no_cpu_baseline:
                from tests import util as _u
                orc = _u.make_oracle(cfg, gain=None, bias_std=0)
                set_blas_threads(args.blas_threads)
                t0 = time.perf_counter()
                orc.sample_model(lab, lm, None, None, SB, 16)
                extra['sample']['cpu_port_us_per_step'] = (time.perf_counter() - t0) * 1e6 / 16
        sections = {k: {'ms': round(v[0] / P, 3), 'launches': v[1] // P} for k, v in sec.items()}
        extra['sections_ms_persistent'] = sections
        diag = {}
        for key in ('sec_scan_fwd', 'sec_scan_bwd', 'fwdA', 'fwdB', 'attn_fwd', 'bwd1', 'bwd2', 'attn_bwd',
                    'gru_bwd_pre', 'readout', 'output', 'dread', 'dh_readout', 'wgrad', 'tail_dctx',
                    'tail_encoder_bwd', 'tail_weight_grads', 'tail_bias_grads', 'tail_speaker'):
            ms, n = prof(key)
            diag[key] = {'ms': round(ms / P, 3), 'launches': n // P}
        extra['breakdown_ms_per_phase_launch_mode'] = diag
        lib.parrot_set_profiling(h.ptr, 0)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        Ts = min(T, args.ref_frames)
        cores = set_blas_threads(args.blas_threads)
        sec = oracle_step_time(cfg, B, Ts, U, 2, 1, with_adam=False)
        cpu = {'value': B * Ts / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': 'numpy float32 restatement of the Theano CPU path (oracle), fwd+bwd, B=%d T=%d U=%d '
                         '(per-step cost is T-independent), %d BLAS threads of %d host cores'
                         % (B, Ts, U, cores, os.cpu_count() or 0)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'bf16x3 (fp32 operands split hi+lo in bf16, 3 tcgen05 MMAs, fp32 accumulate)',
            'data': 'synthetic',
            'config': {'workload': 'Parrot base (BASELINE configs[1]): 1xBiGRU enc E=%d + 3xGRU dec H=%d, '
                                   'batch=%d/GPU, T_text=%d, T_frames=%d, %s cost, weak feedback, fwd+bwd+Adam'
                                   % (cfg['encoder_dim'], cfg['rnn_h_dim'], B, U, T, cfg['which_cost']),
                       'global_batch': B * world, 'parallelism': 'dp%d' % world,
                       'parameters': 'init scale N(0, 0.01^2) (train.py:30-31)' if args.gain is None
                       else 'trained-like N(0, (%g/sqrt(fan_in))^2)' % args.gain,
                       'l2': 'working set (13 GB workspace) far exceeds the 126 MB L2; no explicit flush',
                       'encoder_time_axis': 0},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / K},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': roof,
            'cpu_baseline': cpu,
        }
        line.update(extra)
        sys.stdout.flush()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
