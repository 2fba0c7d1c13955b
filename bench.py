#!/usr/bin/env python
"""Benchmark of the THA4 poser hot path on B200 (contract: see the task's bench.py section).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload teacher_b1|student_b64|teacher_b16]

A "step" is one poser forward over one batch of synthetic input.  Default workload (N=1) is BASELINE.json
configs[1]: the full five-network poser (mode_07), batch 1, the lambda_00 character image, one random pose per
step, eyebrow cache hot (the image does not change between frames, as in the reference's GUI).  Metric: 512x512
RGBA frames/sec.  One JSON line is printed by rank 0.

  value        device-timed throughput with image and poses already resident in HBM;
  e2e          the same through the public API with pinned HOST buffers: H2D of image + pose and D2H of the posed
               frame inside the timed region, every step;
  roofline     the dominant kernel (implicit-GEMM convolution: tensor bound), measured in a separate profiled pass
               with CUDA events inside the library; roofline_tail is the fused grid_sample + decoder kernel of the
               upscaler (HBM bound), the kernel BASELINE.json's metric names;
  cpu_baseline the CPU oracle (a PyTorch-CPU port of the reference path, oracle/) on this box's host cores.
`--impl reference` times that CPU port alone, as the reference arm.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    'teacher_b1': dict(mode='mode_07', batch=1, desc='full poser mode_07 forward, batch=1, lambda_00 image, random poses, eyebrow cache hot'),
    'teacher_b1_nocache': dict(mode='mode_07', batch=1, nocache=True, desc='full poser mode_07 forward, batch=1, the image changes every frame (eyebrow-decomposer cache always misses: 645.9 GFLOP/frame)'),
    'pose_sweep_512': dict(mode='mode_07', total=512, desc='BASELINE configs[3]: 512-pose sweep of the lambda_00 image, contiguous shards of 512/N frames per GPU (micro-batches of 32), no collective'),
    'teacher_b16': dict(mode='mode_07', batch=16, desc='full poser mode_07 forward, batch=16 pose sweep on the lambda_00 image'),
    'student_b64': dict(mode='mode_14', batch=64, desc='distilled student mode_14 forward, batch=64, lambda_00 weights, fp16 tensor-core products'),
    'distill_b1': dict(mode='distill', batch=1, desc='body-morpher distill step: teacher mode_07 fwd + student fwd/bwd + gradient all-reduce + Adam, per-GPU batch 1 (reference-faithful: total batch <= 8)'),
}
TEACHER_GFLOP_PER_FRAME = 625.9   # cache-hot


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p['bf16_tflops'], tflops_sustained=p.get('bf16_tflops_sustained'), source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        return dict(sm_mhz=statistics.median(self.samples) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def load_inputs(workload):
    from tha4_b200 import image_util, synthetic
    png = os.path.join(ROOT, 'tests', 'golden', 'data', 'lambda_00.png')
    image = image_util.load_poser_image(png) if os.path.exists(png) else synthetic.synthetic_image(0, 1)[0]
    return image


def load_state_dicts(mode):
    from tha4_b200 import synthetic
    if mode == 'mode_07':
        return synthetic.teacher_state_dicts(0), 'random-init (seeded) teacher weights of the reference architecture'
    data = os.path.join(ROOT, 'tests', 'golden', 'data')
    paths = {k: os.path.join(data, 'lambda_00_%s.pt' % k) for k in ('face_morpher', 'body_morpher')}
    if all(os.path.exists(p) for p in paths.values()):
        return {k: torch.load(p, map_location='cpu') for k, p in paths.items()}, 'shipped lambda_00 student weights'
    return synthetic.student_state_dicts(0), 'random-init (seeded) student weights'


def cpu_threads():
    """Host threads for the CPU arm.  PyTorch-CPU convs on these 128-core boxes get *slower* past a few dozen threads
    (measured: 128 threads are 40x slower than 8 on the teacher), so the arm uses min(cores, 32) and says so."""
    return max(1, min(os.cpu_count() or 1, int(os.environ.get('THA4_CPU_THREADS', '32'))))


def cpu_port_fps(mode, sds, image, poses, batch, frames_budget, threads, seconds_budget=20.0):
    """Times the CPU oracle (PyTorch-CPU port of the reference path).  Returns (fps, frames, seconds)."""
    from oracle import tha4_oracle
    torch.set_num_threads(threads)
    fn = getattr(tha4_oracle, mode + '_outputs')
    with torch.no_grad():
        dec = None
        if mode == 'mode_07':   # eyebrow cache hot, as in the GPU arm
            dec = tha4_oracle.eyebrow_decomposer(sds['eyebrow_decomposer'], image.unsqueeze(0)[:, :, 64:192, 192:320])
        b = min(batch, 2)
        img_b = image.unsqueeze(0).expand(b, -1, -1, -1).contiguous()
        kw = dict(cached_decomposer_output=[t.expand(b, -1, -1, -1) for t in dec]) if dec is not None else {}
        fn(sds, img_b, poses[:b], **kw)            # warm-up
        t0 = time.perf_counter()
        done = 0
        while done < frames_budget and (done == 0 or time.perf_counter() - t0 < seconds_budget):
            fn(sds, img_b, poses[done % 8:done % 8 + b] if poses.shape[0] >= 8 + b else poses[:b], **kw)
            done += b
        dt = time.perf_counter() - t0
    return done / dt, done, dt


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path (its PyTorch-CPU port in oracle/)."""
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    from tha4_b200 import synthetic
    sds, _ = load_state_dicts(wl['mode'])
    image = load_inputs(wl)
    poses = synthetic.random_poses(64, seed=1234)
    threads = cpu_threads()
    per_step_frames = 1 if wl['mode'] == 'mode_07' else 2
    from oracle import tha4_oracle
    torch.set_num_threads(threads)
    fn = getattr(tha4_oracle, wl['mode'] + '_outputs')
    b = per_step_frames
    img_b = image.unsqueeze(0).expand(b, -1, -1, -1).contiguous()
    kw = {}
    times = []
    with torch.no_grad():
        if wl['mode'] == 'mode_07':     # eyebrow cache hot, as in the GPU arm (mode_07.py:56-68)
            dec = tha4_oracle.eyebrow_decomposer(sds['eyebrow_decomposer'], img_b[:, :, 64:192, 192:320])
            kw = dict(cached_decomposer_output=dec)
        t_start = time.perf_counter()
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            fn(sds, img_b, poses[(i * b) % 32:(i * b) % 32 + b], **kw)
            if i >= args.warmup:
                times.append((time.perf_counter() - t0) / b)
            if times and time.perf_counter() - t_start > 150.0:     # bounded sample: keep the arm within minutes
                break
    spf = sum(times) / len(times)
    line = {
        'impl': 'reference', 'metric': '512x512 RGBA frames/sec', 'value': 1.0 / spf, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * spf * wl.get('batch', wl.get('total', 1)), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['desc'], 'batch': wl.get('batch', wl.get('total', 1))},
        'cpu_baseline': {'value': 1.0 / spf, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d timed steps (of %d requested) of %d frame(s) each of the same workload (PyTorch-CPU port of the reference path, '
                                   '%d of %d host threads)' % (len(times), args.steps, per_step_frames, threads, os.cpu_count() or 1)},
        'e2e': {'value': 1.0 / spf, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def run_torch_cuda(args, rank):
    """Context number (not part of the contract): the oracle's PyTorch ops executed on the GPU = what the reference's own
    PyTorch-CUDA eager path does on this box (same ops, cuDNN/cuBLAS kernels, TF32 convs allowed as by torch's default)."""
    if rank != 0:
        return
    from oracle import tha4_oracle
    from tha4_b200 import synthetic
    wl = WORKLOADS[args.workload]
    mode = 'mode_14' if wl['mode'] == 'mode_14' else 'mode_07'
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    sds, _ = load_state_dicts(mode)
    sds = {k: {kk: vv.to(dev) for kk, vv in v.items()} for k, v in sds.items()}
    B = wl.get('batch', 16)
    image = load_inputs(wl).to(dev).unsqueeze(0).expand(B, -1, -1, -1).contiguous()
    poses = synthetic.random_poses((args.warmup + args.steps) * B, seed=1234).to(dev)
    fn = getattr(tha4_oracle, mode + '_outputs')
    _orig = tha4_oracle.base_grid
    tha4_oracle.base_grid = lambda n, h, w, dtype=torch.float32: _orig(n, h, w, dtype).to(dev)
    tha4_oracle._timestep_embedding_zero_orig = tha4_oracle._timestep_embedding_zero
    tha4_oracle._timestep_embedding_zero = lambda n, c: tha4_oracle._timestep_embedding_zero_orig(n, c).to(dev)
    kw = {}
    with torch.no_grad():
        if mode == 'mode_07':
            kw = dict(cached_decomposer_output=tha4_oracle.eyebrow_decomposer(sds['eyebrow_decomposer'], image[:, :, 64:192, 192:320]))
        for i in range(args.warmup):
            fn(sds, image, poses[i * B:(i + 1) * B], **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            fn(sds, image, poses[(args.warmup + i) * B:(args.warmup + i + 1) * B], **kw)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    emit({'impl': 'torch_cuda_eager', 'metric': '512x512 RGBA frames/sec', 'value': args.steps * B / (ms / 1000.0),
                      'unit': 'frames/s', 'ms_per_step': ms / args.steps, 'steps': args.steps, 'warmup': args.warmup,
                      'config': {'workload': wl['desc']},
          'note': 'PyTorch eager on the same GPU running the oracle port (the ops the reference dispatches); context only'})


_REAL_STDOUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries (NCCL prints its version banner there) share fd 1, so
    everything else is sent to stderr and the JSON line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='tha4_b200', choices=['tha4_b200', 'reference', 'torch_cuda'])
    ap.add_argument('--workload', default='teacher_b1', choices=sorted(WORKLOADS))
    ap.add_argument('--strict', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--option', action='append', default=[], help='library option name=value (developer A/B runs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.impl == 'torch_cuda':
        run_torch_cuda(args, rank)
        return

    wl = WORKLOADS[args.workload]
    B = wl['batch'] if 'batch' in wl else max(1, wl['total'] // world)     # fixed total: strong scaling
    strong = 'total' in wl
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    from tha4_b200 import synthetic
    from tha4_b200.poser.modes import mode_07, mode_14
    distill = wl['mode'] == 'distill'
    sds, weights_desc = load_state_dicts('mode_07' if distill else wl['mode'])
    image = load_inputs(wl)
    nposes = (args.warmup + args.steps) * B
    poses = synthetic.random_poses(nposes, seed=1234 + rank)
    poser = (mode_14 if wl['mode'] == 'mode_14' else mode_07).create_poser(device, state_dicts=sds)
    poser.get_modules()
    ctx = poser.get_context()
    ctx.set_option('strict', args.strict)
    for kv in args.option:
        k, v = kv.split('=')
        ctx.set_option(k, int(v))
    distiller = None
    if distill:
        from tha4_b200.distill import BodyMorpherDistiller
        student_sds, sdesc = load_state_dicts('mode_14')
        distiller = BodyMorpherDistiller(poser, mode_14.load_body_morpher(None, student_sds['body_morpher']))
        weights_desc += '; student: ' + sdesc
    DISTILL_W, DISTILL_LR = [0.0, 1.0, 1.0, 0.0], 1e-4        # phase 1 of the body schedule: warp + grid-change terms

    img_dev = image.unsqueeze(0).expand(B, -1, -1, -1).contiguous().to(device)
    poses_dev = poses.to(device)
    img_alt = None
    if wl.get('nocache'):          # a second image that differs in one pixel value: the cache comparison fails every frame
        img_alt = img_dev.clone()
        img_alt[:, 0, 0, 0] += 1.0 / 512.0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        if distiller is not None:
            return distiller.train_step(img_dev, poses_dev[i * B:(i + 1) * B], DISTILL_W, DISTILL_LR, want_losses=False)
        return poser.get_posing_outputs(img_alt if (img_alt is not None and (i & 1)) else img_dev, poses_dev[i * B:(i + 1) * B])

    # ---------------- device-resident timing ----------------
    with torch.no_grad():
        for i in range(args.warmup):
            step_resident(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = ctx.counter('kernel_launches')
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_resident(args.warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ctx.counter('kernel_launches') - l0
        clocks = sampler.stop()

        # ---------------- end to end through the public API with host buffers ----------------
        img_host = image.unsqueeze(0).expand(B, -1, -1, -1).contiguous().pin_memory()
        poses_host = poses.pin_memory()
        out_host = torch.empty((B, 4, 512, 512), dtype=torch.float32).pin_memory()
        img_in = torch.empty_like(img_dev)
        pose_in = torch.empty((B, 45), device=device)

        img_in2 = torch.empty_like(img_dev) if img_alt is not None else None

        def step_e2e(i):
            img_cur = img_in
            if img_alt is not None and (i & 1):      # a different tensor object with different content: the cache must miss
                img_cur = img_in2
                img_cur.copy_(img_host, non_blocking=True)
                img_cur[:, 0, 0, 0] += 1.0 / 512.0
            else:
                img_cur.copy_(img_host, non_blocking=True)
            pose_in.copy_(poses_host[i * B:(i + 1) * B], non_blocking=True)
            if distiller is not None:        # result of a training step = its loss terms, read back on the host
                distiller.train_step(img_cur, pose_in, DISTILL_W, DISTILL_LR, want_losses=True)
                return
            out = poser.pose(img_cur, pose_in)
            out_host.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()      # the caller consumes the frame on the host

        for i in range(args.warmup):
            step_e2e(i)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(args.steps):
            step_e2e(args.warmup + i)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)

        # ---------------- profiled pass for the roofline objects (rank 0) ----------------
        prof = {}
        if rank == 0 or distiller is not None:    # a distillation step contains the gradient all-reduce: every rank takes part
            if rank == 0:
                ctx.set_option('profile', 2)
            for i in range(args.steps):
                step_resident(args.warmup + i)
            torch.cuda.synchronize()
            if rank == 0:
                for cat in ('conv', 'norm', 'tail', 'attn', 'siren'):
                    prof[cat] = {w: ctx.counter('prof_%s_%s' % (w, cat)) for w in ('us', 'launches', 'flops', 'bytes')}
                ctx.set_option('profile', 0)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        cl = torch.tensor([float(launches)], device=device)
        dist.all_reduce(cl)
        launches = int(cl[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    frames = args.steps * B * world
    value = frames / (ms / 1000.0)
    e2e_value = frames / (ms_e2e / 1000.0)
    line = {
        'metric': 'distillation examples/sec' if distill else '512x512 RGBA frames/sec', 'value': value,
        'unit': 'examples/s' if distill else 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak', 'vs_baseline': None,
        'dtype': 'f16/tf32 operands (10-bit mantissa), f32 accumulate, f32 storage outside conv operands' if wl['mode'] in ('mode_07', 'distill') and not args.strict else
                 ('f32 (3xTF32)' if wl['mode'] == 'mode_07' else 'f16 products, f32 accumulate'),
        'data': 'synthetic poses; ' + weights_desc + '; lambda_00.png character image',
        'config': {'workload': wl['desc'], 'batch_per_gpu': B, 'parallelism': ('data parallel, one NCCL all-reduce of the 1.33 MB flat gradient per step (dp%d)' if distill else 'frames sharded, no collective (dp%d)') % world,
                   'l2': 'packed weights (657 MB teacher) and activations exceed the 126 MB L2; no explicit flush'},
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': B * (4 * 512 * 512 * 4 + 45 * 4),
                'd2h_bytes_per_step': 32 if distill else B * 4 * 512 * 512 * 4, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'clocks': clocks,
    }
    if wl['mode'] == 'mode_07':
        line['teacher_tflops_effective'] = (645.9 if wl.get('nocache') else TEACHER_GFLOP_PER_FRAME) * value / world / 1000.0
    # roofline objects from the profiled pass
    if prof.get('conv', {}).get('us', 0) > 0:
        c = prof['conv']
        ach = c['flops'] / (c['us'] * 1e-6) / 1e12
        line['roofline'] = {'kernel': 'conv_tc_kernel (implicit-GEMM conv: TMA + tcgen05.mma kind::f16/tf32, TMEM accumulator; all conv launches of the step)', 'bound': 'tensor', 'achieved': ach,
                            'peak': peaks['tflops'], 'unit': 'TFLOP/s', 'frac': ach / peaks['tflops'], 'traffic': None,
                            'peak_source': peaks['source'] + ' dense bf16 burst (= the f16 operand rate; kind::tf32 layers peak at half of it)',
                            'avg_launch_us': c['us'] / max(1, c['launches']), 'launches_per_step': c['launches'] / args.steps,
                            'share_of_profiled_kernel_time': c['us'] / max(1.0, sum(v['us'] for v in prof.values()))}
    if prof.get('tail', {}).get('us', 0) > 0:
        t = prof['tail']
        ach = t['bytes'] / (t['us'] * 1e-6) / 1e9
        line['roofline_tail'] = {'kernel': 'tail_kernel (fused head conv + grid_sample + blend, all 4 teacher sites)', 'bound': 'hbm',
                                 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': ach / peaks['hbm_gbs'],
                                 'traffic': None, 'peak_source': peaks['source'], 'avg_launch_us': t['us'] / max(1, t['launches'])}
    if prof.get('siren', {}).get('us', 0) > 0:
        line['siren_us_per_step'] = prof['siren']['us'] / args.steps
    line['kernel_time_us_per_step'] = {k: v['us'] / args.steps for k, v in prof.items() if v['us'] > 0}
    if 'roofline' not in line:
        s = prof.get('siren', {})
        line['roofline'] = {'kernel': 'siren fused MLP kernels', 'bound': 'tensor', 'achieved': None, 'peak': peaks['tflops'],
                            'unit': 'TFLOP/s', 'frac': None, 'traffic': None, 'us_per_step': s.get('us', 0) / max(1, args.steps)}
        if s.get('us', 0) > 0:
            ach = 37.89e9 * B * args.steps / (s['us'] * 1e-6) / 1e12
            line['roofline'].update(achieved=ach, frac=ach / peaks['tflops'])

    if not args.no_cpu_baseline and world == 1 and not distill:
        threads = cpu_threads()
        budget = 6 if wl['mode'] == 'mode_07' else 12
        fps, nfr, dt = cpu_port_fps(wl['mode'], sds, image, poses, B, budget, threads)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                                'sample': '%d frames of the same workload in %.1f s (PyTorch-CPU port of the reference path in oracle/, '
                                          '%d of %d host threads; /root/reference itself is pure Python and does not exist on the GPU box)' % (nfr, dt, threads, os.cpu_count() or 1)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
