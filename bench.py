"""bench.py -- frames/s of Vid2VidModelG.inference at 2048x1024 (BASELINE config 4:
`--label_nc 35 --loadSize 2048 --n_scales_spatial 3 --use_instance --fg --use_single_G`,
scripts/street/test_2048.sh) on synthetic 35-label sequences with random-init weights.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg2|tiny]

A step = one steady-state generated frame (all three scales).  One process per GPU; N > 1 runs N
independent clips (inference has no exchange step: "replicas only", weak scaling) and reports the
whole-job frames/s = N * K / max-over-ranks time.  Prints ONE JSON line on rank 0.

  value        device-resident inputs, CUDA-event timed (what the kernels can do)
  e2e          the public API with HOST tensors: pinned label maps -> H2D, ..., generated frame -> D2H
  roofline     the tcgen05 conv kernel: algorithmic conv FLOPs of a frame / summed device time of its launches
               (per-launch CUDA events, v2v_plan_profile) against MEASURED_PEAKS.json bf16 sustained
  cpu_baseline oracle port (PyTorch CPU restatement of the reference, oracle/generator_oracle.py) timed on the
               host cores on one steady-state frame of the same workload
`--impl reference` times that CPU port as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, n_scales, ngf, loadSize)
    'cfg4': dict(W=2048, H=1024, n_scales=3, ngf=128, desc='label2city 2048x1024 inference, n_scales_spatial=3 --fg --use_single_G'),
    'cfg2': dict(W=512, H=256, n_scales=1, ngf=128, desc='label2city 512x256 inference (use_single_G), n_scales_spatial=1 --fg'),
    'tiny': dict(W=256, H=128, n_scales=2, ngf=32, desc='plumbing: 256x128, 2 scales, ngf 32'),
}


def make_opt_for(wl):
    from vid2vid_b200.utils import make_opt
    w = WORKLOADS[wl]
    return make_opt(label_nc=35, use_instance=True, fg=True, fg_labels=[26], n_scales_spatial=w['n_scales'], ngf=w['ngf'],
                    use_single_G=True, loadSize=w['W'], dataroot='datasets/Cityscapes/', gpu_ids=[0])


def frame_macs(wl):
    from vid2vid_b200 import networks as NW
    opt = make_opt_for(wl)
    opt.gpu_ids = []
    w = WORKLOADS[wl]
    total = 0.0
    for s in range(w['n_scales']):
        sc = 2 ** (w['n_scales'] - 1 - s)
        total += NW.build_netG(opt, s).conv_macs(1, w['H'] // sc, w['W'] // sc)
    return total


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md): NVML every 50 ms, nvidia-smi fallback."""

    REASONS = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'sw_power_cap': 0x4}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        self.recording = False      # NVML init and the first (slow) queries happen during warm-up, outside the timed region
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    # the same queries whether or not we are recording, so their first-call costs are paid in warm-up
                    clk = self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM)
                    try:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    if self.recording:
                        self.sm.append(clk)
                        for n, bit in self.REASONS.items():
                            if mask & bit:
                                self.reasons.add(n)
                    time.sleep(0.05)
                elif not self.recording:
                    time.sleep(0.02)
                else:
                    q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
                        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
                    o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                       capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                    self.sm.append(int(float(o[0])))
                    self.max_mhz = int(float(o[1]))
                    for n, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], o[2:]):
                        if v.strip().lower().startswith('active'):
                            self.reasons.add(n)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.sm and self.nvml is not None:        # very short timed region: one sample right after it
            try:
                self.sm.append(self.nvml.nvmlDeviceGetClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            except Exception:
                pass
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(sm)}


def reduce_times(times_ms, world, device='cpu'):
    """Device-timed durations -> the slowest rank's (replicas finish when the last one does).  One all-reduce(MAX) on the
    job's process group (NCCL on the GPU box, gloo in the CPU tests); the data path itself has no collective."""
    import torch
    t = torch.tensor(list(times_ms), device=device, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1381.2), d.get('hbm_gbs', 6570.9), 'measured (MEASURED_PEAKS.json, bf16 sustained)'
    return 1400.0, 6650.0, 'fallback (B200_PROFILING.md)'


# ----------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vid2vid_b200 import _lib as L
    from vid2vid_b200.model_g import Vid2VidModelG
    from vid2vid_b200.utils import synth_label_sequence

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    wl = WORKLOADS[args.workload]
    opt = make_opt_for(args.workload)
    opt.gpu_ids = [local_rank]
    torch.manual_seed(1234 + rank)
    model = Vid2VidModelG().initialize(opt)
    tG = opt.n_frames_G
    K, Wm = args.steps, args.warmup
    n_frames = 2 * (K + Wm) + tG + 2
    seq = synth_label_sequence(n_frames, wl['H'], wl['W'], label_nc=35, block=64, seed=rank)   # (1, T, 1, H, W)
    seq_pinned = seq.pin_memory()
    seq_dev = seq.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: first-frame generator + W frames (also instantiates the CUDA graphs)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t = 0
    for _ in range(max(Wm, 3)):
        A = seq_dev[:, t:t + tG]
        model.inference(A, None, A)
        t += 1
    # ---- timed: device-resident inputs
    # (a generation-2 garbage collection over the module tree costs tens of ms -- several frames -- when it lands inside
    # a 0.4 s timed region; collect now, keep the collector off until both timed regions are done)
    import gc
    gc.collect()
    gc.disable()
    barrier()
    sampler.recording = True
    l0 = L.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        A = seq_dev[:, t:t + tG]
        model.inference(A, None, A)
        t += 1
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches = L.LAUNCHES[0] - l0
    # ---- timed: end to end through the public API with host tensors
    out_host = torch.empty((1, 3, wl['H'], wl['W']), dtype=torch.float32).pin_memory()
    for _ in range(2):
        A = seq_pinned[:, t:t + tG].to(dev, non_blocking=True)
        fb, _ = model.inference(A, None, A)
        out_host.copy_(fb, non_blocking=True)
        t += 1
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(K):
        A_h = seq_pinned[:, t:t + tG]
        A = A_h.to(dev, non_blocking=True)            # H2D: label ids of the tG-frame window
        I = A_h.to(dev, non_blocking=True)            # H2D: instance ids (same synthetic map)
        fb, _ = model.inference(A, None, I)
        out_host.copy_(fb, non_blocking=True)         # D2H: the generated frame (test.py's tensor2im(.cpu()))
        t += 1
    e3.record()
    barrier()
    sampler.stop_flag = True
    gc.enable()
    ms_e2e = e2.elapsed_time(e3)
    h2d = 2 * tG * wl['H'] * wl['W'] * 4
    d2h = 3 * wl['H'] * wl['W'] * 4

    ms_dev, ms_e2e = reduce_times([ms_dev, ms_e2e], world, dev)
    if rank != 0:
        return

    # ---- roofline of the conv kernel (per-launch CUDA events over one more frame)
    conv_ms, conv_macs, other_ms, top, conv_launches = 0.0, 0.0, 0.0, {}, 0
    kind_ms = {}
    for s in range(wl['n_scales']):
        net = getattr(model, 'netG%d' % s)
        for ent in net._plans().values():
            for kind, ms, macs in ent['plan'].profile():
                kind_ms[kind] = kind_ms.get(kind, 0.0) + ms
                if kind == 1:
                    conv_ms += ms
                    conv_macs += macs
                    conv_launches += 1
                else:
                    other_ms += ms
    peak_tf, peak_gbs, peak_src = peaks()
    achieved_tf = 2.0 * conv_macs / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    fmacs = frame_macs(args.workload)
    frame_ms = ms_dev / K
    # DRAM bytes per conv launch: not measurable here (needs ncu); taken from the committed ncu capture of this command
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'conv_dram_traffic.json')))[args.workload]
        traffic, traffic_src = tj['dram_bytes_per_launch_avg'], tj['source']
    except Exception:
        pass
    out = {
        'metric': 'frames/sec at 2048x1024 inference' if args.workload == 'cfg4' else 'frames/sec inference (%s)' % args.workload,
        'value': world * K / (ms_dev * 1e-3), 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': Wm,
        'ms_per_step': frame_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic 35-label blocky sequences, random-init weights (N(0,0.02))',
        'config': {'workload': wl['desc'], 'parallelism': 'replicas x%d (independent clips, no collective)' % world,
                   'frames_per_step': 1, 'conv_flops_per_frame': 2 * fmacs,
                   'l2_policy': 'per-frame working set (weights 0.8 GB + activations) is far larger than the 126 MB L2'},
        'e2e': {'value': world * K / (ms_e2e * 1e-3), 'unit': 'frames/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': launches,
        'clocks': sampler.summary(),
        'roofline': {'bound': 'tensor', 'kernel': 'conv_umma_kernel (all launches of one frame)', 'achieved': achieved_tf,
                     'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': achieved_tf / peak_tf, 'peak_source': peak_src, 'traffic': traffic,
                     'traffic_unit': 'bytes per launch (dram read + write, ncu, mean over the launches of a frame)',
                     'traffic_source': traffic_src, 'launches_per_frame': conv_launches,
                     'algorithmic_flops_per_launch_avg': 2.0 * conv_macs / max(conv_launches, 1),
                     'avg_launch_us': 1e3 * conv_ms / max(conv_launches, 1),
                     'conv_kernel_ms_per_frame': conv_ms, 'other_kernels_ms_per_frame': other_ms,
                     'ms_by_kernel_kind': {str(k): round(v, 4) for k, v in sorted(kind_ms.items())},
                     'frame_flops_over_frame_time_tflops': 2 * fmacs / (frame_ms * 1e-3) / 1e12},
    }
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_frames_per_s(args.workload, steps=1, warm=0, budget_s=25.0)
    print(json.dumps(out))


# ----------------------------------------------------------------------------------------------- CPU port
def cpu_frames_per_s(workload, steps, warm, budget_s=150.0):
    """Oracle port on the host cores.  A step is one steady-state frame of the workload restricted to a horizontal band
    of the image (full width, height H*f, f in {1/8, 1/4, 1/2, 1}): every layer does the same work per pixel, so
    frames/s = f / seconds-per-band.  f is the largest fraction for which steps + warm bands fit in `budget_s`
    (calibrated on one 1/8 band); previous-frame state is zero-filled (the cost does not depend on the content)."""
    import torch
    from oracle import generator_oracle as GO
    from vid2vid_b200 import networks as NW
    from vid2vid_b200.utils import synth_label_sequence
    wl = WORKLOADS[workload]
    # PyTorch's CPU convolutions stop scaling (and then regress) well before 128 threads on the GPU hosts
    cores = min(os.cpu_count(), int(os.environ.get('V2V_CPU_THREADS', '32')))
    torch.set_num_threads(cores)
    opt = make_opt_for(workload)
    opt.gpu_ids = []
    torch.manual_seed(0)
    sds = [NW.build_netG(opt, s).state_dict() for s in range(wl['n_scales'])]

    def run(frac, n, seed):
        H = max(32, int(wl['H'] * frac) // 32 * 32)
        orc = GO.ModelGOracle(opt, sds)
        orc.fake_B_prev = [torch.zeros(2, 3, H // 2 ** s, wl['W'] // 2 ** s) for s in range(wl['n_scales'])]
        seq = synth_label_sequence(n + 2, H, wl['W'], label_nc=35, block=32, seed=seed)
        t0 = time.time()
        with torch.no_grad():
            for t in range(n):
                orc.inference(seq[:, t:t + 3], seq[:, t:t + 3])
        return (time.time() - t0) / n, H

    run(1.0 / 16, 1, 0)                              # warm the thread pool / allocator
    t8, _ = run(1.0 / 8, 1, 1)                       # calibration
    frac = 1.0 / 16
    for f in (1.0, 0.5, 0.25, 0.125):
        if t8 * 8 * f * (steps + warm) <= budget_s:
            frac = f
            break
    if warm:
        run(frac, warm, 2)
    dt, H = run(frac, steps, 3)
    model = ''
    try:
        model = [l.split(':')[1].strip() for l in open('/proc/cpuinfo') if l.startswith('model name')][0]
    except Exception:
        pass
    eff = H / wl['H']
    return {'value': eff / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': '%d step(s), each one steady-state frame of a %dx%d band (%.3f of the %s frame), torch %s CPU fp32, %d threads, %s'
                      % (steps, wl['W'], H, eff, wl['desc'], torch.__version__, cores, model),
            'seconds_per_step': dt}


def run_reference(args, rank, world):
    if rank != 0:
        return
    K, Wm = args.steps, args.warmup
    r = cpu_frames_per_s(args.workload, steps=K, warm=min(Wm, 1), budget_s=180.0)
    wl = WORKLOADS[args.workload]
    out = {'impl': 'reference',
           'metric': 'frames/sec at 2048x1024 inference' if args.workload == 'cfg4' else 'frames/sec inference (%s)' % args.workload,
           'value': r['value'], 'unit': 'frames/s', 'n_gpus': world, 'steps': K, 'warmup': min(Wm, 1),
           'ms_per_step': 1e3 / r['value'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
           'data': 'synthetic 35-label blocky sequences, random-init weights',
           'config': {'workload': wl['desc'], 'parallelism': 'host CPU, rank 0 only'},
           'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
           'e2e': {'value': r['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    run_ours(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
