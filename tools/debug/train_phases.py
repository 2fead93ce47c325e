"""CPU issue time vs GPU completion time of the phases of one training step (is the step launch bound, and where?)."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, ROOT)
import torch
import bench
from vid2vid_b200 import flownet as FN
from vid2vid_b200.model_d import Vid2VidModelD
from vid2vid_b200.model_g import Vid2VidModelG
from vid2vid_b200.trainer import Trainer
from vid2vid_b200.utils import make_opt, synth_label_sequence

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
H, Wd = wl['H'], wl['W']
opt = make_opt(label_nc=35, use_instance=True, fg=True, fg_labels=[26], n_scales_spatial=wl['n_scales'], ngf=wl['ngf'], num_D=3,
               n_scales_temporal=2, n_frames_D=3, isTrain=True, no_vgg=True, gpu_ids=[0], n_frames_total=30, dataroot='datasets/Cityscapes/', loadSize=Wd)
torch.manual_seed(1234)
G = Vid2VidModelG().initialize(opt); D = Vid2VidModelD().initialize(opt); F = FN.FlowNet().initialize(opt)
tr = Trainer(opt, G, D, F, world=1)
tG = opt.n_frames_G
A = synth_label_sequence(14, H, Wd, label_nc=35, block=64, seed=0).cuda()
B = torch.rand(1, 14, 3, H, Wd).cuda() * 2 - 1
for t in range(5):
    tr.step(A[:, t:t + tG], B[:, t:t + tG], A[:, t:t + tG])
torch.cuda.synchronize()


def phase(name, fn, acc):
    t0 = time.perf_counter()
    r = fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0])
    a[0] += (t1 - t0) * 1e3; a[1] += (t2 - t0) * 1e3
    return r


acc = {}
n = 4
for t in range(5, 5 + n):
    a, b = A[:, t:t + tG], B[:, t:t + tG]
    lg, ld, ldt, _, _ = phase('forward + losses', lambda: tr.losses(a, b, a), acc)
    phase('zero grads', lambda: tr.grads.zero(), acc)
    phase('loss_G.backward', lambda: lg.backward(), acc)
    phase('zero D grads', lambda: [tr.grads.zero(g) for g in range(1, len(tr.grads.groups))], acc)
    phase('loss_D.backward', lambda: ld.backward(), acc)
    phase('loss_D_T.backward', lambda: [x.backward() for x in ldt], acc)
    phase('optimizers', lambda: (G.optimizer_G.step(), D.optimizer_D.step(), [getattr(D, 'optimizer_D_T%d' % s).step() for s in range(len(ldt))]), acc)
print('%-22s %12s %16s' % ('phase', 'CPU issue ms', 'until GPU done ms'))
for k, (c, g) in acc.items():
    print('%-22s %12.1f %16.1f' % (k, c / n, g / n))
print('%-22s %12.1f %16.1f' % ('sum', sum(v[0] for v in acc.values()) / n, sum(v[1] for v in acc.values()) / n))
t0 = time.perf_counter()
for t in range(9, 12):
    tr.step(A[:, t:t + tG], B[:, t:t + tG], A[:, t:t + tG])
torch.cuda.synchronize()
print('plain step: %.1f ms' % ((time.perf_counter() - t0) / 3 * 1e3))

# CUDA runtime / torch op view of one step: which host-side calls take the time (cudaMalloc? synchronisations? launches?)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t = 12
    tr.step(A[:, t - 1:t - 1 + tG], B[:, t - 1:t - 1 + tG], A[:, t - 1:t - 1 + tG])
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.self_cpu_time_total)
print('top host-side entries by self CPU time (one step)')
for r in rows[:28]:
    print('%9.2f ms %6d x  %s' % (r.self_cpu_time_total / 1e3, r.count, r.key[:90]))
