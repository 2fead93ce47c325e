import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, torch.nn as nn
import bf16_emul as E
from vid2vid_b200 import networks as NW
from vid2vid_b200.utils import det_fill_
BN = NW.get_norm_layer('batch')
mods = NW._down(16, 32, BN) + NW._down(32, 64, BN) + NW._up(64, 32, BN) + NW._up(32, 16, BN)
runner = det_fill_(NW.SequentialRunner(mods), seed=5).cuda(); runner.precision = 'precise'
x = torch.randn(1, 16, 32, 64, generator=torch.Generator().manual_seed(1)).cuda()
xr = x.clone().requires_grad_(True)
out = runner(xr)
g = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda()
(out * g).sum().backward()
ours = runner.seq[10].bias.grad.clone()
manual = (g * (out > 0)).sum(dim=(0, 2, 3))
E.ROUND[0], E.GRAD[0] = False, True
for p in runner.parameters(): p.grad = None
xd = x.clone().double().requires_grad_(True)
ref = E.run_units(list(runner.seq), xd)
(ref * g.double()).sum().backward()
refg = runner.seq[10].bias.grad.clone()
manual_ref = (g.double() * (ref > 0)).sum(dim=(0, 2, 3))
print('ours  ', ours[:6].tolist())
print('manual', manual[:6].tolist())
print('ref   ', refg[:6].tolist())
print('manref', manual_ref[:6].tolist())
print('mask mismatches', ((out > 0) != (ref > 0)).sum().item(), 'of', out.numel(), ' max|out-ref|', (out.double() - ref).abs().max().item())
print('diff ours-ref', (ours - refg).abs().tolist())
print('diff manual-ref', (manual - refg).abs().tolist())
names = [n for n, _ in runner.named_parameters()]
