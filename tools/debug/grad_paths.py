"""Debug: which output path of CompositeGenerator carries the gradient error (one output's gradient at a time)."""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases as C
from oracle import generator_oracle as GO
from vid2vid_b200.utils import det_fill_
c = C.CASES['g0_small']
inp, img_prev, mask = C.gen_inputs(c['label_nc'], c['h'], c['w'], c['seed'])
gs = [torch.randn(1, ch, c['h'], c['w'], generator=torch.Generator().manual_seed(20 + i)) for i, ch in enumerate((3, 2, 1, 3, 16, 16, 8))]
keys = ['model_up_flow.7.bias', 'model_up_flow.4.bias', 'model_up_img.7.bias', 'model_up_img.4.bias', 'model_down_seg.2.bias', 'indv_up.4.bias', 'model_final_flow.1.weight']
for k_out in range(7):
    net = det_fill_(C.build_module(c), seed=c['seed'])
    sd = {k: v.clone().double().requires_grad_(k.split('.')[-1] in ('weight', 'bias')) for k, v in net.state_dict().items()}
    torch.set_default_dtype(torch.float64)
    ref = GO.composite_generator(sd, inp.double(), img_prev.double(), mask.double(), False, n_downsampling=3, n_blocks=9, use_fg_model=True)
    (ref[k_out] * gs[k_out].double()).sum().backward()
    torch.set_default_dtype(torch.float32)
    net = net.cuda(); net.precision = 'precise'
    out = net(inp.cuda(), img_prev.cuda(), mask.cuda(), None, None, None, False)
    (out[k_out] * gs[k_out].cuda()).sum().backward()
    line = []
    for k in keys:
        p = dict(net.named_parameters())[k].grad
        r = sd[k].grad
        if r is None or p is None:
            line.append('%s: none' % k); continue
        line.append('%s %.1e (|ref| %.1e)' % (k.replace('model_', ''), ((p.double().cpu() - r).norm() / max(r.norm().item(), 1e-30)).item(), r.norm().item()))
    print(C.GEN_OUT_NAMES[k_out], ' | '.join(line), flush=True)
