"""CPU study (test infrastructure, uses oracle/): how much output error each candidate tensor-core operand
format of the conv stack would cause, frame by frame through the recurrence, against an fp64 evaluation of the same
oracle.  Convolutions are evaluated in fp64 on operands rounded the way the format would round them, so only the operand
precision differs.   python tools/precision_study.py [frames] [ngf] [H] [W] [scales] [flow-head scale] [modes]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import generator_oracle as GO      # noqa: E402
from vid2vid_b200 import networks as NW        # noqa: E402
from vid2vid_b200.utils import make_opt, synth_label_sequence   # noqa: E402


def rnd(x, kind):
    if kind == 'bf16':
        return x.float().bfloat16().double()
    if kind == 'fp16':
        return x.float().half().double()
    if kind == 'tf32':     # truncation to 10 mantissa bits
        xi = x.float().view(torch.int32) & ~0x1FFF
        return xi.view(torch.float32).double()
    if kind == 'fp32':
        return x.float().double()
    raise ValueError(kind)


def split(x, kind):
    hi = rnd(x, kind)
    lo = rnd(x.float().double() - hi, kind)
    return hi, lo


MODE = ['fp64']


def conv_any(fn, x, w, b, **kw):
    m = MODE[0]
    if m == 'fp64':
        return fn(x, w, b, **kw)
    if m in ('bf16', 'fp16', 'tf32', 'fp32'):
        return fn(rnd(x, m), rnd(w, m), b, **kw)
    if m in ('bf16x3', 'fp16x3'):
        k = m[:4]
        xh, xl = split(x, k)
        wh, wl = split(w, k)
        return fn(xh, wh, b, **kw) + fn(xl, wh, None, **kw) + fn(xh, wl, None, **kw)
    raise ValueError(m)


def _conv(x, sd, p, stride=1, padding=0):
    return conv_any(F.conv2d, x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def _deconv(x, sd, p):
    return conv_any(F.conv_transpose2d, x, sd[p + '.weight'], sd[p + '.bias'], stride=2, padding=1, output_padding=1)


GO._conv, GO._deconv = _conv, _deconv


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    ngf = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    H = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    W = int(sys.argv[4]) if len(sys.argv) > 4 else 128
    S = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    fscale = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
    modes = sys.argv[7].split(',') if len(sys.argv) > 7 else ['fp64', 'fp32', 'fp16x3', 'bf16x3', 'fp16', 'tf32', 'bf16']
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(8)
    opt = make_opt(label_nc=35, use_instance=True, fg=True, fg_labels=[26], n_scales_spatial=S, ngf=ngf, use_single_G=True,
                   loadSize=512, dataroot='City', gpu_ids=[], no_first_img=True)
    torch.manual_seed(0)
    sds = [{k: v.double() for k, v in NW.build_netG(opt, s).state_dict().items()} for s in range(S)]
    seq = synth_label_sequence(frames + 3, H, W, label_nc=35, block=16, seed=3).double()
    for sd in sds:      # optional: shrink the random flow heads (tests/cases.py condition_flow_heads)
        for k in sd:
            if k.startswith('model_final_flow'):
                sd[k] = sd[k] * fscale
    outs = {}
    for m in modes:
        MODE[0] = m
        orc = GO.ModelGOracle(opt, sds)
        res = []
        with torch.no_grad():
            for t in range(frames):
                fb, _ = orc.inference(seq[:, t:t + 3], seq[:, t:t + 3])
                res.append(fb.clone())
        outs[m] = res
        if m != 'fp64':
            line = ' '.join('%.1e/%.1e' % ((r - q).abs().max().item(), (r - q).abs().mean().item()) for r, q in zip(res, outs['fp64']))
            print('%-7s max/mean |d img| per frame: %s' % (m, line), flush=True)


if __name__ == '__main__':
    main()
