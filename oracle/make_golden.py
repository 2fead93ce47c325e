"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(imported from /root/reference through oracle/ref_shim.py) on the seeded cases of
tests/cases.py.  Run in the build container:   python -m oracle.make_golden [case ...]

The fixtures are what pins oracle/generator_oracle.py and the CUDA path on the GPU box,
where the reference tree does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from oracle import ref_shim                                     # noqa: E402
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cases as C                                               # noqa: E402
from vid2vid_b200.utils import make_opt, det_fill_, synth_label_sequence   # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def coarse_feats(c):
    g = torch.Generator().manual_seed(c['seed'] + 5)
    h2, w2, ngf = c['h'] // 2, c['w'] // 2, c['ngf']
    return (torch.randn(1, 2 * ngf, h2, w2, generator=g), torch.randn(1, 2 * ngf, h2, w2, generator=g),
            torch.randn(1, ngf, h2, w2, generator=g))


def run_case(name, c):
    N = ref_shim.networks()
    kind = c['kind']
    out = {}
    torch.manual_seed(0)
    if c.get('align_corners'):
        # original PyTorch-0.4 semantics: force the flag the reference omits
        import torch.nn.functional as F
        orig = F.grid_sample
        F.grid_sample = lambda *a, **k: orig(*a, **{**k, 'align_corners': True})
    try:
        with torch.no_grad():
            if kind in ('composite', 'compositeLocal'):
                nc = 3 * (c['label_nc'] + 1)
                opt = make_opt(ngf=c['ngf'], n_blocks=c.get('n_blocks', 9), n_blocks_local=c.get('n_blocks_local', 3),
                               fg=c['fg'], no_flow=c.get('no_flow', False), n_downsample_G=c['nd'], gpu_ids=[])
                net = det_fill_(N.define_G(nc, 3, 6, c['ngf'], kind, c['nd'], 'batch', c.get('scale', 0), [], opt),
                                seed=c['seed'])
                inp, img_prev, mask = C.gen_inputs(c['label_nc'], c['h'], c['w'], c['seed'], block=c.get('block', 4))
                if kind == 'composite':
                    res = net(inp, img_prev, mask, None, None, None, False)
                else:
                    res = net(inp, img_prev, mask, *coarse_feats(c), False)
                ss = c.get('subsample', 1)
                for nme, t in zip(C.GEN_OUT_NAMES, res):
                    if t is None:
                        continue
                    if ss > 1 and t.shape[1] > 3:     # big feature maps: keep a strided sample + channel means
                        out[nme + '_sub'] = _np(t[:, :, ::ss, ::ss])
                        out[nme + '_cmean'] = _np(t.mean(dim=(2, 3)))
                    else:
                        out[nme] = _np(t)
            elif kind in ('global', 'local'):
                opt = make_opt(n_blocks=c['n_blocks'], n_blocks_local=c.get('n_blocks_local', 3), gpu_ids=[])
                net = det_fill_(N.define_G(c['input_nc'], 3, 0, c['ngf'], kind, c['nd'], 'instance', 0, [], opt),
                                seed=c['seed'])
                g = torch.Generator().manual_seed(c['seed'] + 1)
                lab = synth_label_sequence(1, c['h'], c['w'], label_nc=c['input_nc'], block=4, seed=c['seed'])
                x = torch.zeros(1, c['input_nc'], c['h'], c['w']).scatter_(1, lab[:, 0].long(), 1.0)
                out['out'] = _np(net(x))
            elif kind == 'D':
                net = det_fill_(N.define_D(c['input_nc'], c['ndf'], c['n_layers'], 'batch', c['num_D'], True, []),
                                seed=c['seed'])
                g = torch.Generator().manual_seed(c['seed'] + 1)
                x = torch.randn(c['batch'], c['input_nc'], c['h'], c['w'], generator=g)
                for i, tower in enumerate(net(x)):
                    for j, t in enumerate(tower):
                        out['t%d_l%d' % (i, j)] = _np(t)
            elif kind == 'inference':
                opt = C.inference_opt(c)
                single = det_fill_(N.define_G(c['label_nc'], 3, 0, 16, 'global', 2, 'instance', 0, [], opt),
                                   seed=c['seed'] + 100)
                m = ref_shim.make_model_G(opt, single)
                for s in range(c['n_scales']):
                    det_fill_(getattr(m, 'netG%d' % s), seed=c['seed'] + s)
                    C.condition_flow_heads(getattr(m, 'netG%d' % s), c['flow_weight_scale'])
                tG = opt.n_frames_G
                seq = synth_label_sequence(c['n_gen'] + tG - 1, c['h'], c['w'], label_nc=c['label_nc'], block=8,
                                           seed=c['seed'])
                for t in range(c['n_gen']):
                    A = seq[:, t:t + tG]
                    fake_B, real_A = m.inference(A, None, A)
                    out['fake_B_%d' % t] = _np(fake_B)
                for si in range(c['n_scales']):
                    out['prev_state_%d' % si] = _np(m.fake_B_prev[si])
            else:
                raise ValueError(kind)
    finally:
        if c.get('align_corners'):
            F.grid_sample = orig
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **out)
    print('%-14s %7.1f KB  %s' % (name, os.path.getsize(path) / 1024, sorted(out)))


def dump_keys():
    """state_dict key / shape lists of the reference modules -> tests/golden/state_dict_keys.json."""
    import json
    N = ref_shim.networks()
    out = {}
    for name, c in C.KEY_CASES.items():
        kind = c['kind']
        if kind in ('composite', 'compositeLocal'):
            opt = make_opt(ngf=c['ngf'], n_blocks=c.get('n_blocks', 9), n_blocks_local=c.get('n_blocks_local', 3),
                           fg=c['fg'], no_flow=c.get('no_flow', False), n_downsample_G=c['nd'], gpu_ids=[])
            net = N.define_G(3 * (c['label_nc'] + 1), 3, 6, c['ngf'], kind, c['nd'], 'batch', c.get('scale', 0), [], opt)
        elif kind in ('global', 'local'):
            opt = make_opt(n_blocks=c['n_blocks'], n_blocks_local=c.get('n_blocks_local', 3), gpu_ids=[])
            net = N.define_G(c['input_nc'], 3, 0, c['ngf'], kind, c['nd'], 'instance', 0, [], opt)
        else:
            net = N.define_D(c['input_nc'], c['ndf'], c['n_layers'], 'batch', c['num_D'], True, [])
        out[name] = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    with open(os.path.join(OUT, 'state_dict_keys.json'), 'w') as f:
        json.dump(out, f)
    print('state_dict_keys.json', {k: len(v) for k, v in out.items()})


def flownet2_golden():
    """Reference FlowNet2 (with the op stand-ins of ref_shim.install_flownet2_ops) and the FlowNet wrapper's
    compute_flow_and_conf on seeded inputs, with the deterministic weights of flownet2_oracle.det_state_dict ->
    tests/golden/flownet2_keys.json + flownet2_small.npz."""
    import json
    import types
    from oracle import flownet2_oracle as FO
    F2 = ref_shim.flownet2_class()
    net = F2().eval()
    keys = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    with open(os.path.join(OUT, 'flownet2_keys.json'), 'w') as f:
        json.dump(keys, f)
    net.load_state_dict(FO.det_state_dict(keys, seed=7))
    g = torch.Generator().manual_seed(11)
    pair = torch.rand(1, 3, 2, 64, 128, generator=g)
    base = torch.rand(1, 3, 80, 64, generator=g)                    # height 80: exercises the resize-to-64 path
    im1, im2 = base, torch.roll(base, shifts=(1, 2), dims=(2, 3)) * 0.9 + 0.05
    from models.flownet import FlowNet                              # noqa (reference module)
    from models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d   # noqa
    stub = types.SimpleNamespace(flowNet=net, resample=Resample2d())
    stub.norm = lambda t: FlowNet.norm(stub, t)
    with torch.no_grad():
        flow = net(pair)
        wflow, wconf = FlowNet.compute_flow_and_conf(stub, im1, im2)
    out = dict(pair=_np(pair), flow=_np(flow), im1=_np(im1), im2=_np(im2), wflow=_np(wflow), wconf=_np(wconf))
    path = os.path.join(OUT, 'flownet2_small.npz')
    np.savez_compressed(path, **out)
    print('flownet2_small %7.1f KB flow rms %.3f wflow rms %.3f conf mean %.3f' % (
        os.path.getsize(path) / 1024, float(flow.pow(2).mean().sqrt()), float(wflow.pow(2).mean().sqrt()), float(wconf.mean())))


if __name__ == '__main__':
    names = sys.argv[1:] or (list(C.CASES) + ['keys'])
    for n in names:
        if n == 'keys':
            dump_keys()
        elif n == 'flownet2':
            flownet2_golden()
        else:
            run_case(n, C.CASES[n])
