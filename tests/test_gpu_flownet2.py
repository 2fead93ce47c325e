"""GPU parity of the FlowNet2 stack through the plan runtime, precise (split-bf16 x3) mode:
against the committed reference fixture (tests/golden/flownet2_small.npz: the unmodified reference FlowNet2 and the vid2vid
FlowNet wrapper on seeded inputs) and against the oracle on larger seeded inputs, whole cascade and each sub-network."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import flownet2_oracle as FO
from vid2vid_b200 import flownet as FN

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
# Stated tolerance: flow in pixels, max |d| <= 2e-3 * max(1, max|ref|) (the cascade is ~60 convolutions deep with no norm layers)
TOL = 2e-3


def _net():
    keys = json.load(open(os.path.join(GOLD, 'flownet2_keys.json')))
    sd = FO.det_state_dict(keys, seed=7)
    net = FN.FlowNet2()
    net.load_state_dict(sd)
    return net.cuda(), sd


def _close(a, b, name, tol=TOL):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    assert torch.isfinite(a).all(), name
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    print('%-34s max|d|=%.3e mean|d|=%.3e  ref max %.3f rms %.3f' % (name, err, float((a - b).abs().mean()), float(b.abs().max()),
                                                                    float(b.pow(2).mean().sqrt())))
    assert err <= tol * scale, '%s: max |d| = %.3e (scale %.2f)' % (name, err, scale)


def test_flownet2_vs_reference_fixture():
    net, _ = _net()
    g = np.load(os.path.join(GOLD, 'flownet2_small.npz'))
    with torch.no_grad():
        flow = net(torch.from_numpy(g['pair']).cuda())
        flow2 = net(torch.from_numpy(g['pair']).cuda())
    assert torch.equal(flow, flow2)
    _close(flow, g['flow'], 'FlowNet2 flow (64x128)')
    w = FN.FlowNet()
    w.flowNet = net
    wflow, wconf = w.compute_flow_and_conf(torch.from_numpy(g['im1']).cuda(), torch.from_numpy(g['im2']).cuda())
    _close(wflow, g['wflow'], 'wrapper flow (80 -> 64 rows and back)')
    assert float(np.abs(wconf.cpu().numpy() - g['wconf']).mean()) < 5e-3       # thresholded mask: pixels on the 0.02 edge may flip
    f5, c5 = w.forward(torch.from_numpy(g['im1']).cuda().unsqueeze(1), torch.from_numpy(g['im2']).cuda().unsqueeze(1))
    assert f5.shape == (1, 1, 2, 80, 64) and c5.shape == (1, 1, 1, 80, 64)
    assert torch.equal(f5[:, 0], wflow)


def test_flownet2_subnetworks_vs_oracle():
    """Each sub-network on its own seeded input (batch 2, 128x192) against the oracle."""
    net, sd = _net()
    from vid2vid_b200.plan import Plan
    g = torch.Generator().manual_seed(5)
    N, H, W = 2, 128, 192
    x6 = torch.rand(N, 6, H, W, generator=g) - 0.5
    x12 = torch.rand(N, 12, H, W, generator=g) - 0.5
    x11 = torch.rand(N, 11, H, W, generator=g) - 0.5
    dev = torch.device('cuda', 0)
    new = lambda c, h, w: torch.empty((N, c, h, w), device=dev)
    with torch.no_grad():
        out = new(2, H // 4, W // 4)
        net._sub_plan('flownetc', N, H, W, dev).run([x6.cuda(), out])
        _close(out, FO.flownet_c(sd, 'flownetc.', x6), 'FlowNetC flow2')
        out = new(2, H // 4, W // 4)
        xs = x12.cuda()
        net._sub_plan('flownets_1', N, H, W, dev).run([xs[:, :6].contiguous(), xs[:, 6:9].contiguous(), xs[:, 9:11].contiguous(),
                                                       xs[:, 11:12].contiguous(), out])
        _close(out, FO.flownet_s(sd, 'flownets_1.', x12), 'FlowNetS flow2')
        out = new(2, H // 4, W // 4)
        net._sub_plan('flownets_d', N, H, W, dev).run([x6.cuda(), out])
        _close(out, FO.flownet_sd(sd, 'flownets_d.', x6), 'FlowNetSD flow2')
        out = new(2, H, W)
        xf = x11.cuda()
        # fusion stack: x[:, :3] comes from a 6-channel tensor in the product path
        x6f = torch.cat([xf[:, :3], torch.zeros(N, 3, H, W, device=dev)], 1).contiguous()
        net._sub_plan('flownetfusion', N, H, W, dev).run([x6f, xf[:, 3:5].contiguous(), xf[:, 5:7].contiguous()] +
                                                         [xf[:, 7 + i:8 + i].contiguous() for i in range(4)] + [out])
        _close(out, FO.flownet_fusion(sd, 'flownetfusion.', x11), 'FlowNetFusion flow')


def test_flownet2_cascade_vs_oracle_batch2():
    net, sd = _net()
    g = torch.Generator().manual_seed(6)
    base = torch.rand(2, 3, 128, 192, generator=g)
    pair = torch.stack([base, torch.roll(base, shifts=(2, 3), dims=(2, 3)) * 0.95 + 0.02], 2)       # (2,3,2,H,W)
    with torch.no_grad():
        ref = FO.flownet2(sd, pair)
        out = net(pair.cuda())
    _close(out, ref, 'FlowNet2 cascade (2 x 128x192)')
    w = FN.FlowNet()
    w.flowNet = net
    with torch.no_grad():
        rf, rc = FO.flow_and_conf(sd, pair[:, :, 0].contiguous(), pair[:, :, 1].contiguous())
        of, oc = w.compute_flow_and_conf(pair[:, :, 0].contiguous().cuda(), pair[:, :, 1].contiguous().cuda())
    _close(of, rf, 'wrapper flow 128x192')
    assert float((oc.cpu() - rc).abs().mean()) < 5e-3
