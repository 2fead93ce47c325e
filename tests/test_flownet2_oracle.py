"""Pins oracle/flownet2_oracle.py (the reference flow network behind the training targets).
  * against the unmodified reference modules (models/flownet2_pytorch/models.py, models/flownet.py) when /root/reference
    is present -- the three native ops are CUDA-only there and are supplied by oracle/flowops_oracle.c on both sides,
    so this pins the network wiring, layer hyper-parameters and the wrapper logic, not those ops;
  * against the committed fixture tests/golden/flownet2_small.npz (reference output on seeded inputs with weights
    regenerated from (seed, key, shape)) everywhere else.  CPU only."""
import json
import os
import types

import numpy as np
import pytest
import torch

from oracle import flownet2_oracle as FO
from oracle import ref_shim

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def _close(a, b, name, tol=2e-4):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= tol * scale, '%s: max |d| = %.3e (scale %.2f)' % (name, err, scale)


def test_flownet2_oracle_vs_golden():
    keys = json.load(open(os.path.join(GOLD, 'flownet2_keys.json')))
    assert len(keys) == 220 and sum(int(np.prod(s)) for _, s in keys) == 162518834      # models.py:16 parameter count
    g = np.load(os.path.join(GOLD, 'flownet2_small.npz'))
    sd = FO.det_state_dict(keys, seed=7)
    with torch.no_grad():
        flow = FO.flownet2(sd, torch.from_numpy(g['pair']))
        wflow, wconf = FO.flow_and_conf(sd, torch.from_numpy(g['im1']), torch.from_numpy(g['im2']))
    _close(flow, g['flow'], 'FlowNet2 flow')
    _close(wflow, g['wflow'], 'wrapper flow (80 -> 64 rows and back)')
    # the confidence is a thresholded mask: allow the few pixels that sit on the 0.02 threshold to flip
    assert float(np.abs(wconf.numpy() - g['wconf']).mean()) < 2e-3


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present (GPU box)')
def test_flownet2_oracle_vs_reference_random_init():
    F2 = ref_shim.flownet2_class()
    torch.manual_seed(3)
    net = F2().eval()                                   # the reference's own Xavier / uniform-bias initialisation
    sd = net.state_dict()
    pair = torch.rand(2, 3, 2, 64, 64)
    with torch.no_grad():
        ref = net(pair)
        out = FO.flownet2(sd, pair)
    _close(out, ref, 'FlowNet2 (batch 2, 64x64)')
    # sub-networks on their own inputs (eval mode returns a 1-tuple)
    x6, x12 = torch.rand(1, 6, 64, 128) - 0.5, torch.rand(1, 12, 64, 128) - 0.5
    with torch.no_grad():
        _close(FO.flownet_c(sd, 'flownetc.', x6), net.flownetc(x6)[0], 'FlowNetC')
        _close(FO.flownet_s(sd, 'flownets_1.', x12), net.flownets_1(x12)[0], 'FlowNetS')
        _close(FO.flownet_sd(sd, 'flownets_d.', x6), net.flownets_d(x6)[0], 'FlowNetSD')
    xf = torch.rand(1, 11, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        _close(FO.flownet_fusion(sd, 'flownetfusion.', xf), net.flownetfusion(xf), 'FlowNetFusion')
    # the vid2vid wrapper (models/flownet.py:43-58), un-resized and resized paths
    from models.flownet import FlowNet                                                        # noqa (reference module)
    from models.flownet2_pytorch.networks.resample2d_package.resample2d import Resample2d     # noqa
    stub = types.SimpleNamespace(flowNet=net, resample=Resample2d())
    stub.norm = lambda t: FlowNet.norm(stub, t)
    for h, w in ((64, 64), (70, 64)):
        im1 = torch.rand(1, 3, h, w)
        im2 = torch.roll(im1, shifts=(1, 1), dims=(2, 3))
        with torch.no_grad():
            rf, rc = FlowNet.compute_flow_and_conf(stub, im1, im2)
            of, oc = FO.flow_and_conf(sd, im1, im2)
        _close(of, rf, 'wrapper flow %dx%d' % (h, w))
        assert float((oc - rc).abs().mean()) < 2e-3
