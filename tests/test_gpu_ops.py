"""GPU parity of the stand-alone operators (C ABI group 1) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import flowops
from oracle import generator_oracle as GO
from vid2vid_b200 import ops
from vid2vid_b200.utils import synth_label_sequence

pytestmark = pytest.mark.gpu


def test_correlation_flownetc_params():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 64, 20, 36, generator=g)
    b = torch.randn(2, 64, 20, 36, generator=g)
    out = ops.Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)(a.cuda(), b.cuda())
    ref = flowops.correlation(a.numpy(), b.numpy())
    # fp32 dot products of length 64 in a different summation order: tolerance 1e-5 absolute (values O(0.1))
    np.testing.assert_allclose(out.cpu().numpy(), ref, atol=1e-5, rtol=1e-5)


def test_correlation_256_channels():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(1, 256, 8, 40, generator=g)
    b = torch.randn(1, 256, 8, 40, generator=g)
    out = ops.correlation(a.cuda(), b.cuda())
    np.testing.assert_allclose(out.cpu().numpy(), flowops.correlation(a.numpy(), b.numpy()), atol=2e-5, rtol=1e-5)


def test_resample2d_bit_exact():
    g = torch.Generator().manual_seed(2)
    img = torch.randn(2, 3, 33, 47, generator=g)
    flow = torch.randn(2, 2, 33, 47, generator=g) * 4
    out = ops.Resample2d()(img.cuda(), flow.cuda())
    ref = flowops.resample2d(img.numpy(), flow.numpy())
    assert np.array_equal(out.cpu().numpy(), ref)      # same expression order, double weights -> bit exact


def test_channelnorm_bit_exact():
    g = torch.Generator().manual_seed(3)
    for c in (2, 3):
        x = torch.randn(2, c, 19, 23, generator=g)
        out = ops.ChannelNorm()(x.cuda())
        assert np.array_equal(out.cpu().numpy(), flowops.channelnorm(x.numpy()))


@pytest.mark.parametrize('ac', [False, True])
def test_resample_grid_sample(ac):
    g = torch.Generator().manual_seed(4)
    img = torch.randn(2, 3, 24, 40, generator=g)
    flow = torch.randn(2, 2, 24, 40, generator=g) * 3
    out = ops.resample(img.cuda(), flow.cuda(), align_corners=ac)
    ref = GO.resample(img, flow, align_corners=ac)
    # fp32 coordinate arithmetic mirrors ATen; tolerance 2e-5 absolute on N(0,1) images
    assert (out.cpu() - ref).abs().max().item() < 2e-5


def test_onehot_edges_pyramid_mask_exact():
    lab = synth_label_sequence(3, 32, 48, label_nc=35, block=4, seed=5)
    ref = GO.encode_input(lab, lab, 35, True)
    out = ops.onehot_edges(lab.cuda(), lab.cuda(), 35, True)
    assert torch.equal(out.cpu(), ref)
    pyr_ref = GO.build_pyr(ref, 3)
    p1 = ops.avgpool3s2(out)
    p2 = ops.avgpool3s2(p1)
    assert (p1.cpu() - pyr_ref[1]).abs().max().item() < 1e-6
    assert (p2.cpu() - pyr_ref[2]).abs().max().item() < 1e-6
    m = ops.fg_mask(p1, 2, [26, 3])
    assert (m.cpu() - GO.compute_mask(pyr_ref[1], 2, [26, 3])).abs().max().item() < 1e-6
