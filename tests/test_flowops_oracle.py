"""CPU: the plain-C restatement of the reference's native FlowNet2 ops (oracle/flowops_oracle.c).  The reference holds no
tests or vectors for these ops and its CUDA sources cannot be built here, so the restatement is pinned two ways:
against independent PyTorch formulations on random data, and against known-answer vectors worked out by hand from the CUDA
sources' text (index / clamp rules, channel order, normalisation and the 32-lane summation order)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import flowops


def test_resample2d_equals_grid_sample_align_corners_true():
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 3, 20, 28, generator=g)
    flow = torch.randn(2, 2, 20, 28, generator=g) * 3
    out = flowops.resample2d(img.numpy(), flow.numpy())
    ys, xs = torch.meshgrid(torch.arange(20.), torch.arange(28.), indexing='ij')
    gx = (xs + flow[:, 0]) / (28 - 1) * 2 - 1
    gy = (ys + flow[:, 1]) / (20 - 1) * 2 - 1
    ref = F.grid_sample(img, torch.stack([gx, gy], -1), mode='bilinear', padding_mode='border', align_corners=True)
    np.testing.assert_allclose(out, ref.numpy(), atol=2e-5)


def test_correlation_equals_shifted_dot_products():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(1, 40, 12, 16, generator=g)
    b = torch.randn(1, 40, 12, 16, generator=g)
    out = flowops.correlation(a.numpy(), b.numpy(), pad=20, k=1, max_disp=20, s1=1, s2=2)
    assert out.shape == (1, 441, 12, 16)
    bp = F.pad(b, (20, 20, 20, 20))
    ref = torch.zeros(1, 441, 12, 16)
    for tj in range(-10, 11):
        for ti in range(-10, 11):
            sh = bp[:, :, 20 + 2 * tj: 20 + 2 * tj + 12, 20 + 2 * ti: 20 + 2 * ti + 16]
            ref[:, (tj + 10) * 21 + (ti + 10)] = (a * sh).sum(1) / 40
    np.testing.assert_allclose(out, ref.numpy(), atol=2e-6)


def test_channelnorm_equals_l2_norm():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 9, 11, generator=g)
    np.testing.assert_allclose(flowops.channelnorm(x.numpy()), x.norm(dim=1, keepdim=True).numpy(), rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
# Known-answer vectors worked out BY HAND from the reference's CUDA sources (the reference ships none): they pin the
# index / clamp / channel-order / summation-order rules of the restatement to the kernels' text rather than to another
# implementation of ours.
def test_resample2d_hand_computed_floor_and_clamp_rules():
    """resample2d_kernel.cu:40-58: xf = x + dx; alpha = xf - floor(xf); xL = clamp(floor(xf)), xR = clamp(floor(xf) + 1) against
    the OUTPUT extent; val = (1-a)(1-b) v[yT][xL] + a(1-b) v[yT][xR] + (1-a) b v[yB][xL] + a b v[yB][xR]."""
    img = np.array([[[[1., 2., 3.], [4., 5., 6.]]]], np.float32)
    flow = np.zeros((1, 2, 2, 3), np.float32)
    flow[0, 0, 0, 0] = 0.5       # (y0,x0): xf 0.5 -> a .5, xL 0, xR 1            -> .5*1 + .5*2            = 1.5
    flow[0, 0, 0, 1] = -1.5      # (y0,x1): xf -0.5 -> floor -1, a .5, xL 0, xR 0 -> 1
    flow[0, 0, 0, 2] = 0.5       # (y0,x2): xf 2.5 -> a .5, xL 2, xR clamp(3)=2  -> 3
    flow[0, 1, 1, 0] = -1.0      # (y1,x0): yf 0 -> b 0, yT 0                     -> 1
    flow[0, 1, 1, 1] = 0.25      # (y1,x1): yf 1.25 -> b .25, yT 1, yB clamp(2)=1 -> 5
    flow[0, 0, 1, 2] = -0.5; flow[0, 1, 1, 2] = -0.5   # (y1,x2): xf 1.5, yf .5: .25*(2+3+5+6) = 4
    out = flowops.resample2d(img, flow)
    assert out.tolist() == [[[[1.5, 1.0, 3.0], [1.0, 5.0, 4.0]]]]


def test_channelnorm_hand_computed():
    x = np.array([[[[3., 0.]], [[4., 5.]], [[0., 12.]]]], np.float32)           # (1,3,1,2): sqrt(9+16), sqrt(25+144)
    assert flowops.channelnorm(x).tolist() == [[[[5.0, 13.0]]]]


def test_correlation_hand_computed_channel_order_and_normalisation():
    """correlation_cuda_kernel.cu:104-143 with pad 20, kernel 1, max_disp 20, stride2 2: 21 x 21 displacements, output channel
    tc = (tj + 10) * 21 + (ti + 10) with tj the VERTICAL displacement (outer), value / (C * k * k)."""
    a = np.zeros((1, 2, 3, 5), np.float32)
    b = np.zeros((1, 2, 3, 5), np.float32)
    a[0, :, 1, 1] = [2., 3.]
    b[0, :, 1, 1] = [5., 7.]          # same pixel: displacement (0, 0)          -> (2*5 + 3*7) / 2 = 15.5 at channel 220
    b[0, :, 1, 3] = [1., 1.]          # two pixels to the right: ti = +1 (s2 = 2) -> (2 + 3) / 2 = 2.5 at channel 10*21 + 11
    b[0, :, 2, 1] = [4., 0.]          # one pixel down is NOT on the stride-2 displacement grid -> never sampled
    b[0, :, 0, 1] = [0., 0.]
    out = flowops.correlation(a, b, pad=20, k=1, max_disp=20, s1=1, s2=2)
    assert out.shape == (1, 441, 3, 5)
    assert out[0, 220, 1, 1] == 15.5 and out[0, 10 * 21 + 11, 1, 1] == 2.5
    nz = np.argwhere(out != 0)
    assert sorted(map(tuple, nz.tolist())) == [(0, 220, 1, 1), (0, 221, 1, 1)]
    b2 = np.zeros_like(b)
    a2 = np.zeros_like(a)
    a2[0, :, 0, 1] = [1., 1.]
    b2[0, :, 2, 1] = [6., 2.]         # two pixels down from (0,1): tj = +1 -> channel 11*21 + 10 = 241, value (6 + 2) / 2 = 4
    out2 = flowops.correlation(a2, b2, pad=20, k=1, max_disp=20, s1=1, s2=2)
    assert out2[0, 11 * 21 + 10, 0, 1] == 4.0 and np.count_nonzero(out2) == 1


def test_correlation_hand_computed_summation_order():
    """correlation_cuda_kernel.cu:121-141: the block has 32 threads; thread `lane` sums channels lane, lane + 32, ... and a
    shuffle-down tree (warpReduceSum :16-21) joins the 32 partials.  fp32 vector built so that the order is visible: products
    1e8 (channel 0), 1 (channel 32, same lane 0: absorbed, 1e8 + 1 == 1e8 in fp32), -1e8 (channel 1, lane 1).  Per-lane-then-tree
    gives 1e8 + (-1e8) = 0; a sequential sum over channels would give (1e8 - 1e8 + ... + 1) / 64 = 1/64."""
    C = 64
    a = np.zeros((1, C, 1, 1), np.float32)
    b = np.ones((1, C, 1, 1), np.float32)
    a[0, 0], a[0, 32], a[0, 1] = 1e8, 1.0, -1e8
    out = flowops.correlation(a, b, pad=20, k=1, max_disp=20, s1=1, s2=2)
    assert out[0, 220, 0, 0] == 0.0
    assert np.float32(np.float32(np.float32(1e8) + np.float32(-1e8)) + np.float32(1.0)) / np.float32(C) == np.float32(1.0 / 64)
