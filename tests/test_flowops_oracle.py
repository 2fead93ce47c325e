"""CPU: the plain-C restatement of the reference's native FlowNet2 ops (oracle/flowops_oracle.c) against
independent PyTorch formulations.  The reference holds no vectors for these ops, so this
cross-check is the pin the restatement gets ("parity unpinned" by reference tests)."""
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
