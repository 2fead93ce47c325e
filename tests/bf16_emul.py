"""Test helper: PyTorch evaluation of a layer list with bf16 rounding inserted at exactly the points
where the CUDA path stores bf16 (inputs, packed weights, raw conv outputs, activations), fp32
everywhere else.  Against this emulation the kernels must agree to ~1 bf16 ulp, which separates
"kernel is wrong" from "bf16 operands differ from the fp32 oracle" (the stated end-to-end tolerance)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from vid2vid_b200 import networks as NW
from vid2vid_b200 import _lib as L


ROUND = [True]      # False: no rounding anywhere and fp64 arithmetic -- the exact reference the precise mode is compared with
GRAD = [False]      # True: keep the autograd graph to the module parameters (reference gradients for the backward kernels)


def _param(t):
    return t if GRAD[0] else t.detach()


def r16(t):
    if not ROUND[0]:
        return t.double()
    return t.to(torch.bfloat16).to(torch.float32)


def _act(x, act, slope):
    if act == L.ACT_RELU:
        return F.relu(x)
    if act == L.ACT_LRELU:
        return F.leaky_relu(x, slope)
    if act == L.ACT_TANH:
        return torch.tanh(x)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def _conv(x, conv, pmode, pad, with_bias):
    w = r16(_param(conv.weight).float())
    b = _param(conv.bias).to(w.dtype) if (with_bias and conv.bias is not None) else None
    if isinstance(conv, nn.ConvTranspose2d):
        return F.conv_transpose2d(x, w, b, stride=conv.stride, padding=conv.padding, output_padding=conv.output_padding)
    if pmode == L.PAD_REFLECT and pad:
        return F.conv2d(F.pad(x, (pad,) * 4, mode='reflect'), w, b, stride=conv.stride)
    return F.conv2d(x, w, b, stride=conv.stride, padding=conv.padding)


def _norm(raw, norm):
    """Statistics from the fp32 accumulators, applied to the bf16-stored raw values."""
    if isinstance(norm, nn.BatchNorm2d):
        mean = raw.mean(dim=(0, 2, 3), keepdim=True)
        var = raw.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        g, b = _param(norm.weight).view(1, -1, 1, 1).to(raw.dtype), _param(norm.bias).view(1, -1, 1, 1).to(raw.dtype)
    else:
        mean = raw.mean(dim=(2, 3), keepdim=True)
        var = raw.var(dim=(2, 3), unbiased=False, keepdim=True)
        g, b = 1.0, 0.0
    scale = g / torch.sqrt(var + norm.eps)
    shift = b - mean * scale
    return r16(raw) * scale + shift


def run_units(mods, x, adds=()):
    """x already bf16-rounded fp32.  Returns the bf16-rounded output activation."""
    units = NW._units(mods)
    for k, u in enumerate(units):
        last = k == len(units) - 1
        if u[0] == 'conv':
            _, conv, pmode, pad, norm, act, slope = u
            if norm is None:
                x = r16(_act(_conv(x, conv, pmode, pad, True), act, slope))
            else:
                y = _act(_norm(_conv(x, conv, pmode, pad, False), norm), act, slope)
                if last:
                    for a in adds:
                        y = y + a
                x = r16(y)
        else:
            cb = u[1].conv_block
            h = r16(_act(_norm(_conv(x, cb[1], L.PAD_REFLECT, 1, False), cb[2]), L.ACT_RELU, 0))
            y = _norm(_conv(h, cb[5], L.PAD_REFLECT, 1, False), cb[6]) + x
            if last:
                for a in adds:
                    y = y + a
            x = r16(y)
    return x


def run_head(mods, x, scale=1.0):
    (u,) = NW._units(mods)
    _, conv, pmode, pad, norm, act, slope = u
    return _act(_conv(x, conv, pmode, pad, True), act, slope) * scale
