"""TEST INFRASTRUCTURE -- not product code.

Imports the *unmodified* reference (NVIDIA/vid2vid at /root/reference) on CPU under the
installed PyTorch with the harness-level shims of SURVEY.md 8(c).  Only usable in the build
container (the GPU box has no /root/reference); used by oracle/make_golden.py to produce the
committed fixtures in tests/golden/ and by the `not gpu` tests that pin oracle/*.py against
the reference itself when it is present.

Shims (none changes reference arithmetic):
  1. Tensor.cuda / Module.cuda -> identity  (reference hard-codes .cuda(gpu_id),
     models/networks.py:93,113,219,228; models/vid2vid_model_G.py:90,100,110)
  2. torch.cuda.FloatTensor/ByteTensor -> CPU tensor types (vid2vid_model_G.py:94,
     base_model.py:14,147)
  3. fractions.gcd = math.gcd (models/models.py:7-8)
  4. Tensor.get_device() on CPU returns -1 already; `.cuda(-1)` is swallowed by shim 1.
"""
import os
import sys
import math
import fractions

REF_ROOT = os.environ.get('V2V_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'models'))


_installed = False


def install():
    """Apply the shims and put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError('reference tree not found at %s' % REF_ROOT)
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.ByteTensor = torch.ByteTensor
    if not hasattr(fractions, 'gcd'):
        fractions.gcd = math.gcd
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def networks():
    install()
    import models.networks as ref_networks   # noqa  (reference module)
    return ref_networks


def model_G_class():
    install()
    from models.vid2vid_model_G import Vid2VidModelG
    return Vid2VidModelG


def make_model_G(opt, single_G=None):
    """Build the reference Vid2VidModelG without touching disk: load_network and
    load_single_G are stubbed (SURVEY 8c shim 4).  `single_G` is the module returned by
    the stubbed load_single_G (the caller builds it with networks().define_G)."""
    cls = model_G_class()
    import contextlib
    import io

    class _M(cls):
        def load_network(self, *a, **k):
            return None

        def load_single_G(self):
            return single_G

    m = _M()
    with contextlib.redirect_stdout(io.StringIO()):
        m.initialize(opt)
    return m
