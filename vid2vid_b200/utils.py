"""Small host-side helpers shared by the modules, the tests and bench.py.

Nothing here touches the GPU.  `make_opt` mirrors the argparse namespace the
reference builds in options/base_options.py:11-128 + options/test_options.py:4-15
(only the fields the generator / discriminator hot path reads).
"""
import zlib
from types import SimpleNamespace

import torch


def make_opt(**kw):
    """Option namespace with the reference's defaults (options/base_options.py:14-84)."""
    opt = SimpleNamespace(
        # base_options.py:14-20
        dataroot='datasets/Cityscapes/', batchSize=1, loadSize=512, fineSize=512,
        input_nc=3, label_nc=0, output_nc=3,
        # base_options.py:22-26
        netG='composite', ngf=128, ndf=64, n_blocks=9, n_downsample_G=3,
        gpu_ids=[0], n_gpus_gen=1, name='synthetic', dataset_mode='temporal',
        checkpoints_dir='./checkpoints', norm='batch',
        # base_options.py:43-59
        use_instance=False, label_feat=False, feat_num=3, n_blocks_local=3,
        n_local_enhancers=1, n_frames_G=3, n_scales_spatial=1, no_first_img=False,
        use_single_G=False, fg=False, fg_labels=[26], no_flow=False,
        openpose_only=False, densepose_only=False, add_face_disc=False,
        load_pretrain='', debug=True, fp16=False,
        # test_options.py
        which_epoch='latest', use_real_img=False, isTrain=False,
        # train_options.py (subset read by Vid2VidModelD)
        num_D=2, n_layers_D=3, n_frames_D=3, n_scales_temporal=3, no_ganFeat=False,
        no_vgg=True, lambda_feat=10.0, lambda_F=10.0, lambda_T=10.0,
        gan_mode='ls', continue_train=False, niter_fix_global=0, lr=0.0002,
        beta1=0.5, TTUR=False, max_frames_per_gpu=1, n_frames_total=30,
        max_frames_backpropagate=1,
        # not a reference option: random-initialised networks (no checkpoints exist offline); False = the reference's behaviour
        synthetic_weights=True,
    )
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt


def det_fill_(module, seed=0):
    """Fill every parameter/buffer of `module` with values that depend only on
    (seed, tensor name, shape) -- not on construction order or RNG history -- so the
    reference modules (oracle side) and ours can be given identical random weights.

    Distributions follow the reference's initialiser (models/networks.py:15-21):
    conv weights N(0, 0.02), norm gamma N(1, 0.02); biases get small non-zero values
    so that affine shifts are exercised.
    """
    sd = module.state_dict()
    with torch.no_grad():
        for name, t in sd.items():
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            leaf = name.rsplit('.', 1)[-1]
            if leaf == 'num_batches_tracked':
                t.zero_()
            elif leaf == 'running_mean':
                t.zero_()
            elif leaf == 'running_var':
                t.fill_(1.0)
            elif t.dim() >= 2:
                t.copy_(torch.randn(t.shape, generator=g) * 0.02)
            elif leaf == 'weight':          # norm gamma
                t.copy_(1.0 + torch.randn(t.shape, generator=g) * 0.02)
            else:                           # conv bias / norm beta
                t.copy_(torch.randn(t.shape, generator=g) * 0.05)
    return module


def synth_label_sequence(n_frames, H, W, label_nc=35, block=8, seed=0, drift=1):
    """Blocky synthetic label maps: ids U{0..label_nc-1} on a coarse grid,
    nearest-upsampled by `block`, shifted by `drift` coarse cells per frame so consecutive
    frames differ.  Returns float tensor (1, n_frames, 1, H, W) of label ids (the
    reference feeds label ids as floats, data/test_dataset.py / vid2vid_model_G.py:95)."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = (H + block - 1) // block, (W + block - 1) // block
    base = torch.randint(0, label_nc, (gh, gw + drift * n_frames), generator=g)
    frames = []
    for t in range(n_frames):
        m = base[:, t * drift: t * drift + gw]
        m = m.repeat_interleave(block, 0).repeat_interleave(block, 1)[:H, :W]
        frames.append(m)
    return torch.stack(frames).float().view(1, n_frames, 1, H, W)
