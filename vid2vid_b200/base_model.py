"""Host-side helpers of models/base_model.py that train.py's outer loop calls between steps (train.py:105-128): checkpoint
files and the learning-rate / frame-budget schedules.  No tensor math; shared by Vid2VidModelG and Vid2VidModelD."""
import math
import os

import torch


class HostScheduleMixin:
    """Expects self.opt (checkpoints_dir, name, lr, niter, niter_decay, beta1, max_frames_*), self.old_lr and, for the
    generator, self.n_scales / self.n_frames_bp / self.n_frames_load / self.n_frames_per_gpu."""

    @property
    def save_dir(self):
        return os.path.join(self.opt.checkpoints_dir, self.opt.name)                 # base_model.py:15

    def save_network(self, network, network_label, epoch_label, gpu_ids=None):
        """base_model.py:43-48: <save_dir>/<epoch>_net_<label>.pth holding the CPU state_dict.  The reference moves the
        module to the CPU and back; here the tensors are copied instead, so the parameters keep their device addresses and the
        plans built on them stay valid (same file contents)."""
        os.makedirs(self.save_dir, exist_ok=True)
        path = os.path.join(self.save_dir, '%s_net_%s.pth' % (epoch_label, network_label))
        torch.save({k: v.detach().to('cpu', copy=True) for k, v in network.state_dict().items()}, path)
        return path

    def update_learning_rate(self, epoch, model):
        """base_model.py:154-159: linear decay to zero over niter_decay epochs after the first niter."""
        lr = self.opt.lr * (1 - (epoch - self.opt.niter) / self.opt.niter_decay)
        for group in getattr(self, 'optimizer_' + model).param_groups:
            group['lr'] = lr
        print('update learning rate: %f -> %f' % (self.old_lr, lr))
        self.old_lr = lr

    def update_fixed_params(self):
        """base_model.py:161-167: from epoch niter_fix_global on every scale is trained, with a fresh optimizer over all of them."""
        params = []
        for s in range(self.n_scales):
            params += list(getattr(self, 'netG' + str(s)).parameters())
        self.optimizer_G = self._make_adam(params, lr=self.old_lr, betas=(self.opt.beta1, 0.999))
        self.finetune_all = True
        print('------------ Now finetuning all scales -----------')

    def update_training_batch(self, ratio):
        """base_model.py:169-181: every niter_step epochs more frames are back-propagated through and more are loaded per GPU
        (one process per GPU here: n_gpus = 1 in the reference's n_frames_load = n_gpus * n_frames_per_gpu)."""
        nfb, nfl = self.n_frames_bp, self.n_frames_load
        if nfb < nfl:
            nfb = min(self.opt.max_frames_backpropagate, 2 ** ratio)
            self.n_frames_bp = nfl // int(math.ceil(float(nfl) / nfb))
            print('-------- Updating number of backpropagated frames to %d ----------' % self.n_frames_bp)
        if self.n_frames_per_gpu < self.opt.max_frames_per_gpu:
            self.n_frames_per_gpu = min(self.n_frames_per_gpu * 2, self.opt.max_frames_per_gpu)
            self.n_frames_load = getattr(self, 'n_gpus', 1) * self.n_frames_per_gpu
            print('-------- Updating number of frames per gpu to %d ----------' % self.n_frames_per_gpu)

    @staticmethod
    def _make_adam(params, **kw):
        params = list(params)
        if params and all(p.is_cuda for p in params):
            try:
                return torch.optim.Adam(params, fused=True, **kw)
            except (TypeError, RuntimeError):
                pass
        return torch.optim.Adam(params, **kw)
