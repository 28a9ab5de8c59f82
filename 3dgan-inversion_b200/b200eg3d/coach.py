"""The two per-step inversion loops of the reference, assembled on the b200eg3d operators.

    PTIStep            one pivotal-tuning step: single_id_coach.py:64-77 (synthesis -> calc_loss -> backward -> Adam on all
                       generator parameters, base_coach.py:96-126); LPIPS needs pretrained weights and is the caller's extra term
    ProjectionStep     one w-projection iteration: w_projector.py:160-268 (two synthesis calls, warping loss, feature distance,
                       noise regulariser, three Adam optimisers, noise normalisation) with the reference's learning-rate /
                       w-noise schedule (:173-181)

Both can run eagerly (every kernel launched from Python) or as a replayed CUDA graph (graphs.GraphedStep): the kernels and
their order are identical, which tests/test_gpu_graphed.py checks step by step.
"""
import math

import numpy as np
import torch

from . import losses, optim, projector
from .graphs import GraphedStep


class PTIStep:
    """Pivotal tuning of one image.  step(ws, c, real) -> loss (a device scalar; no host sync)."""

    def __init__(self, G, lr=3e-4, graphed=True, example=None, pt_l2_lambda=1.0, depth_tv_lambda=1.0, extra_loss=None,
                 noise_mode='const', force_fp32=True):
        self.G = G
        self.params = [p for n, p in G.named_parameters() if '.mapping.' not in n]     # base_coach.py:96-99 tunes G; mapping is unused
        for p in self.params:
            p.requires_grad_(True)
        self.opt = optim.Adam(self.params, lr=lr)             # one launch per step, device-side step counter (graph-capturable)
        self.l2, self.tv, self.extra = pt_l2_lambda, depth_tv_lambda, extra_loss
        self.noise_mode, self.force_fp32 = noise_mode, force_fp32
        self.graph = None
        if graphed:
            if example is None:
                raise ValueError('graphed=True needs example inputs (ws, c, real) to size the static buffers')
            self.graph = GraphedStep(self._eager, list(example), optimizer=self.opt, warmup=3)

    def _eager(self, ws, c, real):
        out = self.G.synthesis(ws, c, noise_mode=self.noise_mode, force_fp32=self.force_fp32)
        loss = losses.pti_loss(out, real, pt_l2_lambda=self.l2, depth_tv_lambda=self.tv)
        if self.extra is not None:                      # e.g. the caller's LPIPS term (base_coach.py:111-116)
            loss = loss + self.extra(out, real)
        if self.graph is None:
            self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return loss

    def step(self, ws, c, real):
        return self.graph(ws, c, real) if self.graph is not None else self._eager(ws, c, real)

    __call__ = step


def projection_schedule(step, num_steps, w_std, cam_preheat_steps=0, initial_learning_rate=0.01, lr_rampdown_length=0.25,
                        lr_rampup_length=0.05, initial_noise_factor=0.05, noise_ramp_length=0.75):
    """(lr, w_noise_scale) of iteration `step`: w_projector.py:173-179."""
    t = (step - cam_preheat_steps) / (num_steps - cam_preheat_steps)
    w_noise_scale = w_std * initial_noise_factor * max(0.0, 1.0 - t / noise_ramp_length) ** 2
    lr_ramp = min(1.0, (1.0 - t) / lr_rampdown_length)
    lr_ramp = 0.5 - 0.5 * np.cos(lr_ramp * np.pi)
    lr_ramp = lr_ramp * min(1.0, t / lr_rampup_length)
    return initial_learning_rate * lr_ramp, w_noise_scale


class ProjectionStep:
    """State and iteration body of the w-projection (w_projector.py:112-268) around a pose predictor.

    pose_params: tensors the camera optimiser updates (the reference: cam_predictor.parameters()); pose_fn() -> [1,3,3]
    rotation (rot6d_to_rotmat(cam_predictor(target)), euler2rot(...), ...).  feature_fn / torch_vgg are the caller's
    feature networks (VGG16-LPIPS / torchvision VGG16.features in the reference; both need pretrained weights).
    The learning rate of the latent/noise optimiser is a device scalar so that a captured graph follows the schedule.
    """

    def __init__(self, G, w_start, pose_params, pose_fn, init_ext, intrinsic, target_images, feature_fn, torch_vgg,
                 first_inv_lr=5e-3, cam_lr=1e-4, translation_lr=1e-4, regularize_noise_weight=1e5, graphed=True, seed=0):
        dev = w_start.device
        self.G = G.requires_grad_(False)
        self.noise_bufs = {n: b for n, b in G.backbone.synthesis.named_buffers() if 'noise_const' in n}
        self.noise_bufs2 = {n: b for n, b in G.superresolution.named_buffers() if 'noise_const' in n}
        gen = torch.Generator().manual_seed(seed)
        for b in list(self.noise_bufs.values()) + list(self.noise_bufs2.values()):      # w_projector.py:128-133
            b.copy_(torch.randn(b.shape, generator=gen))
            b.requires_grad = True
        self.w_opt = w_start.detach().clone().to(torch.float32).requires_grad_(True)
        self.translation_opt = torch.zeros(1, 3, device=dev, requires_grad=True)
        self.pose_fn = pose_fn
        self.lr = torch.tensor(first_inv_lr, device=dev, dtype=torch.float32)
        self.optimizer = optim.Adam([self.w_opt] + list(self.noise_bufs.values()) + list(self.noise_bufs2.values()),
                                    betas=(0.9, 0.999), lr=self.lr)
        self.cam_optimizer = optim.Adam(list(pose_params), lr=cam_lr, betas=(0.9, 0.999))
        self.translation_optimizer = optim.Adam([self.translation_opt], lr=translation_lr)
        self.init_ext, self.intrinsic = init_ext, intrinsic
        self.w2c = torch.linalg.inv(init_ext.reshape(4, 4)).contiguous()              # constant; linalg.inv cannot be graph-captured
        self.target_images = target_images
        t255 = (target_images + 1) / 2 * 255
        if t255.shape[2] > 256:
            t255 = torch.nn.functional.interpolate(t255, size=(256, 256), mode='area')
        with torch.no_grad():
            self.target_features = feature_fn(t255)
        self.feature_fn, self.torch_vgg, self.reg_w = feature_fn, torch_vgg, regularize_noise_weight
        self.parts = None
        self.graph = None              # _eager runs during the warm-up / capture below
        self.graph = GraphedStep(self._eager, [torch.zeros_like(self.w_opt)], optimizer=self, warmup=3) if graphed else None

    def zero_grad(self, set_to_none=True):
        for o in (self.optimizer, self.cam_optimizer, self.translation_optimizer):
            o.zero_grad(set_to_none=set_to_none)

    def _eager(self, w_noise):
        loss, parts = projector.projection_step_loss(self.G, self.w_opt, self.pose_fn(), self.translation_opt, self.init_ext,
                                                     self.intrinsic, self.target_images, self.target_features, self.feature_fn,
                                                     self.torch_vgg, self.noise_bufs, self.noise_bufs2, w_noise=w_noise,
                                                     regularize_noise_weight=self.reg_w, w2c=self.w2c)
        if self.graph is None:
            self.zero_grad()
        loss.backward()
        self.cam_optimizer.step()                       # w_projector.py:255-257 (order as in the reference)
        self.optimizer.step()
        self.translation_optimizer.step()
        projector.normalize_noise_(list(self.noise_bufs.values()) + list(self.noise_bufs2.values()))
        self.parts = parts
        return loss

    def step(self, w_noise, lr=None):
        if lr is not None:
            self.lr.fill_(float(lr))
        return self.graph(w_noise) if self.graph is not None else self._eager(w_noise)

    __call__ = step


def project(G, target, *, w_start, w_std, pose_params, pose_fn, feature_fn, torch_vgg, num_steps=400, device=None, graphed=True,
            seed=0, **schedule):
    """w_projector.project (w_projector.py:28-275) for an already-encoded start latent: runs `num_steps` iterations with the
    reference's schedule and returns (ws [1, num_ws, 512], cam [1, 25]) as the reference does (:270-275).
    target: [3, H, W] in [-1, 1]."""
    device = device or w_start.device
    init_ext = torch.tensor([1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 2.7, 0, 0, 0, 1], dtype=torch.float32, device=device).reshape(1, 4, 4)
    intrinsic = torch.tensor([4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1], dtype=torch.float32, device=device)
    st = ProjectionStep(G, w_start, pose_params, pose_fn, init_ext, intrinsic, target.unsqueeze(0).to(device).float().contiguous(),
                        feature_fn, torch_vgg, graphed=graphed, seed=seed)
    gen = torch.Generator(device='cpu').manual_seed(seed + 1)
    for it in range(num_steps):
        lr, scale = projection_schedule(it, num_steps, w_std, **schedule)
        w_noise = (torch.randn(st.w_opt.shape, generator=gen) * scale).to(device)
        st.step(w_noise, lr=lr)
    with torch.no_grad():
        ext = projector.assemble_extrinsic(pose_fn(), st.translation_opt)
        cam = torch.cat([ext.reshape(-1, 16), intrinsic.reshape(1, 9)], dim=-1)
        ws = (st.w_opt + w_noise).repeat(1, G.backbone.num_ws, 1)
    return ws.detach(), cam.detach()
