"""Generate tests/golden/stage1_iter.npz: ONE iteration of the reference's w-projection loop, run from its own source.
TEST INFRASTRUCTURE ONLY (build container; needs /root/reference).

    python oracle/make_goldens_stage1_iter.py

training/projectors/w_projector.py cannot be imported or run as a whole here (it downloads VGG16, needs pretrained encoders and a
GPU).  Its per-iteration statements (the body of `for step in tqdm(range(num_steps))`, w_projector.py:160-268) are therefore taken
from the file with `ast` and executed unmodified in a namespace this script prepares: the unmodified reference TriPlaneGenerator
(tiny architecture, CPU), the reference's own calc_warping_loss / rot6d_to_rotmat / RaySampler, seeded stand-ins for the three
pretrained networks (tests/golden_util.py) and the loop state the reference sets up in :60-135.  Shims: Tensor.cuda = identity.
Recorded: the loss and its parts, and every optimised quantity after the iteration's three Adam steps and the noise normalisation.
"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('EG3D_REFERENCE', '/root/reference')
for p in (REF, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
torch.Tensor.cuda = lambda self, *a, **k: self

import math  # noqa: E402
import synth_params as sp  # noqa: E402
import eg3d_oracle as oracle  # noqa: E402
from golden_util import stage1_feature_fn, stage1_feature_net, stage1_iter_setup, stage1_noise_init, stage1_pose_net  # noqa: E402
from training.triplane import TriPlaneGenerator  # noqa: E402
from training.warping_loss import calc_warping_loss  # noqa: E402
from training.volumetric_rendering.ray_sampler import RaySampler  # noqa: E402
from utils.camera_utils import rot6d_to_rotmat, compute_rotation_matrix_from_quaternion, euler2rot  # noqa: E402
from configs import hyperparameters  # noqa: E402


def loop_body():
    src = open(os.path.join(REF, 'training', 'projectors', 'w_projector.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'project'][0]
    loop = [n for n in fn.body if isinstance(n, ast.For)][-1]
    return compile(ast.Module(body=loop.body, type_ignores=[]), 'w_projector.py', 'exec')


def main():
    cfg = stage1_iter_setup()
    R, S = cfg['R'], cfg['S']
    G = TriPlaneGenerator(rendering_kwargs=cfg['rk'], **cfg['gk']).eval().requires_grad_(False).float()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), cfg['param_seed'])
    G.neural_rendering_resolution = R
    for m in G.modules():
        if hasattr(m, 'noise_strength'):
            m.noise_strength.data.fill_(0.05)
    feat, torch_vgg, cam_predictor = stage1_feature_net(11), stage1_feature_net(12), stage1_pose_net(13)
    cam_predictor.requires_grad_(True)
    target = cfg['target']
    device = torch.device('cpu')
    # ---- loop state, as the reference prepares it before the loop (w_projector.py:75-135)
    init_ext = torch.Tensor([1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 2.7, 0, 0, 0, 1]).reshape(-1, 4, 4)
    intrinsic = torch.tensor([4.2647, 0, 0.5, 0, 4.2647, 0.5, 0, 0, 1]).unsqueeze(0)
    canonical_cam = torch.cat([init_ext.reshape(-1, 16), intrinsic], dim=-1)
    noise_bufs = {name: buf for (name, buf) in G.backbone.synthesis.named_buffers() if 'noise_const' in name}
    noise_bufs2 = {name: buf for (name, buf) in G.superresolution.named_buffers() if 'noise_const' in name}
    target_images_contiguous = target.contiguous()
    target_images = (((target + 1) / 2) * 255).unsqueeze(0).to(torch.float32)
    target_images = F.interpolate(target_images, size=(256, 256), mode='area')
    vgg16 = stage1_feature_fn(feat)
    with torch.no_grad():
        target_features = vgg16(target_images, resize_images=False, return_lpips=True)
    w_opt = cfg['w_start'].clone().requires_grad_(True)
    translation_opt = cfg['translation'].clone().requires_grad_(True)
    optimizer = torch.optim.Adam([w_opt] + list(noise_bufs.values()) + list(noise_bufs2.values()), betas=(0.9, 0.999), lr=hyperparameters.first_inv_lr)
    cam_optimizer = torch.optim.Adam(cam_predictor.parameters(), lr=hyperparameters.cam_lr_6d, betas=(0.9, 0.999))
    translation_optimizer = torch.optim.Adam([translation_opt], lr=hyperparameters.translation_lr)
    stage1_noise_init(list(noise_bufs.items()) + list(noise_bufs2.items()))
    for buf in list(noise_bufs.values()) + list(noise_bufs2.values()):
        buf.requires_grad = True
    # ---- RNG: the body draws randn_like(w_opt) (replaced by the fixture's w_noise) and, inside the two synthesis calls, the depth noise
    w_noise_fixed = cfg['w_noise']
    real_randn_like = torch.randn_like
    torch.randn_like = lambda t, **k: (w_noise_fixed / ns['w_noise_scale']).to(t.dtype) if t is w_opt else real_randn_like(t, **k)
    draws = []
    real_rand, real_rand_like = torch.rand, torch.rand_like
    dgen = torch.Generator().manual_seed(23)

    def rand(*size, **k):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        t = real_rand(shape, generator=dgen)
        draws.append(t.clone())
        return t
    torch.rand = rand
    torch.rand_like = lambda t, **k: rand(t.shape)
    ns = dict(torch=torch, np=np, F=F, math=math, os=os, G=G, cam_predictor=cam_predictor, vgg16=vgg16, torch_vgg=torch_vgg, layers='14',
              ray_generator=RaySampler(), calc_warping_loss=calc_warping_loss, rot6d_to_rotmat=rot6d_to_rotmat, euler2rot=euler2rot,
              compute_rotation_matrix_from_quaternion=compute_rotation_matrix_from_quaternion, hyperparameters=hyperparameters,
              global_config=types.SimpleNamespace(use_quaternions=False, use_6d=True, visualize_opt_process=False, visualize_warp_process=False),
              radius=2.7, init_ext=init_ext, intrinsic=intrinsic, canonical_cam=canonical_cam, noise_bufs=noise_bufs, noise_bufs2=noise_bufs2,
              target_images=target_images, target_images_contiguous=target_images_contiguous, target_features=target_features, w_opt=w_opt,
              translation_opt=translation_opt, optimizer=optimizer, cam_optimizer=cam_optimizer, translation_optimizer=translation_optimizer,
              num_steps=cfg['num_steps'], w_std=cfg['w_std'], initial_noise_factor=0.05, noise_ramp_length=0.75, lr_rampdown_length=0.25,
              lr_rampup_length=0.05, initial_learning_rate=0.01, regularize_noise_weight=1e5, step=cfg['step'], outdir='/tmp', w_name='x',
              device=device, PIL=None)
    exec(loop_body(), ns)
    torch.rand, torch.rand_like, torch.randn_like = real_rand, real_rand_like, real_randn_like
    assert len(draws) == 4, [tuple(d.shape) for d in draws]           # (rand_like, rand) of the predicted render, then of the canonical render
    out = {'loss': np.float64(ns['loss'].item()), 'dist': np.float64(ns['dist'].item()), 'warp_loss': np.float64(ns['warp_loss'].item()),
           'reg_loss': np.float64(float(ns['reg_loss'])), 'lr': np.float64(ns['lr']), 'w_noise_scale': np.float64(ns['w_noise_scale']),
           'w_opt_after': w_opt.detach().numpy(), 'translation_after': translation_opt.detach().numpy(),
           'pose_w_after': cam_predictor[2].weight.detach().numpy(), 'pose_b_after': cam_predictor[2].bias.detach().numpy(),
           'grad_w_opt': w_opt.grad.numpy(), 'grad_translation': translation_opt.grad.numpy(), 'grad_pose_b': cam_predictor[2].bias.grad.numpy(),
           'pred_cam': ns['pred_cam'].detach().numpy(), 'image_sub8': ns['pred_dict']['image'].detach()[..., ::8, ::8].numpy(),
           'depth': ns['pred_depths'].detach().numpy()}
    for i, d in enumerate(draws):
        out[f'draw{i}'] = d.numpy()
    for name, buf in list(noise_bufs.items()) + list(noise_bufs2.items()):
        key = name.replace('.', '_')
        st = max(1, buf.shape[0] // 32)
        out['noise_after_' + key] = buf.detach()[::st, ::st].numpy().copy()          # <= 32 x 32 sub-sample + moments of the whole buffer
        out['noise_after_mom_' + key] = np.array([buf.detach().double().sum().item(), buf.detach().double().square().sum().item()])
        out['noise_grad_norm_' + key] = np.float64(buf.grad.norm().item())
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'stage1_iter.npz'), **out)
    print('loss', out['loss'], 'dist', out['dist'], 'warp', out['warp_loss'], 'reg', out['reg_loss'], 'lr', out['lr'])
    print('wrote stage1_iter.npz', os.path.getsize(os.path.join(ROOT, 'tests', 'golden', 'stage1_iter.npz')) // 1024, 'kB')


if __name__ == '__main__':
    main()
