"""Generate tests/golden/stage1_warp.npz by running the UNMODIFIED reference (training/warping_loss.py:calc_warping_loss,
training/volumetric_rendering/ray_sampler.py:RaySampler) on CPU in the build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_goldens_stage1.py          (needs /root/reference; the fixtures it writes are committed)

Shims (probe-only, nothing of the reference is copied): Tensor.cuda = identity (ray_sampler.py:38, warping_loss.py:14-15);
G is a stub whose synthesis() returns a fixed canonical image; the feature network is a small seeded conv stack with 22
children standing in for torchvision VGG16.features (weights are not available offline) -- rebuilt from the same seed by
the tests (tests/golden_util.py:stage1_feature_net).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('EG3D_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

torch.Tensor.cuda = lambda self, *a, **k: self

from training.warping_loss import calc_warping_loss  # noqa: E402
from training.volumetric_rendering.ray_sampler import RaySampler  # noqa: E402

import stage1_oracle as s1  # noqa: E402
import synth_params as sp  # noqa: E402
from golden_util import stage1_feature_net, stage1_inputs  # noqa: E402


class StubG:
    def __init__(self, img):
        self.img = img

    def synthesis(self, ws, c, **kw):
        return {'image': self.img}


def literal_noise_reg(noise_bufs):
    """Literal transcription target for the regulariser statement block (w_projector.py:221-237) -- used only to check the
    oracle restatement; the loop below follows the reference line by line."""
    import torch.nn.functional as F
    reg_loss = 0.0
    for v in noise_bufs.values():
        noise = v[None, None, :, :]
        while True:
            reg_loss += (noise * torch.roll(noise, shifts=1, dims=3)).mean() ** 2
            reg_loss += (noise * torch.roll(noise, shifts=1, dims=2)).mean() ** 2
            if noise.shape[2] <= 8:
                break
            noise = F.avg_pool2d(noise, kernel_size=2)
    return reg_loss


def main():
    out = {}
    for name, R, H in (('a', 32, 64), ('b', 128, 512)):
        inp = stage1_inputs(name, R, H)
        vgg = stage1_feature_net()
        ext = inp['extrinsic'].clone().requires_grad_(True)
        depth = inp['depth'].clone().requires_grad_(True)
        can = inp['can_image'].clone().requires_grad_(True)
        loss, warped = calc_warping_loss(torch.zeros(1, 14, 512), torch.zeros(1, 25), ext, inp['init_ext'], inp['intrinsic'], depth,
                                         inp['target'], StubG(can), vgg, RaySampler(), layers='14')
        loss.backward()
        # oracle restatement on the same inputs
        ext2 = inp['extrinsic'].clone().requires_grad_(True)
        depth2 = inp['depth'].clone().requires_grad_(True)
        loss2, warped2 = s1.warping_loss(inp['can_image'], ext2, inp['init_ext'], inp['intrinsic'], depth2, inp['target'], vgg, '14')
        loss2.backward()
        print(name, 'loss', loss.item(), 'oracle', loss2.item(), 'warped max diff', (warped - warped2).abs().max().item(),
              'd_ext rel', ((ext.grad - ext2.grad).norm() / ext.grad.norm()).item(),
              'd_depth rel', ((depth.grad - depth2.grad).norm() / depth.grad.norm()).item())
        uv, _ = s1.warp_uv(inp['extrinsic'], inp['init_ext'], inp['intrinsic'], inp['depth'])
        out[f'{name}_loss'] = np.float64(loss.item())
        out[f'{name}_warped_sub'] = warped.detach()[:, :, ::4, ::4].numpy().astype(np.float32)
        out[f'{name}_d_ext'] = ext.grad.numpy().astype(np.float32)
        out[f'{name}_d_depth_sub'] = depth.grad[:, :, ::4, ::4].numpy().astype(np.float32)
        out[f'{name}_uv_sub'] = uv.detach().reshape(R, R, 2)[::4, ::4].numpy().astype(np.float32)
    # noise regulariser: oracle vs the literal statement block
    g = torch.Generator().manual_seed(5)
    bufs = {f'n{i}': torch.randn(r, r, generator=g) for i, r in enumerate([4, 8, 8, 16, 16, 32, 64, 128, 256])}
    a, b = literal_noise_reg(bufs), s1.noise_regularizer(list(bufs.values()))
    print('noise reg', float(a), float(b))
    out['noise_reg'] = np.float64(float(a))
    # create_samples: run the reference function's own source (single_id_coach.py imports lpips / mrcfile, absent here)
    import ast
    src = open(os.path.join(REF, 'training', 'coaches', 'single_id_coach.py')).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'create_samples'][0]
    ns = {'np': np, 'torch': torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'single_id_coach.py', 'exec'), ns)
    for N in (16, 256):
        ref_s, _, _ = ns['create_samples'](N=N, voxel_origin=[0, 0, 0], cube_length=1.0)
        mine, _, _ = s1.create_samples(N=N, voxel_origin=[0, 0, 0], cube_length=1.0)
        print('create_samples', N, 'max diff vs oracle', (ref_s - mine).abs().max().item())
        out[f'samples_n{N}'] = ref_s[0, ::(1 if N == 16 else 4099)].numpy().astype(np.float32)
    # pose helpers: the reference's own rot6d_to_rotmat (utils/camera_utils.py:259-273) and LinePlaneCollision
    # (training/warping_loss.py:58-72) on seeded inputs; the extrinsic assembly of w_projector.py:160-171 is inline code,
    # transcribed literally below (the .cuda() calls are identities under the shim)
    from utils.camera_utils import rot6d_to_rotmat
    from training.warping_loss import LinePlaneCollision
    g = torch.Generator().manual_seed(8)
    x6 = torch.randn(2, 6, generator=g)          # not 3 rows: the reference calls torch.cross without dim (first size-3 dimension)
    out['rot6d_in'] = x6.numpy()
    out['rot6d_out'] = rot6d_to_rotmat(x6.clone()).numpy()
    n_, pp_, rd_, rp_ = [torch.randn(50, 3, generator=g) for _ in range(4)]
    out['lpc_in'] = torch.stack([n_, pp_, rd_, rp_]).numpy()
    out['lpc_out'] = LinePlaneCollision(n_, pp_, rd_, rp_).numpy()
    pred_rotmat = rot6d_to_rotmat(x6[:1].clone())
    translation_opt = torch.tensor([[0.02, -0.01, 0.03]])
    radius = 2.7
    pred_ext_tmp = torch.eye(4).unsqueeze(0).repeat(pred_rotmat.shape[0], 1, 1).cuda()
    pred_translation = -radius * pred_rotmat[:, :3, 2]
    pred_ext_tmp[:, :3, :3] = pred_rotmat
    translation_opt_world = -torch.bmm(pred_ext_tmp[:, :3, :3], translation_opt.unsqueeze(-1)) * 2.7
    tmp_translation = translation_opt_world.squeeze(-1) + pred_translation
    tmp_translation = tmp_translation / torch.norm(tmp_translation, dim=-1) * 2.7
    pred_ext = torch.eye(4).unsqueeze(0).cuda()
    pred_ext[:, :3, 3] = tmp_translation
    pred_ext[:, :3, :3] = pred_ext_tmp[:, :3, :3]
    out['ext_translation'] = translation_opt.numpy()
    out['ext_out'] = pred_ext.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'stage1_warp.npz'), **out)
    print('wrote stage1_warp.npz')


if __name__ == '__main__':
    main()
