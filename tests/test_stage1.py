"""Stage-1 (w-projection) caller-side pieces: the oracle against the fixtures generated from the reference's own
calc_warping_loss (CPU), and the CUDA kernels against the oracle / fixtures (GPU)."""
import os

import numpy as np
import pytest
import torch

import stage1_oracle as s1
from golden_util import stage1_feature_net, stage1_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'stage1_warp.npz')
CASES = {'a': (32, 64), 'b': (128, 512)}


def relerr(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('name', ['a', 'b'])
def test_oracle_matches_reference_fixture(name):
    fx = np.load(GOLD)
    R, H = CASES[name]
    inp = stage1_inputs(name, R, H)
    ext = inp['extrinsic'].clone().requires_grad_(True)
    depth = inp['depth'].clone().requires_grad_(True)
    loss, warped = s1.warping_loss(inp['can_image'], ext, inp['init_ext'], inp['intrinsic'], depth, inp['target'], stage1_feature_net(), '14')
    loss.backward()
    assert abs(loss.item() - float(fx[f'{name}_loss'])) <= 2e-6 * abs(float(fx[f'{name}_loss']))
    assert np.abs(warped.detach()[:, :, ::4, ::4].numpy() - fx[f'{name}_warped_sub']).max() < 1e-4
    assert relerr(ext.grad, torch.from_numpy(fx[f'{name}_d_ext'])) < 1e-4
    assert relerr(depth.grad[:, :, ::4, ::4], torch.from_numpy(fx[f'{name}_d_depth_sub'])) < 1e-4
    uv, _ = s1.warp_uv(inp['extrinsic'], inp['init_ext'], inp['intrinsic'], inp['depth'])
    assert np.abs(uv.reshape(R, R, 2)[::4, ::4].numpy() - fx[f'{name}_uv_sub']).max() < 1e-5


def test_oracle_noise_regulariser_fixture():
    g = torch.Generator().manual_seed(5)
    bufs = [torch.randn(r, r, generator=g) for r in [4, 8, 8, 16, 16, 32, 64, 128, 256]]
    assert abs(float(s1.noise_regularizer(bufs)) - float(np.load(GOLD)['noise_reg'])) < 1e-6


@pytest.fixture(scope='module')
def b2():
    import b200eg3d
    assert torch.cuda.is_available()
    b200eg3d.ops.library_info()
    return b200eg3d


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['a', 'b'])
def test_warp_uv_fwd_bwd(b2, name):
    R, H = CASES[name]
    inp = stage1_inputs(name, R, H)
    g = torch.Generator().manual_seed(3)
    up = torch.randn(R * R, 2, generator=g)
    ext = inp['extrinsic'].clone().requires_grad_(True)
    depth = inp['depth'].clone().requires_grad_(True)
    uv_ref, _ = s1.warp_uv(ext, inp['init_ext'], inp['intrinsic'], depth)
    (uv_ref * up).sum().backward()
    ext_d = inp['extrinsic'].clone().cuda().requires_grad_(True)
    depth_d = inp['depth'].clone().cuda().requires_grad_(True)
    uv = b2.projector.warp_uv(ext_d, inp['init_ext'].cuda(), inp['intrinsic'].cuda(), depth_d)
    (uv * up.cuda()).sum().backward()
    assert (uv.cpu() - uv_ref).abs().max().item() < 2e-5            # fp32, uv in [-1, 1]
    assert relerr(depth_d.grad, depth.grad) < 1e-4
    assert relerr(ext_d.grad, ext.grad) < 1e-4
    assert ext_d.grad.shape == ext.grad.shape and float(ext_d.grad[0, 3].abs().max()) == 0.0


@pytest.mark.gpu
def test_warp_uv_reports_degenerate_intersection(b2):
    inp = stage1_inputs('a', 32, 64)
    # a depth map that puts the surface points in the plane through the canonical origin orthogonal to its normal: n . v == 0
    ext = inp['init_ext'].clone().cuda()
    depth = torch.zeros(1, 1, 32, 32, device='cuda')                # surface point == predicted camera origin == canonical origin
    with pytest.raises(RuntimeError, match='no intersection'):
        b2.projector.warp_uv(ext, inp['init_ext'].cuda(), inp['intrinsic'].cuda(), depth)
    uv = b2.projector.warp_uv(ext, inp['init_ext'].cuda(), inp['intrinsic'].cuda(), depth, check_intersection=False)
    assert uv.shape == (32 * 32, 2)


class _StubG:
    def __init__(self, img):
        self.img = img

    def synthesis(self, ws, c, **kw):
        return {'image': self.img}


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['a', 'b'])
def test_calc_warping_loss_against_reference_fixture(b2, name):
    """Same call as training/warping_loss.py:calc_warping_loss (stub G returning the fixture's canonical image)."""
    fx = np.load(GOLD)
    R, H = CASES[name]
    inp = {k: v.cuda() for k, v in stage1_inputs(name, R, H).items()}
    ext = inp['extrinsic'].clone().requires_grad_(True)
    depth = inp['depth'].clone().requires_grad_(True)
    vgg = stage1_feature_net().cuda()
    torch.backends.cudnn.allow_tf32 = False          # the stand-in feature net must run in fp32 like the CPU reference (|.| and ReLU kinks)
    loss, warped = b2.projector.calc_warping_loss(torch.zeros(1, 14, 512, device='cuda'), torch.zeros(1, 25, device='cuda'), ext,
                                                  inp['init_ext'], inp['intrinsic'], depth, inp['target'], _StubG(inp['can_image']), vgg,
                                                  b2.RaySampler(), layers='14')
    loss.backward()
    assert abs(loss.item() - float(fx[f'{name}_loss'])) <= 1e-3 * abs(float(fx[f'{name}_loss']))     # cuDNN convs of the stand-in feature net
    assert np.abs(warped.detach()[:, :, ::4, ::4].cpu().numpy() - fx[f'{name}_warped_sub']).max() < 1e-3
    assert relerr(ext.grad, torch.from_numpy(fx[f'{name}_d_ext'])) < 1e-2
    assert relerr(depth.grad[:, :, ::4, ::4], torch.from_numpy(fx[f'{name}_d_depth_sub'])) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize('sizes', [[4], [8, 16], [4, 8, 8, 16, 16, 32, 32, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512], [1024, 4]])
def test_noise_regulariser_fwd_bwd(b2, sizes):
    g = torch.Generator().manual_seed(len(sizes))
    bufs = [torch.randn(r, r, generator=g) for r in sizes]
    ref_in = [b.clone().double().requires_grad_(True) for b in bufs]
    ref = s1.noise_regularizer(ref_in)
    (ref * 1e5).backward()
    dev_in = [b.clone().cuda().requires_grad_(True) for b in bufs]
    reg = b2.projector.noise_regularizer(dev_in)
    (reg * 1e5).backward()
    assert abs(reg.item() - ref.item()) <= 1e-4 * abs(ref.item()) + 1e-9
    for a, b in zip(dev_in, ref_in):
        assert relerr(a.grad, b.grad) < 1e-3, a.shape


@pytest.mark.gpu
def test_normalize_noise_in_place(b2):
    g = torch.Generator().manual_seed(9)
    bufs = [torch.randn(r, r, generator=g) * (1 + i) + 0.3 * i for i, r in enumerate([4, 8, 64, 512])]
    ref = s1.normalize_noise(bufs)
    dev = {f'b{i}': b.clone().cuda() for i, b in enumerate(bufs)}
    b2.projector.normalize_noise_(dev)
    for a, b in zip(dev.values(), ref):
        assert (a.cpu() - b).abs().max().item() < 1e-4


def test_pose_helpers_match_reference_fixture():
    """Device-agnostic host helpers of projector.py against outputs of the reference's own functions (CPU)."""
    from b200eg3d import projector
    fx = np.load(GOLD)
    r = projector.rot6d_to_rotmat(torch.from_numpy(fx['rot6d_in']))
    assert np.abs(r.numpy() - fx['rot6d_out']).max() < 1e-6
    n_, pp_, rd_, rp_ = [torch.from_numpy(a) for a in fx['lpc_in']]
    assert np.abs(projector.LinePlaneCollision(n_, pp_, rd_, rp_).numpy() - fx['lpc_out']).max() < 1e-5
    ext = projector.assemble_extrinsic(projector.rot6d_to_rotmat(torch.from_numpy(fx['rot6d_in'][:1])), torch.from_numpy(fx['ext_translation']))
    assert ext.shape == (1, 4, 4) and np.abs(ext.numpy() - fx['ext_out']).max() < 1e-6
    with pytest.raises(RuntimeError, match='no intersection'):
        projector.LinePlaneCollision(torch.tensor([[0., 0., 1.]]), torch.zeros(1, 3), torch.tensor([[1., 0., 0.]]), torch.zeros(1, 3))
    net = stage1_feature_net()
    x = torch.rand(1, 3, 32, 32)
    assert torch.equal(projector.get_features(x, net, '14'), s1.get_features(x, net, '14'))
    with pytest.raises(ValueError):
        projector.get_features(x, net, '5')


@pytest.mark.gpu
def test_projection_iteration_matches_reference_loop_body(b2):
    """One whole w-projection iteration (w_projector.py:160-268) on the b200eg3d path against tests/golden/stage1_iter.npz, which
    oracle/make_goldens_stage1_iter.py recorded by EXECUTING the reference's own loop body (extracted with ast) around the unmodified
    reference generator / calc_warping_loss / rot6d_to_rotmat: loss and its three parts, the gradients of the latent, the pose
    predictor and the translation, and every optimised quantity after the three Adam steps and the noise normalisation."""
    import synth_params as sp
    from golden_util import stage1_feature_fn, stage1_iter_setup, stage1_noise_init, stage1_pose_net
    from b200eg3d import projector
    from b200eg3d.coach import ProjectionStep
    fx = np.load(os.path.join(os.path.dirname(GOLD), 'stage1_iter.npz'))
    cfg = stage1_iter_setup()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    G = b2.TriPlaneGenerator(rendering_kwargs=cfg['rk'], **cfg['gk']).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), cfg['param_seed'])
    G = G.cuda()
    G.neural_rendering_resolution = cfg['R']
    for m in G.modules():
        if hasattr(m, 'noise_strength'):
            m.noise_strength.data.fill_(0.05)
    feat, torch_vgg, pose = stage1_feature_net(11).cuda(), stage1_feature_net(12).cuda(), stage1_pose_net(13).cuda()
    target = cfg['target'].cuda()
    t255 = torch.nn.functional.interpolate(((target + 1) / 2 * 255).unsqueeze(0), size=(256, 256), mode='area')
    c0 = sp.camera(0.0, 0.0).cuda()
    st = ProjectionStep(G, cfg['w_start'].cuda(), list(pose.parameters()), lambda: projector.rot6d_to_rotmat(pose(t255)),
                        c0[:, :16].reshape(1, 4, 4).contiguous(), c0[0, 16:25].contiguous(), target.unsqueeze(0).contiguous(),
                        stage1_feature_fn(feat), torch_vgg, first_inv_lr=8e-3, cam_lr=6e-6, translation_lr=2e-4, graphed=False)
    named = list(st.noise_bufs.items()) + list(st.noise_bufs2.items())
    stage1_noise_init(named)
    with torch.no_grad():
        st.translation_opt.copy_(cfg['translation'].cuda())
    # the four uniform draws of the two renders (predicted camera, canonical camera), in the order the reference consumed them
    G.renderer.fixed_noise = [(torch.from_numpy(fx['draw0']), torch.from_numpy(fx['draw1'])), (torch.from_numpy(fx['draw2']), torch.from_numpy(fx['draw3']))]
    loss = st.step(cfg['w_noise'].cuda(), lr=float(fx['lr']))
    parts = {k: float(v) for k, v in st.parts.items()}
    print('loss', loss.item(), float(fx['loss']), parts, float(fx['dist']), float(fx['warp_loss']), float(fx['reg_loss']))
    assert abs(parts['dist'] - float(fx['dist'])) <= 2e-3 * float(fx['dist'])
    assert abs(parts['warp'] - float(fx['warp_loss'])) <= 2e-3 * float(fx['warp_loss'])
    assert abs(parts['reg'] - float(fx['reg_loss'])) <= 1e-4 * float(fx['reg_loss'])
    assert abs(loss.item() - float(fx['loss'])) <= 1e-3 * float(fx['loss'])
    assert relerr(st.w_opt.grad, torch.from_numpy(fx['grad_w_opt'])) < 1e-2
    assert relerr(st.translation_opt.grad, torch.from_numpy(fx['grad_translation'])) < 1e-2
    assert relerr(pose[2].bias.grad, torch.from_numpy(fx['grad_pose_b'])) < 1e-2
    # after the optimiser steps: Adam's first update is +-lr per element, so compare the MOVES (a sign flip of a near-zero gradient
    # is a full 2 lr on that element)
    for mine, start, key in ((st.w_opt, cfg['w_start'], 'w_opt_after'), (st.translation_opt, cfg['translation'], 'translation_after')):
        move, ref = mine.detach().cpu() - start, torch.from_numpy(fx[key]) - start
        assert relerr(move, ref) < 0.15, key
    assert relerr(pose[2].bias.detach() - stage1_pose_net(13)[2].bias.cuda(), torch.from_numpy(fx['pose_b_after']) - stage1_pose_net(13)[2].bias) < 0.15
    for name, buf in named:
        key = name.replace('.', '_')
        stride = max(1, buf.shape[0] // 32)
        assert relerr(buf.detach()[::stride, ::stride], torch.from_numpy(fx['noise_after_' + key])) < 2e-2, name
        mom = fx['noise_after_mom_' + key]
        assert abs(buf.detach().double().square().sum().item() - mom[1]) <= 1e-3 * mom[1], name          # unit second moment after w_projector.py:262-268
        assert relerr(buf.grad.norm().reshape(1), torch.tensor([float(fx['noise_grad_norm_' + key])])) < 2e-2, name
