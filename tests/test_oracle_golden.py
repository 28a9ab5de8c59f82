"""CPU: the oracle restatement reproduces the fixtures generated from the real reference
(oracle/make_goldens.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

import eg3d_oracle as oracle
import synth_params as sp
from golden_util import build_param_dict, check_image, check_param_grads, load_case, oracle_noise_args


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


FAST = ['tiny_r64_s16', 'tiny_r32_s8_n2_white', 'tiny_r64_s12_noimp', 'tiny_r32_s8_random', 'tiny_r32_s8_auto', 'tiny_r32_s8_disp']
SLOW = ['full_r64_s16']


def run_oracle(case, P, ws, c, **kw):
    nr, dd = oracle_noise_args(case)
    return oracle.synthesis(P, ws, c, case.rk, case.R, case.u_strat, case.u_imp, noise_mode=case.noise_mode, noise_random=nr,
                            density_draws=dd, **kw)


@pytest.mark.parametrize('name', FAST + SLOW)
def test_oracle_forward_matches_reference_fixture(name, golden_dir):
    case = load_case(golden_dir, name)
    P = build_param_dict(case)
    with torch.no_grad():
        out = run_oracle(case, P, case.ws, case.c, return_planes=True)
    assert np.abs(out['image_raw'].numpy() - case.fx['image_raw']).max() < 2e-5
    assert np.abs(out['image_depth'].numpy() - case.fx['image_depth']).max() < 2e-5
    d_sub, d_tile = check_image(out['image'], case.fx, 5e-5)
    assert d_sub < 5e-5 and d_tile < 5e-5
    assert np.abs(out['planes'][..., ::8, ::8].numpy() - case.fx['planes_sub8']).max() < 5e-5
    mom = case.fx['image_mom']
    assert abs(out['image'].double().sum().item() - mom[0]) < 1e-4 * max(1.0, abs(mom[0]))


@pytest.mark.parametrize('name', ['tiny_r64_s16', 'tiny_r32_s8_n2_white', 'tiny_r32_s8_random', 'tiny_r32_s8_auto', 'tiny_r32_s8_disp'])
def test_oracle_gradients_match_reference_fixture(name, golden_dir):
    case = load_case(golden_dir, name)
    P = build_param_dict(case, requires_grad=True)
    ws = case.ws.clone().requires_grad_(True)
    c = case.c.clone().requires_grad_(True)
    out = run_oracle(case, P, ws, c)
    loss = oracle.pti_loss(out, case.t512, case.t_raw)
    assert abs(loss.item() - case.fx['loss'][0]) < 1e-5
    loss.backward()
    assert rel_l2(ws.grad.numpy(), case.fx['grad_ws']) < 1e-3
    assert rel_l2(c.grad.numpy(), case.fx['grad_c']) < 1e-3
    worst = check_param_grads({k: v.grad for k, v in P.items()}, case.fx, 1e-3)
    assert max(v[0] for v in worst.values()) < 1e-2, worst


def test_fixture_inventory(golden_dir):
    for n in sp.GOLDEN_CASES:
        assert os.path.exists(os.path.join(golden_dir, n + '.npz')), n
    for n in ('manifest_full.json', 'manifest_tiny.json', 'mapping.npz', 'stage1_warp.npz'):
        assert os.path.exists(os.path.join(golden_dir, n)), n


def test_fixtures_resolve_channels(golden_dir):
    """The v2 fixtures are channel-resolved: permuting two output channels of one weight gradient is detected."""
    case = load_case(golden_dir, 'tiny_r64_s16')
    P = build_param_dict(case, requires_grad=True)
    out = run_oracle(case, P, case.ws, case.c)
    oracle.pti_loss(out, case.t512, case.t_raw).backward()
    grads = {k: (v.grad.clone() if v.grad is not None else None) for k, v in P.items()}
    check_param_grads(grads, case.fx, 1e-3)
    g = grads['backbone.synthesis.b16.conv1.weight']
    g[[3, 5]] = g[[5, 3]]
    with pytest.raises(AssertionError):
        check_param_grads(grads, case.fx, 1e-3)
    img = out['image'].detach().clone()
    img[0, 1, 17, 33] += 0.05                                  # an odd pixel: invisible to a [::2, ::2] (or [::4, ::4]) sub-sample
    d_sub, d_tile = check_image(img, case.fx, 1e-3)
    assert d_sub < 5e-5 and d_tile > 1e-4                      # ... but it moves its 16x16 tile sum
