"""CPU: the oracle restatement reproduces the fixtures generated from the real reference
(oracle/make_goldens.py).  No GPU, no /root/reference needed."""
import os

import numpy as np
import pytest
import torch

import eg3d_oracle as oracle
import synth_params as sp
from golden_util import build_param_dict, load_case

def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


FAST = ['tiny_r64_s16', 'tiny_r32_s8_n2_white', 'tiny_r64_s12_noimp']
SLOW = ['full_r64_s16']


@pytest.mark.parametrize('name', FAST + SLOW)
def test_oracle_forward_matches_reference_fixture(name, golden_dir):
    case = load_case(golden_dir, name)
    P = build_param_dict(case)
    with torch.no_grad():
        out = oracle.synthesis(P, case.ws, case.c, case.rk, case.R, case.u_strat, case.u_imp, return_planes=True)
    assert np.abs(out['image_raw'].numpy() - case.fx['image_raw']).max() < 2e-5
    assert np.abs(out['image_depth'].numpy() - case.fx['image_depth']).max() < 2e-5
    assert np.abs(out['image'][..., ::4, ::4].numpy() - case.fx['image_sub4']).max() < 5e-5
    assert np.abs(out['planes'][..., ::8, ::8].numpy() - case.fx['planes_sub8']).max() < 5e-5
    mom = case.fx['image_mom']
    assert abs(out['image'].double().sum().item() - mom[0]) < 1e-4 * max(1.0, abs(mom[0]))


@pytest.mark.parametrize('name', ['tiny_r64_s16', 'tiny_r32_s8_n2_white'])
def test_oracle_gradients_match_reference_fixture(name, golden_dir):
    case = load_case(golden_dir, name)
    P = build_param_dict(case, requires_grad=True)
    ws = case.ws.clone().requires_grad_(True)
    c = case.c.clone().requires_grad_(True)
    out = oracle.synthesis(P, ws, c, case.rk, case.R, case.u_strat, case.u_imp)
    loss = oracle.pti_loss(out, case.t512, case.t_raw)
    assert abs(loss.item() - case.fx['loss'][0]) < 1e-5
    loss.backward()
    assert rel_l2(ws.grad.numpy(), case.fx['grad_ws']) < 1e-3
    assert rel_l2(c.grad.numpy(), case.fx['grad_c']) < 1e-3
    for i, n in enumerate(case.fx['grad_names']):
        g = P[str(n)].grad
        assert g is not None, n
        ssq = g.double().square().sum().item()
        ref = case.fx['grad_mom'][i][1]
        assert abs(ssq - ref) <= 2e-3 * max(ref, 1e-20), (n, ssq, ref)


def test_fixture_inventory(golden_dir):
    for n in FAST + SLOW + ['full_r128_s48', 'full_r256_s96']:
        assert os.path.exists(os.path.join(golden_dir, n + '.npz')), n
