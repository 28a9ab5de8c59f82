"""CPU oracle for the EG3D tri-plane generator hot path (TEST INFRASTRUCTURE ONLY).

This file is a functional restatement, in plain torch CPU ops, of what the
reference computes on the path  TriPlaneGenerator.synthesis()  ->  StyleGAN2
backbone -> ImportanceRenderer -> SuperresolutionHybrid8X.  It is *not* part of
the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.  The product path (b200eg3d) never does.

Parity status: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, imported read-only in the build container by oracle/make_goldens.py and
committed under tests/golden/ (forward outputs and gradients, two generator
sizes).  tests/test_oracle_golden.py re-checks that on every run.

Every function cites the reference file:line whose behaviour it restates.
All functions are dtype-generic (fp32 / fp64) and differentiable by autograd.

The parameter container is a flat ``dict[str, Tensor]`` keyed exactly like the
reference ``G.state_dict()`` (e.g. ``backbone.synthesis.b8.conv0.affine.weight``).
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)


# ----------------------------------------------------------------------------
# Elementwise / FIR primitives

def fir_1331(dtype=torch.float32):
    """[1,3,3,1] (x) [1,3,3,1] / 64.  torch_utils/ops/upfirdn2d.py:72-116 (4 taps < 8 => outer product, normalised)."""
    k = torch.tensor([1.0, 3.0, 3.0, 1.0], dtype=torch.float64)
    f = torch.outer(k, k)
    return (f / f.sum()).to(dtype)


def upfirdn2d(x, f, up=1, down=1, pad=(0, 0, 0, 0), gain=1.0):
    """Zero-insert upsample, pad/crop, FIR (true convolution), decimate.
    torch_utils/ops/upfirdn2d.py:169-213 (_upfirdn2d_ref).  pad = (x0, x1, y0, y1)."""
    n, c, h, w = x.shape
    px0, px1, py0, py1 = pad
    if up > 1:
        z = x.new_zeros(n, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.reshape(n, c, h * up, w * up)
    x = F.pad(x, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    x = x[:, :, max(-py0, 0): x.shape[2] - max(-py1, 0), max(-px0, 0): x.shape[3] - max(-px1, 0)]
    k = (f.to(x.dtype) * gain).flip([0, 1])  # F.conv2d correlates; flipping gives convolution
    k = k[None, None].repeat(c, 1, 1, 1)
    x = F.conv2d(x, k, groups=c)
    return x[:, :, ::down, ::down]


def upsample2d(x, f):
    """x2 FIR upsampling used on the RGB / plane skip path.
    torch_utils/ops/upfirdn2d.py:315-350: pad = [(fw+up-1)//2, (fw-up)//2]*2 = [2,1,2,1], gain = up**2."""
    return upfirdn2d(x, f, up=2, pad=(2, 1, 2, 1), gain=4.0)


def bias_act(x, b=None, act='linear', gain=None, clamp=None, alpha=0.2):
    """y = clamp(act(x + b) * gain).  torch_utils/ops/bias_act.py:93-122; kernel bias_act.cu:28-151."""
    if gain is None:
        gain = SQRT2 if act == 'lrelu' else 1.0
    if b is not None:
        x = x + b.to(x.dtype).reshape(1, -1, *([1] * (x.ndim - 2)))
    if act == 'lrelu':
        x = F.leaky_relu(x, alpha)
    elif act != 'linear':
        raise ValueError(act)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def fully_connected(x, weight, bias, lr_mul=1.0, act='linear'):
    """training/networks_stylegan2.py:114-127 (FullyConnectedLayer.forward): weight gain lr_mul/sqrt(fan_in), bias gain lr_mul."""
    w = weight.to(x.dtype) * (lr_mul / math.sqrt(weight.shape[1]))
    y = x @ w.t()
    b = bias.to(x.dtype) * lr_mul if bias is not None else None
    if act == 'linear':
        return y + b if b is not None else y
    return bias_act(y, b, act=act)


# ----------------------------------------------------------------------------
# Modulated convolution (training/networks_stylegan2.py:34-91) + resampling
# (torch_utils/ops/conv2d_resample.py:48-143)

def modulated_conv2d(x, weight, styles, noise=None, up=1, demodulate=True, f=None):
    """Fused-path semantics of modulated_conv2d (networks_stylegan2.py:58-91, fp32 branch).

    up == 1: correlation with padding k//2 (flip_weight=True, conv2d_resample.py:134-136).
    up == 2: stride-2 transposed conv (true convolution, conv2d_resample.py:113-127) giving (2H+1)^2,
             then 4x4 FIR with pad [1,1,1,1] and gain 4 (conv2d_resample.py:128) giving (2H)^2.
    """
    n, cin, h, w_ = x.shape
    cout, _, kh, kw = weight.shape
    w = weight.to(x.dtype)[None] * styles.to(x.dtype).reshape(n, 1, cin, 1, 1)          # [N,O,I,k,k]
    if demodulate:
        d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()                              # [N,O]
        w = w * d.reshape(n, cout, 1, 1, 1)
    outs = []
    for i in range(n):                                                                  # groups=N in the reference
        xi = x[i:i + 1]
        wi = w[i]
        if up == 1:
            yi = F.conv2d(xi, wi, padding=kh // 2)
        else:
            assert up == 2 and kh == 3
            yi = F.conv_transpose2d(xi, wi.transpose(0, 1), stride=2)
            yi = upfirdn2d(yi, f, pad=(1, 1, 1, 1), gain=4.0)
        outs.append(yi)
    y = torch.cat(outs, 0)
    if noise is not None:
        y = y + noise.to(y.dtype)
    return y


def synthesis_layer(P, pre, x, w, res, up, noise_mode, conv_clamp, f, gain=1.0, noise_random=None):
    """training/networks_stylegan2.py:311-330 (SynthesisLayer.forward)."""
    styles = fully_connected(w, P[pre + 'affine.weight'], P[pre + 'affine.bias'])
    noise = None
    if noise_mode == 'const':
        noise = P[pre + 'noise_const'] * P[pre + 'noise_strength']
    elif noise_mode == 'random':
        noise = noise_random * P[pre + 'noise_strength']
    y = modulated_conv2d(x, P[pre + 'weight'], styles, noise=noise, up=up, f=f)
    clamp = conv_clamp * gain if conv_clamp is not None else None
    return bias_act(y, P[pre + 'bias'], act='lrelu', gain=SQRT2 * gain, clamp=clamp)


def torgb_layer(P, pre, x, w, conv_clamp):
    """training/networks_stylegan2.py:353-357 (ToRGBLayer.forward): 1x1 modconv without demodulation."""
    cin = P[pre + 'weight'].shape[1]
    styles = fully_connected(w, P[pre + 'affine.weight'], P[pre + 'affine.bias']) * (1.0 / math.sqrt(cin))
    y = modulated_conv2d(x, P[pre + 'weight'], styles, demodulate=False)
    return bias_act(y, P[pre + 'bias'], act='linear', clamp=conv_clamp)


def synthesis_block(P, pre, x, img, ws, res, first, noise_mode, conv_clamp, f, noise_random=None):
    """training/networks_stylegan2.py:417-461 (SynthesisBlock.forward), 'skip' architecture, fp32."""
    n = ws.shape[0]
    wi = 0
    nr = noise_random or {}
    if first:
        x = P[pre + 'const'].to(ws.dtype)[None].repeat(n, 1, 1, 1)
    else:
        x = synthesis_layer(P, pre + 'conv0.', x, ws[:, wi], res, 2, noise_mode, conv_clamp, f, noise_random=nr.get(pre + 'conv0.'))
        wi += 1
    x = synthesis_layer(P, pre + 'conv1.', x, ws[:, wi], res, 1, noise_mode, conv_clamp, f, noise_random=nr.get(pre + 'conv1.'))
    wi += 1
    if img is not None:
        img = upsample2d(img, f)
    y = torgb_layer(P, pre + 'torgb.', x, ws[:, wi], conv_clamp)
    img = img + y if img is not None else y
    return x, img


def backbone_synthesis(P, ws, noise_mode='const', conv_clamp=256, img_resolution=256, noise_random=None, pre='backbone.synthesis.'):
    """training/networks_stylegan2.py:503-518 (SynthesisNetwork.forward): blocks 4..img_resolution,
    block k reads ws[:, idx : idx+num_conv+num_torgb], idx += num_conv."""
    f = fir_1331(ws.dtype).to(ws.device)
    x = img = None
    w_idx = 0
    res = 4
    while res <= img_resolution:
        first = res == 4
        n_conv = 1 if first else 2
        x, img = synthesis_block(P, f'{pre}b{res}.', x, img, ws[:, w_idx:w_idx + n_conv + 1], res, first,
                                 noise_mode, conv_clamp, f, noise_random)
        w_idx += n_conv
        res *= 2
    return img


def superresolution_8x(P, rgb, x, ws, noise_mode='none', sr_antialias=True, noise_random=None, pre='superresolution.'):
    """training/superresolution.py:45-56 (SuperresolutionHybrid8X.forward); conv_clamp 256 because sr_num_fp16_res>0 (:36-42)."""
    f = fir_1331(ws.dtype).to(ws.device)
    w3 = ws[:, -1:, :].repeat(1, 3, 1)
    if x.shape[-1] != 128:
        x = F.interpolate(x, size=(128, 128), mode='bilinear', align_corners=False, antialias=sr_antialias)
        rgb = F.interpolate(rgb, size=(128, 128), mode='bilinear', align_corners=False, antialias=sr_antialias)
    x, rgb = synthesis_block(P, pre + 'block0.', x, rgb, w3, 256, False, noise_mode, 256, f, noise_random)
    x, rgb = synthesis_block(P, pre + 'block1.', x, rgb, w3, 512, False, noise_mode, 256, f, noise_random)
    return rgb


# ----------------------------------------------------------------------------
# Rays (training/volumetric_rendering/ray_sampler.py:24-73)

def ray_sampler(cam2world, intrinsics, resolution):
    n = cam2world.shape[0]
    dt = cam2world.dtype
    fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
    cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
    idx = (torch.arange(resolution, dtype=torch.float32, device=cam2world.device) * (1.0 / resolution) + (0.5 / resolution)).to(dt)
    # flat ray m = i*R + j  ->  x = (j+.5)/R, y = (i+.5)/R        (ray_sampler.py:45-51)
    y_cam = idx.repeat_interleave(resolution)[None].expand(n, -1)
    x_cam = idx.repeat(resolution)[None].expand(n, -1)
    x_lift = (x_cam - cx[:, None] + cy[:, None] * sk[:, None] / fy[:, None] - sk[:, None] * y_cam / fy[:, None]) / fx[:, None]
    y_lift = (y_cam - cy[:, None]) / fy[:, None]
    pts = torch.stack([x_lift, y_lift, torch.ones_like(x_lift), torch.ones_like(x_lift)], dim=-1)   # [N,M,4]
    world = torch.einsum('nij,nmj->nmi', cam2world, pts)[:, :, :3]
    origin = cam2world[:, :3, 3]
    dirs = F.normalize(world - origin[:, None, :], dim=2)
    return origin[:, None, :].expand(-1, dirs.shape[1], -1), dirs


# ----------------------------------------------------------------------------
# Tri-plane sampling + OSG decoder (renderer.py:39-66, triplane.py:124-136)

def sample_triplane_features(planes, coords, box_warp):
    """planes [N,3,C,H,W], coords [N,P,3] -> mean over the 3 planes of bilinear samples [N,P,C].
    Plane axes after inverting generate_planes() (renderer.py:23-53): P0<-(x,y), P1<-(x,z), P2<-(z,x);
    grid_sample x indexes W; align_corners=False, zero padding (renderer.py:64)."""
    n, _, c, h, w = planes.shape
    g = coords * (2.0 / box_warp)
    uv = torch.stack([g[..., [0, 1]], g[..., [0, 2]], g[..., [2, 0]]], dim=1)           # [N,3,P,2]
    feats = F.grid_sample(planes.reshape(n * 3, c, h, w), uv.reshape(n * 3, 1, -1, 2).to(planes.dtype),
                          mode='bilinear', padding_mode='zeros', align_corners=False)    # [N*3,C,1,P]
    return feats.reshape(n, 3, c, -1).mean(1).permute(0, 2, 1)


def osg_decoder(P, feats, lr_mul=1.0, pre='decoder.net.'):
    """triplane.py:124-136: FC(32->64) softplus FC(64->33); sigma = o[0]; rgb = sigmoid(o[1:])*1.002-0.001."""
    h = F.softplus(fully_connected(feats, P[pre + '0.weight'], P[pre + '0.bias'], lr_mul))
    o = fully_connected(h, P[pre + '2.weight'], P[pre + '2.bias'], lr_mul)
    rgb = torch.sigmoid(o[..., 1:]) * (1 + 2 * 0.001) - 0.001
    return rgb, o[..., 0:1]


def run_model(P, planes, coords, rk, density_draw=None):
    """renderer.py:197-203.  density_draw: the standard-normal tensor the reference draws with randn_like(sigma) when
    rendering_kwargs['density_noise'] > 0 (:201-202); None = no density noise."""
    feats = sample_triplane_features(planes, coords, rk['box_warp'])
    rgb, sigma = osg_decoder(P, feats, rk.get('decoder_lr_mul', 1))
    if rk.get('density_noise', 0) > 0:
        assert density_draw is not None, 'density_noise > 0 needs the randn_like draw'
        sigma = sigma + density_draw.reshape(sigma.shape).to(sigma.dtype) * rk['density_noise']
    return rgb, sigma


# ----------------------------------------------------------------------------
# Ray marching (ray_marcher.py:25-57) and hierarchical sampling (renderer.py:212-308)

def ray_march(colors, densities, depths, white_back=False):
    deltas = depths[:, :, 1:] - depths[:, :, :-1]
    c_mid = (colors[:, :, :-1] + colors[:, :, 1:]) / 2
    s_mid = F.softplus((densities[:, :, :-1] + densities[:, :, 1:]) / 2 - 1)
    t_mid = (depths[:, :, :-1] + depths[:, :, 1:]) / 2
    alpha = 1 - torch.exp(-s_mid * deltas)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2), -2)[:, :, :-1]
    weights = alpha * trans
    rgb = (weights * c_mid).sum(-2)
    wsum = weights.sum(2)
    depth = (weights * t_mid).sum(-2) / wsum
    depth = torch.nan_to_num(depth, float('inf'))
    depth = torch.clamp(depth, torch.min(depths), torch.max(depths))     # global min/max (ray_marcher.py:50)
    if white_back:
        rgb = rgb + 1 - wsum
    return rgb * 2 - 1, depth, weights


def ray_limits_box(ray_o, ray_d, box_side_length):
    """math_utils.py:46-98 (get_ray_limits_box): slab intersection of every ray with the cube of side box_side_length about
    the origin; (-1, -2) marks a miss.  Restated with where() instead of index_select / masked assignment."""
    o, d = ray_o.detach().reshape(-1, 3), ray_d.detach().reshape(-1, 3)
    half = box_side_length / 2
    inv = 1 / d
    neg = inv < 0
    lo = torch.where(neg, torch.full_like(o, half), torch.full_like(o, -half))       # bounds[sign]
    hi = torch.where(neg, torch.full_like(o, -half), torch.full_like(o, half))       # bounds[1 - sign]
    t0, t1 = (lo - o) * inv, (hi - o) * inv
    tmin, tmax = t0[:, 0], t1[:, 0]
    valid = ~((tmin > t1[:, 1]) | (t0[:, 1] > tmax))
    tmin, tmax = torch.max(tmin, t0[:, 1]), torch.min(tmax, t1[:, 1])
    valid = valid & ~((tmin > t1[:, 2]) | (t0[:, 2] > tmax))
    tmin, tmax = torch.max(tmin, t0[:, 2]), torch.min(tmax, t1[:, 2])
    tmin = torch.where(valid, tmin, torch.full_like(tmin, -1))
    tmax = torch.where(valid, tmax, torch.full_like(tmax, -2))
    return tmin.reshape(*ray_o.shape[:-1], 1), tmax.reshape(*ray_o.shape[:-1], 1)


def stratified_depths(n, m, s, ray_start, ray_end, u, dtype, disparity=False):
    """renderer.py:224-247 (sample_stratified): scalar limits, per-ray tensor limits ('auto'), or disparity-space sampling."""
    u = u.to(dtype)
    if disparity:
        t = torch.linspace(0, 1, s, device=u.device).to(dtype).reshape(1, 1, s, 1).repeat(n, m, 1, 1) + u * (1 / (s - 1))
        return 1. / (1. / ray_start * (1. - t) + 1. / ray_end * t)
    if isinstance(ray_start, torch.Tensor):                                          # [N,M,1] each
        steps = (torch.arange(s, dtype=torch.float32, device=u.device) / (s - 1)).to(dtype).reshape(1, 1, s, 1)
        t = ray_start.unsqueeze(2) + steps * (ray_end - ray_start).unsqueeze(2)
        return t + u * ((ray_end - ray_start) / (s - 1)).unsqueeze(-1)
    t = torch.linspace(ray_start, ray_end, s, device=u.device).to(dtype).reshape(1, 1, s, 1).repeat(n, m, 1, 1)
    return t + u * ((ray_end - ray_start) / (s - 1))


def importance_depths(depths, weights, n_imp, u, eps=1e-5):
    """renderer.py:249-308 (sample_importance + sample_pdf), no gradient."""
    with torch.no_grad():
        n, m, s, _ = depths.shape
        z = depths.reshape(n * m, s)
        w = weights.reshape(n * m, -1)
        w = F.max_pool1d(w.unsqueeze(1), 2, 1, padding=1)
        w = F.avg_pool1d(w, 2, 1).squeeze(1) + 0.01
        bins = 0.5 * (z[:, :-1] + z[:, 1:])
        pw = w[:, 1:-1] + eps
        pdf = pw / pw.sum(-1, keepdim=True)
        cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
        u = u.to(z.dtype).contiguous()
        inds = torch.searchsorted(cdf, u, right=True)
        below = (inds - 1).clamp_min(0)
        above = inds.clamp_max(pw.shape[1])
        cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
        bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
        den = cdf_hi - cdf_lo
        den = torch.where(den < eps, torch.ones_like(den), den)
        t = bin_lo + (u - cdf_lo) / den * (bin_hi - bin_lo)
        return t.reshape(n, m, n_imp, 1)


def render(P, planes, ray_o, ray_d, rk, u_strat, u_imp, density_draws=(None, None)):
    """renderer.py:143-195 (ImportanceRenderer.forward).  density_draws: the (coarse, fine) randn_like draws of :201-202."""
    n, m, _ = ray_o.shape
    s = rk['depth_resolution']
    ray_start, ray_end = rk['ray_start'], rk['ray_end']
    if ray_start == 'auto' and ray_end == 'auto':                                    # renderer.py:146-152
        ray_start, ray_end = ray_limits_box(ray_o, ray_d, rk['box_warp'])
        ok = ray_end > ray_start
        if bool(ok.any()):
            ray_start = torch.where(ok, ray_start, ray_start[ok].min())
            ray_end = torch.where(ok, ray_end, ray_start[ok].max())
    t_c = stratified_depths(n, m, s, ray_start, ray_end, u_strat, ray_o.dtype, rk.get('disparity_space_sampling', False))
    pts = (ray_o.unsqueeze(-2) + t_c * ray_d.unsqueeze(-2)).reshape(n, -1, 3)
    rgb_c, sig_c = run_model(P, planes, pts, rk, density_draws[0])
    rgb_c = rgb_c.reshape(n, m, s, -1)
    sig_c = sig_c.reshape(n, m, s, 1)
    n_imp = rk['depth_resolution_importance']
    wb = rk.get('white_back', False)
    if n_imp > 0:
        _, _, w_c = ray_march(rgb_c, sig_c, t_c, wb)
        t_f = importance_depths(t_c, w_c, n_imp, u_imp)
        pts = (ray_o.unsqueeze(-2) + t_f * ray_d.unsqueeze(-2)).reshape(n, -1, 3)
        rgb_f, sig_f = run_model(P, planes, pts, rk, density_draws[1])
        rgb_f = rgb_f.reshape(n, m, n_imp, -1)
        sig_f = sig_f.reshape(n, m, n_imp, 1)
        t_all = torch.cat([t_c, t_f], -2)
        _, order = torch.sort(t_all, dim=-2)
        t_all = torch.gather(t_all, -2, order)
        rgb_all = torch.gather(torch.cat([rgb_c, rgb_f], -2), -2, order.expand(-1, -1, -1, rgb_c.shape[-1]))
        sig_all = torch.gather(torch.cat([sig_c, sig_f], -2), -2, order)
        rgb, depth, w = ray_march(rgb_all, sig_all, t_all, wb)
    else:
        rgb, depth, w = ray_march(rgb_c, sig_c, t_c, wb)
    return rgb, depth, w.sum(2)


# ----------------------------------------------------------------------------
# Whole generator (training/triplane.py:53-90)

def draw_depth_noise(seed, n, m, s, s_imp, tensor_limits=False):
    """The two RNG draws the renderer makes, in order (renderer.py:245 rand_like [N,M,S,1]; :292 rand [N*M,S_imp]).
    tensor_limits ('auto' ray limits, renderer.py:240-242): depths_coarse is a permuted view of a [S,N,M,1] tensor there and
    rand_like preserves that memory layout, so the draws fill [S,N,M,1] order and are read as [N,M,S,1]."""
    g = torch.Generator().manual_seed(seed)
    if tensor_limits:
        u1 = torch.rand(s, n, m, 1, generator=g).permute(1, 2, 0, 3).contiguous()
    else:
        u1 = torch.rand(n, m, s, 1, generator=g)
    return u1, torch.rand(n * m, s_imp, generator=g)


def synthesis(P, ws, c, rk, neural_rendering_resolution, u_strat, u_imp, noise_mode='const',
              conv_clamp=256, return_planes=False, noise_random=None, density_draws=(None, None)):
    """TriPlaneGenerator.synthesis (triplane.py:53-90) with force_fp32=True semantics.
    noise_random: {layer prefix -> [N,1,res,res] standard-normal draw} for noise_mode='random' (networks_stylegan2.py:319)."""
    n = ws.shape[0]
    r = neural_rendering_resolution
    cam2world = c[:, :16].reshape(-1, 4, 4)
    intr = c[:, 16:25].reshape(-1, 3, 3)
    ray_o, ray_d = ray_sampler(cam2world, intr, r)
    planes96 = backbone_synthesis(P, ws, noise_mode=noise_mode, conv_clamp=conv_clamp, noise_random=noise_random)
    planes = planes96.reshape(n, 3, 32, planes96.shape[-2], planes96.shape[-1])
    feat, depth, _ = render(P, planes, ray_o, ray_d, rk, u_strat, u_imp, density_draws)
    feat_img = feat.permute(0, 2, 1).reshape(n, feat.shape[-1], r, r)
    depth_img = depth.permute(0, 2, 1).reshape(n, 1, r, r)
    rgb = feat_img[:, :3]
    sr = superresolution_8x(P, rgb, feat_img, ws, noise_mode=rk['superresolution_noise_mode'],
                            sr_antialias=rk['sr_antialias'])
    out = {'image': sr, 'image_raw': rgb, 'image_depth': depth_img}
    if return_planes:
        out['planes'] = planes96
    return out


# ----------------------------------------------------------------------------
# PTI-step stand-in loss (training/coaches/base_coach.py:101-126, 294-305), LPIPS omitted (weights unavailable)

def tv_norm(depth):
    """base_coach.py:294-305 (compute_tv_norm) restated: mean of squared forward differences in x and y."""
    dx = depth[..., :-1, 1:] - depth[..., :-1, :-1]
    dy = depth[..., 1:, :-1] - depth[..., :-1, :-1]
    return (dx.square() + dy.square()).mean()


def pti_loss(out, target512, target_raw):
    return F.mse_loss(out['image'], target512) + F.mse_loss(out['image_raw'], target_raw) + tv_norm(out['image_depth'])


def calc_loss(out, real_images, pt_l2_lambda=1.0, depth_tv_lambda=1.0):
    """base_coach.py:101-126 (calc_loss) restated without the LPIPS term: the raw-resolution target is the area-resized
    real image (:103), both MSE terms are weighted by pt_l2_lambda (:105-109), the depth TV term is added last (:123-124).
    Returns (loss, [mse_image, mse_raw, tv])."""
    r = out['image_raw'].shape[-1]
    real_r = F.interpolate(real_images, size=(r, r), mode='area')
    a, b, t = F.mse_loss(out['image'], real_images), F.mse_loss(out['image_raw'], real_r), tv_norm(out['image_depth'])
    return pt_l2_lambda * (a + b) + depth_tv_lambda * t, [a, b, t]
