"""Drop-in TriPlaneGenerator for the EG3D inversion hot path, running on the b200eg3d CUDA library.

Mirrors the module tree, parameter / buffer names, method signatures and keyword arguments of the reference
(training/triplane.py:19-110, training/networks_stylegan2.py:96-552, training/superresolution.py:29-56,
training/volumetric_rendering/{renderer,ray_sampler}.py) so that `misc.copy_params_and_buffers(G_ref, G, require_all=True)`
and every call site listed in SURVEY.md section 3.5 work unchanged.  Only the compute differs: activations are NHWC
internally (exposed as channels_last NCHW views, zero-copy) and every stage is a hand-written sm_100a kernel.

Precision: the whole path computes in fp32 regardless of `force_fp32` (the reference's optional fp16 blocks,
networks_stylegan2.py:421-423, are a speed/precision trade the B200 path does not need); the kwarg is accepted.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops
from ._lib import on_device


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def _to_nhwc(x):
    """[N,C,H,W] (any strides) -> contiguous [N,H,W,C]; free when x is already channels_last."""
    return x.permute(0, 2, 3, 1).contiguous()


def _to_nchw_view(x):
    """contiguous [N,H,W,C] -> [N,C,H,W] view (channels_last strides, no copy)."""
    return x.permute(0, 3, 1, 2)


class FullyConnectedLayer(torch.nn.Module):
    """networks_stylegan2.py:96-130."""

    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.activation = activation
        self.lr_multiplier = lr_multiplier
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            return torch.addmm(b.unsqueeze(0), x, w.t())
        x = x.matmul(w.t())
        return ops.bias_act(x, b, act=self.activation)

    def extra_repr(self):
        return f'in_features={self.in_features:d}, out_features={self.out_features:d}, activation={self.activation:s}'


class MappingNetwork(torch.nn.Module):
    """networks_stylegan2.py:175-265 (off the per-step hot path: called once per image, w_projector.py:93)."""

    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.998):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = z_dim, c_dim, w_dim, num_ws
        self.num_layers, self.w_avg_beta = num_layers, w_avg_beta
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx in range(num_layers):
            setattr(self, f'fc{idx}', FullyConnectedLayer(features[idx], features[idx + 1], activation=activation,
                                                          lr_multiplier=lr_multiplier))
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        x = None
        with ops.prof_range('input'):
            if self.z_dim > 0:
                x = normalize_2nd_moment(z.to(torch.float32))
            if self.c_dim > 0:
                y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
                x = torch.cat([x, y], dim=1) if x is not None else y
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if update_emas and self.w_avg_beta is not None:
            with ops.prof_range('update_w_avg'):
                self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            with ops.prof_range('broadcast'):
                x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            with ops.prof_range('truncate'):
                if self.num_ws is None or truncation_cutoff is None:
                    x = self.w_avg.lerp(x, truncation_psi)
                else:
                    x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


class SynthesisLayer(torch.nn.Module):
    """networks_stylegan2.py:276-336.  forward() takes and returns NHWC activations."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False):
        super().__init__()
        if kernel_size != 3 or activation != 'lrelu' or up not in (1, 2):
            raise NotImplementedError('b200eg3d SynthesisLayer supports kernel_size=3, lrelu, up in {1,2}')
        self.in_channels, self.out_channels, self.w_dim, self.resolution = in_channels, out_channels, w_dim, resolution
        self.up, self.use_noise, self.activation, self.conv_clamp = up, use_noise, activation, conv_clamp
        self.register_buffer('resample_filter', ops.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = math.sqrt(2)
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1, x_split=None, bank=None, lidx=-1, want_z=True, fork=False):
        """Returns (z, split) where split is the (hi, lo) bf16 pair of z for the next tensor-core conv, or None; with fork=True
        (z, split, z_second): a second handle of z for its second consumer (ops.modconv_layer).
        bank / lidx: styles and modulated weights come from a pre-computed ops.WeightBank entry (w is then unused).
        want_z=False: the caller's consumers all read the pair (ops.lean_ok): z comes back as a shape-only placeholder."""
        assert noise_mode in ['random', 'const', 'none']
        if not fused_modconv:
            z = self._forward_unfused(x, w, noise_mode, gain)
            return (z, None, z) if fork else (z, None)
        styles = self.affine(w) if bank is None else None
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device)
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return ops.modconv_layer(x, self.weight, styles, self.bias, noise, self.noise_strength if noise is not None else None,
                                 self.up, self.act_gain * gain, clamp, x_split=x_split, bank=bank, lidx=lidx, want_z=want_z, fork=fork)


    def _forward_unfused(self, x, w, noise_mode, gain):
        """fused_modconv=False (networks_stylegan2.py:70-79): scale the activations by the styles, convolve with the SHARED weight,
        scale by the demodulation coefficients afterwards.  Algebraically the same layer; it runs on the op-level look-alikes
        (shims.ops_modules: conv2d_resample, fma, bias_act) and is the cross-check of the fused kernels, not a fast path."""
        from .shims import ops_modules as om
        styles = self.affine(w)
        xn = _to_nchw_view(x)
        n = xn.shape[0]
        wt = self.weight.to(torch.float32)
        dcoefs = ((wt.unsqueeze(0) * styles.reshape(n, 1, -1, 1, 1)).square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([n, 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        y = xn * styles.reshape(n, -1, 1, 1)
        y = om.conv2d_resample(y, wt, f=self.resample_filter, up=self.up, padding=self.padding, flip_weight=(self.up == 1))
        if noise is not None:
            y = om.fma(y, dcoefs.reshape(n, -1, 1, 1), noise)
        else:
            y = y * dcoefs.reshape(n, -1, 1, 1)
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        y = ops.bias_act(y, self.bias, act=self.activation, gain=self.act_gain * gain, clamp=clamp)
        return _to_nhwc(y)


class ToRGBLayer(torch.nn.Module):
    """networks_stylegan2.py:340-360; forward(x, w, img_prev) also applies the FIR-upsampled skip add (:451-457)."""

    def __init__(self, in_channels, out_channels, w_dim, kernel_size=1, conv_clamp=None, channels_last=False):
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError('ToRGBLayer supports kernel_size=1')
        self.in_channels, self.out_channels, self.w_dim, self.conv_clamp = in_channels, out_channels, w_dim, conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))

    def forward(self, x, w, fused_modconv=True, img_prev=None, x_split=None, bank=None, lidx=-1):
        if not fused_modconv:                               # networks_stylegan2.py:70-79 with demodulate=False, then the skip add of :451-457
            from .shims import ops_modules as om
            styles = self.affine(w) * self.weight_gain
            xn = _to_nchw_view(x)
            y = om.conv2d_resample(xn * styles.reshape(xn.shape[0], -1, 1, 1), self.weight.to(torch.float32), padding=0, flip_weight=True)
            y = ops.bias_act(y, self.bias, clamp=self.conv_clamp)
            if img_prev is not None:
                y = y + ops.upsample2d(_to_nchw_view(img_prev), ops.fir_filter(x.device))
            return _to_nhwc(y)
        styles = self.affine(w) * self.weight_gain if bank is None else None
        return ops.torgb_layer(x, self.weight, styles, self.bias, img_prev, self.conv_clamp, x_split=x_split, bank=bank, lidx=lidx)


class SynthesisBlock(torch.nn.Module):
    """networks_stylegan2.py:365-461, 'skip' architecture.  NHWC in, NHWC out."""

    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=256, use_fp16=False, fp16_channels_last=False,
                 fused_modconv_default=True, **layer_kwargs):
        super().__init__()
        if architecture != 'skip':
            raise NotImplementedError("only the 'skip' architecture used by the EG3D pickles is implemented")
        self.in_channels, self.w_dim, self.resolution, self.img_channels = in_channels, w_dim, resolution, img_channels
        self.is_last, self.architecture, self.use_fp16 = is_last, architecture, use_fp16
        self.fused_modconv_default = fused_modconv_default
        self.register_buffer('resample_filter', ops.setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, conv_clamp=conv_clamp)
        self.num_torgb += 1

    def bank_specs(self, w_base):
        """ops.BankSpec of this block's layers, in execution order; w_base = index of the block's first latent."""
        res, specs, wi = self.resolution, [], w_base
        if self.in_channels != 0:
            specs.append(ops.BankSpec(self.conv0.affine, self.conv0.weight, wi, True, 1.0, res // 2, res // 2, 2))
            wi += 1
        specs.append(ops.BankSpec(self.conv1.affine, self.conv1.weight, wi, True, 1.0, res, res, 1))
        specs.append(ops.BankSpec(self.torgb.affine, self.torgb.weight, wi + 1, False, self.torgb.weight_gain, res, res, 1))
        return specs

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, update_emas=False, x_split=None, return_split=False,
                bank=None, bank_base=0, side=None, lean_out=False, **layer_kwargs):
        """x_split / return_split carry the split-bf16 copies of the activations between consecutive blocks;
        bank / bank_base: pre-computed styles + modulated weights (ops.WeightBank) and this block's first entry."""
        w_iter = iter(ws.unbind(dim=1))
        li = bank_base
        if fused_modconv is None:                           # networks_stylegan2.py:425-428
            fused_modconv = self.fused_modconv_default
        if fused_modconv == 'inference_only':
            fused_modconv = not self.training
        layer_kwargs = dict(layer_kwargs, fused_modconv=bool(fused_modconv))
        if not fused_modconv:
            side = None                                     # the cross-check path is plain stream-ordered torch code
        if self.in_channels == 0:
            x = self.const.to(torch.float32).permute(1, 2, 0).unsqueeze(0).repeat([ws.shape[0], 1, 1, 1]).contiguous()
            sp = None
        else:
            # activations between tensor-core layers travel as their split-bf16 pair only (no fp32 copy): conv0 feeds conv1,
            # conv1 feeds this block's ToRGB and the next block's conv0
            lean0 = bool(fused_modconv) and ops.lean_ok(bank, [li + 1])
            x, sp = self.conv0(x, next(w_iter), x_split=x_split, bank=bank, lidx=li, want_z=not lean0, **layer_kwargs)
            li += 1
        lean1 = bool(fused_modconv) and lean_out and ops.lean_ok(bank, [li + 1, li + 2])
        # conv1's output feeds the next block AND this block's ToRGB: the ToRGB layer gets a handle of its own (x_rgb), so that the two
        # gradients are summed inside conv1's activation backward rather than by an accumulation pass in between
        fork1 = lean1 and ops.CONFIG['fork_grads']
        out = self.conv1(x, next(w_iter), x_split=sp, bank=bank, lidx=li, want_z=not lean1, fork=fork1, **layer_kwargs)
        x, sp = out[0], out[1]
        x_rgb = out[2] if fork1 else x
        if side is None:
            img = self.torgb(x_rgb, next(w_iter), img_prev=img, x_split=sp, bank=bank, lidx=li + 1, fused_modconv=bool(fused_modconv))
        else:
            # ToRGB + skip upsample on the second stream: they depend on this block's x only, the next block's convolutions
            # do not depend on them.  Tensors that cross streams are registered with the allocator (record_stream).
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                img = self.torgb(x_rgb, next(w_iter), img_prev=img, x_split=sp, bank=bank, lidx=li + 1)
            for t in (x, x_rgb) + (tuple(sp) if sp is not None else ()):
                if t is not None:
                    t.record_stream(side)
        if return_split:
            return x, img, sp
        return x, img


class SynthesisNetwork(torch.nn.Module):
    """networks_stylegan2.py:469-525."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=4,
                 **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels, self.num_fp16_res = w_dim, img_resolution, img_channels, num_fp16_res
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        self.num_ws = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res // 2] if res > 4 else 0
            block = SynthesisBlock(in_channels, channels_dict[res], w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=(res == self.img_resolution), use_fp16=(res >= fp16_resolution), **block_kwargs)
            self.num_ws += block.num_conv
            if res == self.img_resolution:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward_nhwc(self, ws, **block_kwargs):
        with on_device(ws):                     # streams, events and allocations below refer to the latents' device
            return self._forward_nhwc(ws, **block_kwargs)

    def _forward_nhwc(self, ws, **block_kwargs):
        block_ws = []
        with ops.prof_range('split_ws'):
            ws = ws.to(torch.float32)
            w_idx = 0
            for res in self.block_resolutions:
                block = getattr(self, f'b{res}')
                block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
                w_idx += block.num_conv
        bank, bases = None, [0] * len(self.block_resolutions)
        if ops.CONFIG['bank']:
            specs, w_idx = [], 0
            for i, res in enumerate(self.block_resolutions):
                block = getattr(self, f'b{res}')
                bases[i] = len(specs)
                specs += block.bank_specs(w_idx)
                w_idx += block.num_conv
            bank = ops.make_bank(ws, specs)
        x = img = sp = None
        side = ops.side_stream(ws.device) if (ops.CONFIG['overlap'] and bank is not None and ws.is_cuda) else None
        if side is not None:
            bank.side = side
            for t in bank.w_hi + bank.w_lo + bank.wmod + bank.styles + [bank.zpool, bank.zf, bank.zb]:
                if t is not None:
                    t.record_stream(side)
        for i, (res, cur_ws) in enumerate(zip(self.block_resolutions, block_ws)):
            x, img, sp = getattr(self, f'b{res}')(x, img, cur_ws, x_split=sp, return_split=True, bank=bank, bank_base=bases[i],
                                                  side=side, lean_out=True, **block_kwargs)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            img.record_stream(torch.cuda.current_stream())
        return img

    def forward(self, ws, c=None, **block_kwargs):
        return _to_nchw_view(self.forward_nhwc(ws, **block_kwargs))


class Generator(torch.nn.Module):
    """The StyleGAN2 backbone (networks_stylegan2.py:529-552)."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)


class SuperresolutionHybrid8X(torch.nn.Module):
    """training/superresolution.py:29-56."""

    def __init__(self, channels, img_resolution, sr_num_fp16_res, sr_antialias, num_fp16_res=4, conv_clamp=None,
                 channel_base=None, channel_max=None, **block_kwargs):
        super().__init__()
        assert img_resolution == 512
        use_fp16 = sr_num_fp16_res > 0
        self.input_resolution = 128
        self.sr_antialias = sr_antialias
        self.block0 = SynthesisBlock(channels, 128, w_dim=512, resolution=256, img_channels=3, is_last=False, use_fp16=use_fp16,
                                     conv_clamp=(256 if use_fp16 else None), **block_kwargs)
        self.block1 = SynthesisBlock(128, 64, w_dim=512, resolution=512, img_channels=3, is_last=True, use_fp16=use_fp16,
                                     conv_clamp=(256 if use_fp16 else None), **block_kwargs)
        self.register_buffer('resample_filter', ops.setup_filter([1, 3, 3, 1]))

    def forward_nhwc(self, rgb, x, ws, **block_kwargs):
        with on_device(ws):
            return self._forward_nhwc(rgb, x, ws, **block_kwargs)

    def _make_bank(self, ws3):
        specs = self.block0.bank_specs(0)
        base1 = len(specs)
        specs += self.block1.bank_specs(0)
        return ops.make_bank(ws3.contiguous(), specs), base1

    def prepare_bank(self, ws):
        """Styles and modulated weights of the module's six layers depend on the latents only: made on the second stream ahead of
        time (the caller runs the renderer meanwhile), their backward runs there too, beside the renderer's.  Returns an opaque
        handle for forward_nhwc(..., prepared=handle), or None when there is no second stream to run on."""
        if not (ops.CONFIG['bank'] and ops.CONFIG['overlap'] and ops.CONFIG['early_sr_bank'] and ws.is_cuda):
            return None
        with on_device(ws):
            side, main = ops.side_stream(ws.device), torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ws3 = ws[:, -1:, :].repeat(1, 3, 1)
                bank, base1 = self._make_bank(ws3)
            ws.record_stream(side)
            return ws3, bank, base1, side

    def _forward_nhwc(self, rgb, x, ws, prepared=None, **block_kwargs):
        if x.shape[1] != self.input_resolution:
            size = (self.input_resolution, self.input_resolution)
            x = _to_nhwc(F.interpolate(_to_nchw_view(x), size=size, mode='bilinear', align_corners=False, antialias=self.sr_antialias))
            rgb = _to_nhwc(F.interpolate(_to_nchw_view(rgb), size=size, mode='bilinear', align_corners=False, antialias=self.sr_antialias))
        bank, base1 = None, 0
        if prepared is not None:
            ws, bank, base1, pside = prepared
            main = torch.cuda.current_stream()
            main.wait_stream(pside)                                 # the bank's kernels ran on the second stream
            for t in bank.w_hi + bank.w_lo + bank.wmod + bank.styles + bank.dcoef + [bank.zpool, bank.zf, bank.zb, ws]:
                if t is not None:
                    t.record_stream(main)                           # allocated on the second stream, read by the layers on this one
        else:
            ws = ws[:, -1:, :].repeat(1, 3, 1)
            if ops.CONFIG['bank']:
                bank, base1 = self._make_bank(ws)
        side = ops.side_stream(ws.device) if (ops.CONFIG['overlap'] and bank is not None and ws.is_cuda) else None
        if side is not None:
            bank.side = side
            side.wait_stream(torch.cuda.current_stream())           # rgb (skip input) and the bank were produced on the main stream
            for t in bank.w_hi + bank.w_lo + bank.wmod + bank.styles + [bank.zpool, bank.zf, bank.zb, rgb]:
                if t is not None:
                    t.record_stream(side)
        x, rgb, sp = self.block0(x, rgb, ws, return_split=True, bank=bank, bank_base=0, side=side, lean_out=True, **block_kwargs)
        x, rgb = self.block1(x, rgb, ws, x_split=sp, bank=bank, bank_base=base1, side=side, lean_out=True, **block_kwargs)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            rgb.record_stream(torch.cuda.current_stream())
        return rgb

    def forward(self, rgb, x, ws, **block_kwargs):
        return _to_nchw_view(self.forward_nhwc(_to_nhwc(rgb), _to_nhwc(x), ws, **block_kwargs))


class OSGDecoder(torch.nn.Module):
    """training/triplane.py:113-136.  The renderer fuses this MLP into the sampling kernel; forward() (features in,
    as in the reference) is kept for callers that hold pre-sampled features."""

    def __init__(self, n_features, options):
        super().__init__()
        self.hidden_dim = 64
        self.net = torch.nn.Sequential(
            FullyConnectedLayer(n_features, self.hidden_dim, lr_multiplier=options['decoder_lr_mul']),
            torch.nn.Softplus(),
            FullyConnectedLayer(self.hidden_dim, 1 + options['decoder_output_dim'], lr_multiplier=options['decoder_lr_mul']))

    def forward(self, sampled_features, ray_directions):
        x = sampled_features.mean(1)
        N, M, C = x.shape
        x = self.net(x.view(N * M, C)).view(N, M, -1)
        rgb = torch.sigmoid(x[..., 1:]) * (1 + 2 * 0.001) - 0.001
        return {'rgb': rgb, 'sigma': x[..., 0:1]}


class RaySampler(torch.nn.Module):
    """training/volumetric_rendering/ray_sampler.py:24-73 (device follows the inputs; the reference hard-codes .cuda())."""

    def forward(self, cam2world_matrix, intrinsics, resolution, need_cam_space=False):
        if cam2world_matrix.is_cuda and not need_cam_space and not intrinsics.requires_grad:
            return ops.ray_sampler(cam2world_matrix, intrinsics, resolution)      # one kernel instead of ~15 small launches
        return self.forward_torch(cam2world_matrix, intrinsics, resolution, need_cam_space)

    def forward_torch(self, cam2world_matrix, intrinsics, resolution, need_cam_space=False):
        """The same statements as device-agnostic torch ops (camera-space rays, intrinsics that require a gradient)."""
        N, M = cam2world_matrix.shape[0], resolution ** 2
        dev = cam2world_matrix.device
        cam_locs_world = cam2world_matrix[:, :3, 3]
        fx, fy = intrinsics[:, 0, 0], intrinsics[:, 1, 1]
        cx, cy, sk = intrinsics[:, 0, 2], intrinsics[:, 1, 2], intrinsics[:, 0, 1]
        ar = torch.arange(resolution, dtype=torch.float32, device=dev)
        uv = torch.stack(torch.meshgrid(ar, ar, indexing='ij')) * (1. / resolution) + (0.5 / resolution)
        uv = uv.flip(0).reshape(2, -1).transpose(1, 0).unsqueeze(0).repeat(N, 1, 1)
        x_cam, y_cam = uv[:, :, 0].view(N, -1), uv[:, :, 1].view(N, -1)
        z_cam = torch.ones((N, M), device=dev)
        x_lift = (x_cam - cx.unsqueeze(-1) + cy.unsqueeze(-1) * sk.unsqueeze(-1) / fy.unsqueeze(-1)
                  - sk.unsqueeze(-1) * y_cam / fy.unsqueeze(-1)) / fx.unsqueeze(-1) * z_cam
        y_lift = (y_cam - cy.unsqueeze(-1)) / fy.unsqueeze(-1) * z_cam
        cam_rel_points = torch.stack((x_lift, y_lift, z_cam, torch.ones_like(z_cam)), dim=-1)
        if need_cam_space:                                   # ray_sampler.py:61-70
            return torch.zeros_like(cam_locs_world), F.normalize(cam_rel_points[:, :, :3], dim=2), uv
        world_rel_points = torch.bmm(cam2world_matrix, cam_rel_points.permute(0, 2, 1)).permute(0, 2, 1)[:, :, :3]
        ray_dirs = F.normalize(world_rel_points - cam_locs_world[:, None, :], dim=2)
        ray_origins = cam_locs_world.unsqueeze(1).repeat(1, ray_dirs.shape[1], 1)
        return ray_origins, ray_dirs

    def calculate_xyz_of_depth(self, ray_origin, ray_dirs, depth):
        """ray_sampler.py:75-93: homogeneous world xyz [4, res*res] of a rendered depth map [1,res,res] / [1,1,res,res]
        along rays [1,res*res,3] (or [3,res,res]); used by the warping loss (warping_loss.py:21)."""
        res = depth.shape[-1]
        if ray_origin.shape[0] == 1 and ray_origin.shape[1] == res ** 2:
            ray_origin = ray_origin.squeeze(0).reshape(res, res, 3).permute(2, 0, 1)
        if ray_dirs.shape[0] == 1 and ray_dirs.shape[1] == res ** 2:
            ray_dirs = ray_dirs.squeeze(0).reshape(res, res, 3).permute(2, 0, 1)
        xyz = ray_origin + ray_dirs * depth.squeeze(0)
        ones = torch.ones(1, xyz.shape[1], xyz.shape[2], device=ray_origin.device)
        return torch.cat([xyz, ones], dim=0).reshape(4, res * res)


class ImportanceRenderer(torch.nn.Module):
    """training/volumetric_rendering/renderer.py:136-308 with the sampling / decoding / marching fused into CUDA kernels."""

    def __init__(self):
        super().__init__()
        self.fixed_noise = None        # optional (u_strat [N,M,S,1], u_imp [N*M,S_imp]) replacing the two torch.rand draws

    @staticmethod
    def _planes_nhwc(planes):
        if planes.ndim == 5:           # [N,3,C,H,W] -> [N,H,W,3*C]; a view when produced by SynthesisNetwork.forward
            n, p, c, h, w = planes.shape
            return planes.permute(0, 3, 4, 1, 2).reshape(n, h, w, p * c).contiguous()
        return planes                   # already [N,H,W,96]

    @staticmethod
    def ray_limits_box(rays_o, rays_d, box_side_length):
        """math_utils.py:46-98 (get_ray_limits_box): slab test of every ray against the cube of side box_side_length about the
        origin; returns (t_near, t_far) [N,M,1], (-1, -2) for rays that miss.  No gradient (the reference detaches the rays)."""
        o, d = rays_o.detach().reshape(-1, 3), rays_d.detach().reshape(-1, 3)
        half = box_side_length / 2
        inv = 1 / d
        neg = inv < 0
        t_in = (torch.where(neg, half, -half) - o) * inv          # entry / exit distance through each pair of faces
        t_out = (torch.where(neg, -half, half) - o) * inv
        tmin, tmax = t_in[:, 0], t_out[:, 0]
        hit = ~((tmin > t_out[:, 1]) | (t_in[:, 1] > tmax))
        tmin, tmax = torch.maximum(tmin, t_in[:, 1]), torch.minimum(tmax, t_out[:, 1])
        hit &= ~((tmin > t_out[:, 2]) | (t_in[:, 2] > tmax))
        tmin, tmax = torch.maximum(tmin, t_in[:, 2]), torch.minimum(tmax, t_out[:, 2])
        tmin = torch.where(hit, tmin, -torch.ones_like(tmin))
        tmax = torch.where(hit, tmax, -2 * torch.ones_like(tmax))
        return tmin.reshape(*rays_o.shape[:-1], 1), tmax.reshape(*rays_o.shape[:-1], 1)

    def forward(self, planes, decoder, ray_origins, ray_directions, rendering_options):
        assert rendering_options.get('clamp_mode', 'softplus') == 'softplus'
        pl = self._planes_nhwc(planes)
        dev = pl.device
        N, M, _ = ray_origins.shape
        S = int(rendering_options['depth_resolution'])
        S2 = int(rendering_options['depth_resolution_importance'])
        auto = rendering_options['ray_start'] == rendering_options['ray_end'] == 'auto'
        disparity = bool(rendering_options.get('disparity_space_sampling', False))
        if self.fixed_noise is not None:
            # a (u_strat, u_imp) pair used for every call, or a list of pairs consumed one per call (tests replaying several renders)
            u_strat, u_imp = self.fixed_noise.pop(0) if isinstance(self.fixed_noise, list) else self.fixed_noise
            u_strat = u_strat.to(dev)
            u_imp = u_imp.to(dev) if S2 > 0 else None
        else:
            u_strat = torch.rand([N, M, S, 1], device=dev)                        # renderer.py:245
            u_imp = torch.rand([N * M, S2], device=dev) if S2 > 0 else None       # renderer.py:292
        t_base, delta, t_coarse = None, 0.0, None
        if auto:                                                                  # renderer.py:146-152, sample_stratified :238-242
            ray_start, ray_end = self.ray_limits_box(ray_origins, ray_directions, rendering_options['box_warp'])
            ok = ray_end > ray_start
            lo, hi = torch.where(ok, ray_start, float('inf')).min(), torch.where(ok, ray_start, -float('inf')).max()
            any_ok = ok.any()                                                     # stays on the device: no host sync (the reference calls .item())
            ray_start = torch.where(ok | ~any_ok, ray_start, lo)
            ray_end = torch.where(ok | ~any_ok, ray_end, hi)
            steps = (torch.arange(S, dtype=torch.float32, device=dev) / (S - 1)).reshape(1, 1, S, 1)
            if disparity:
                t = steps + u_strat * (1.0 / (S - 1))
                t_coarse = 1. / (1. / ray_start.unsqueeze(2) * (1. - t) + 1. / ray_end.unsqueeze(2) * t)
            else:
                t_coarse = ray_start.unsqueeze(2) + steps * (ray_end - ray_start).unsqueeze(2) \
                    + u_strat * ((ray_end - ray_start) / (S - 1)).unsqueeze(-1)
        elif disparity:                                                           # renderer.py:230-237
            ray_start, ray_end = float(rendering_options['ray_start']), float(rendering_options['ray_end'])
            t = torch.linspace(0, 1, S, device=dev).reshape(1, 1, S, 1) + u_strat * (1.0 / (S - 1))
            t_coarse = 1. / (1. / ray_start * (1. - t) + 1. / ray_end * t)
        else:
            ray_start, ray_end = float(rendering_options['ray_start']), float(rendering_options['ray_end'])
            t_base = torch.linspace(ray_start, ray_end, S, device=dev)
            delta = (ray_end - ray_start) / (S - 1)
        return ops.render(pl, decoder, ray_origins, ray_directions, rendering_options['box_warp'], t_base, delta, u_strat,
                          u_imp, white_back=rendering_options.get('white_back', False),
                          density_noise=rendering_options.get('density_noise', 0), t_coarse=t_coarse)

    def run_model(self, planes, decoder, sample_coordinates, sample_directions, options):
        pl = self._planes_nhwc(planes)
        rgb, sigma = ops.run_model_nhwc(pl, decoder, sample_coordinates, options['box_warp'])
        if options.get('density_noise', 0) > 0:
            sigma = sigma + torch.randn_like(sigma) * options['density_noise']
        return {'rgb': rgb, 'sigma': sigma}


class TriPlaneGenerator(torch.nn.Module):
    """training/triplane.py:19-110."""

    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, sr_num_fp16_res=0, mapping_kwargs={},
                 rendering_kwargs={}, sr_kwargs={}, **synthesis_kwargs):
        super().__init__()
        self.init_args = ()
        self.init_kwargs = copy.deepcopy(dict(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, img_resolution=img_resolution,
                                              img_channels=img_channels, sr_num_fp16_res=sr_num_fp16_res,
                                              mapping_kwargs=mapping_kwargs, rendering_kwargs=rendering_kwargs,
                                              sr_kwargs=sr_kwargs, **synthesis_kwargs))
        self.z_dim, self.c_dim, self.w_dim = z_dim, c_dim, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.renderer = ImportanceRenderer()
        self.ray_sampler = RaySampler()
        self.backbone = Generator(z_dim, c_dim, w_dim, img_resolution=256, img_channels=32 * 3, mapping_kwargs=mapping_kwargs,
                                  **synthesis_kwargs)
        sr_name = rendering_kwargs.get('superresolution_module', 'training.superresolution.SuperresolutionHybrid8X')
        if not sr_name.endswith('SuperresolutionHybrid8X'):
            raise NotImplementedError(f'superresolution module {sr_name} is not implemented (512-128 EG3D models use Hybrid8X)')
        self.superresolution = SuperresolutionHybrid8X(channels=32, img_resolution=img_resolution, sr_num_fp16_res=sr_num_fp16_res,
                                                       sr_antialias=rendering_kwargs['sr_antialias'], **sr_kwargs)
        self.decoder = OSGDecoder(32, {'decoder_lr_mul': rendering_kwargs.get('decoder_lr_mul', 1), 'decoder_output_dim': 32})
        self.neural_rendering_resolution = 64
        self.rendering_kwargs = rendering_kwargs
        self._last_planes = None

    def mapping(self, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False):
        if self.rendering_kwargs['c_gen_conditioning_zero']:
            c = torch.zeros_like(c)
        return self.backbone.mapping(z, c * self.rendering_kwargs.get('c_scale', 0), truncation_psi=truncation_psi,
                                     truncation_cutoff=truncation_cutoff, update_emas=update_emas)

    def synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                  use_cached_backbone=False, **synthesis_kwargs):
        with on_device(ws):
            return self._synthesis(ws, c, neural_rendering_resolution, update_emas, cache_backbone, use_cached_backbone,
                                   **synthesis_kwargs)

    def _synthesis(self, ws, c, neural_rendering_resolution=None, update_emas=False, cache_backbone=False,
                   use_cached_backbone=False, **synthesis_kwargs):
        cam2world_matrix = c[:, :16].view(-1, 4, 4)
        intrinsics = c[:, 16:25].view(-1, 3, 3)
        if neural_rendering_resolution is None:
            neural_rendering_resolution = self.neural_rendering_resolution
        else:
            self.neural_rendering_resolution = neural_rendering_resolution
        ray_origins, ray_directions = self.ray_sampler(cam2world_matrix, intrinsics, neural_rendering_resolution)
        N, M, _ = ray_origins.shape
        if use_cached_backbone and self._last_planes is not None:
            planes = self._last_planes
        else:
            planes = self.backbone.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)     # channels_last [N,96,256,256]
        if cache_backbone:
            self._last_planes = planes
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        # the super-resolution module's styles and modulated weights need the latents only: started on the second stream here, so that
        # they (and, in the backward pass, their gradients) run beside the renderer instead of in front of the first SR convolution
        prepared = None
        if synthesis_kwargs.get('fused_modconv', True) is not False and hasattr(self.superresolution, 'prepare_bank'):
            prepared = self.superresolution.prepare_bank(ws)
        feature_samples, depth_samples, _ = self.renderer(planes, self.decoder, ray_origins, ray_directions, self.rendering_kwargs)
        H = W = self.neural_rendering_resolution
        feature_nhwc = feature_samples.reshape(N, H, W, feature_samples.shape[-1])                 # [N,M,32] is already NHWC
        depth_image = depth_samples.permute(0, 2, 1).reshape(N, 1, H, W)
        feature_image = _to_nchw_view(feature_nhwc)
        rgb_image = feature_image[:, :3]
        sr_kwargs = {k: v for k, v in synthesis_kwargs.items() if k != 'noise_mode'}
        if prepared is not None:
            sr_kwargs['prepared'] = prepared
        sr_image = self.superresolution(rgb_image, feature_image, ws, noise_mode=self.rendering_kwargs['superresolution_noise_mode'],
                                        **sr_kwargs)
        return {'image': sr_image, 'image_raw': rgb_image, 'image_depth': depth_image}

    def sample(self, coordinates, directions, z, c, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.sample_mixed(coordinates, directions, ws, update_emas=update_emas, **synthesis_kwargs)

    def sample_mixed(self, coordinates, directions, ws, truncation_psi=1, truncation_cutoff=None, update_emas=False, **synthesis_kwargs):
        planes = self.backbone.synthesis(ws, update_emas=update_emas, **synthesis_kwargs)
        planes = planes.view(len(planes), 3, 32, planes.shape[-2], planes.shape[-1])
        return self.renderer.run_model(planes, self.decoder, coordinates, directions, self.rendering_kwargs)

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, neural_rendering_resolution=None, update_emas=False,
                cache_backbone=False, use_cached_backbone=False, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff, update_emas=update_emas)
        return self.synthesis(ws, c, update_emas=update_emas, neural_rendering_resolution=neural_rendering_resolution,
                              cache_backbone=cache_backbone, use_cached_backbone=use_cached_backbone, **synthesis_kwargs)
