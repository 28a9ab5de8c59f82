"""The drop-in seam: turn a reference (un-pickled) TriPlaneGenerator into the B200-native one.

The reference loads its generator with pickle and runs the class source embedded in the pickle
(utils/models_utils.py:21-25, torch_utils/persistence.py:181-229).  The supported way to swap implementations is
the reference's own `reload_modules` pattern (gen_samples.py:146-152): rebuild from init_args/init_kwargs and copy
every parameter and buffer by name.
"""
import copy
import pickle

import torch

from .generator import TriPlaneGenerator


def copy_params_and_buffers(src_module, dst_module, require_all=False):
    """torch_utils/misc.py:157-164."""
    src = dict(list(src_module.named_parameters()) + list(src_module.named_buffers()))
    with torch.no_grad():
        for name, tensor in list(dst_module.named_parameters()) + list(dst_module.named_buffers()):
            if name not in src:
                if require_all:
                    raise KeyError(f'{name} is missing in the source generator')
                continue
            tensor.copy_(src[name].detach().to(tensor.dtype))


def convert_generator(G_ref, device='cuda'):
    """G_ref: reference TriPlaneGenerator (anything exposing init_args / init_kwargs, named parameters and buffers)."""
    kwargs = copy.deepcopy(dict(G_ref.init_kwargs))
    G = TriPlaneGenerator(*G_ref.init_args, **kwargs).eval().requires_grad_(False).to(device)
    copy_params_and_buffers(G_ref, G, require_all=True)
    G.neural_rendering_resolution = G_ref.neural_rendering_resolution
    G.rendering_kwargs = G_ref.rendering_kwargs
    return G


def load_old_G(path, device='cuda'):
    """Replacement for utils/models_utils.py:21-25: un-pickle with the reference on sys.path, then convert."""
    with open(path, 'rb') as f:
        G_ref = pickle.load(f)['G_ema'].eval().float()
    return convert_generator(G_ref, device=device).float()
