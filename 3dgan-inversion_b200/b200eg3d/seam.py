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


def _plain(obj):
    """Recursively turn dict subclasses (dnnlib.EasyDict), tuples and numpy scalars into plain dict / list / python scalars,
    so that the checkpoint holds no class references and loads with torch.load(weights_only=True)."""
    if isinstance(obj, dict):
        return {str(k): _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu()
    if hasattr(obj, 'item') and not isinstance(obj, (str, bytes)):
        try:
            return obj.item()
        except Exception:
            pass
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    raise TypeError(f'cannot store {type(obj).__name__} in a tuned-generator checkpoint')


def save_tuned_G(G, path):
    """Checkpoint of a (PTI-tuned) generator: constructor arguments + state_dict + the attributes callers set after loading.
    The reference never saves the tuned generator (single_id_coach.py:117 is commented out); this is the `state_dict` route
    SURVEY.md 8 f4 asks for.  Plain torch.save of tensors and python scalars -- no pickled code."""
    torch.save({'format': 'b200eg3d.tuned_G.v1', 'init_args': _plain(list(G.init_args)), 'init_kwargs': _plain(dict(G.init_kwargs)),
                'state_dict': {k: v.detach().cpu() for k, v in G.state_dict().items()},
                'neural_rendering_resolution': int(G.neural_rendering_resolution),
                'rendering_kwargs': _plain(dict(G.rendering_kwargs))}, path)


def load_tuned_G(path, device='cuda'):
    ck = torch.load(path, map_location='cpu', weights_only=True)         # tensors + plain containers only: no pickled code runs
    if ck.get('format') != 'b200eg3d.tuned_G.v1':
        raise ValueError(f'{path} is not a b200eg3d tuned-generator checkpoint')
    G = TriPlaneGenerator(*ck['init_args'], **ck['init_kwargs']).eval().requires_grad_(False)
    G.init_kwargs['rendering_kwargs'] = ck['rendering_kwargs']
    G.load_state_dict(ck['state_dict'], strict=True)
    G.neural_rendering_resolution = ck['neural_rendering_resolution']
    G.rendering_kwargs = ck['rendering_kwargs']
    return G.to(device)
