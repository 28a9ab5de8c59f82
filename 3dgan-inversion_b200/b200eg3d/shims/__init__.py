"""Op-level seam (second row): the reference's operator modules and renderer classes, by their own import
paths, backed by the b200eg3d CUDA library.

A generator that was un-pickled WITHOUT being rebuilt runs the class source embedded in the pickle (torch_utils/persistence.py);
that source looks its operators up by import path at call time:

    from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, conv2d_gradfix, fma      (networks_stylegan2.py:17-21)
    from training.volumetric_rendering.ray_marcher import MipRayMarcher2                      (renderer.py:20)
    from training.volumetric_rendering.renderer import ImportanceRenderer                      (triplane.py:14)
    from training.volumetric_rendering.ray_sampler import RaySampler                           (triplane.py:15)

`install()` makes those paths resolve to the look-alikes in this package: existing modules (the real reference on sys.path) get
their public callables replaced, missing ones are created.  `uninstall()` restores everything.  The primary seam remains
seam.convert_generator (the whole module tree is swapped); this one exists for callers that cannot rebuild.
"""
import importlib
import sys
import types

from . import ops_modules, rendering

# import path -> {attribute: replacement}
TABLE = {
    'torch_utils.ops.bias_act': {'bias_act': ops_modules.bias_act, 'activation_funcs': ops_modules.activation_funcs},
    'torch_utils.ops.upfirdn2d': {'upfirdn2d': ops_modules.upfirdn2d, 'setup_filter': ops_modules.setup_filter, 'filter2d': ops_modules.filter2d,
                                  'upsample2d': ops_modules.upsample2d, 'downsample2d': ops_modules.downsample2d},
    'torch_utils.ops.conv2d_resample': {'conv2d_resample': ops_modules.conv2d_resample},
    'torch_utils.ops.conv2d_gradfix': {'conv2d': ops_modules.conv2d, 'conv_transpose2d': ops_modules.conv_transpose2d,
                                       'no_weight_gradients': ops_modules.no_weight_gradients},
    'torch_utils.ops.fma': {'fma': ops_modules.fma},
    'training.volumetric_rendering.ray_marcher': {'MipRayMarcher2': rendering.MipRayMarcher2},
    'training.volumetric_rendering.renderer': {'ImportanceRenderer': rendering.ImportanceRenderer, 'sample_from_planes': rendering.sample_from_planes},
    'training.volumetric_rendering.ray_sampler': {'RaySampler': rendering.RaySampler},
}
_SAVED = []        # (module, attribute, previous value or _MISSING)
_CREATED = []      # module names this package put into sys.modules
_MISSING = object()


def _get_or_create(path):
    try:
        return importlib.import_module(path)
    except Exception:
        parts = path.split('.')
        for i in range(1, len(parts) + 1):
            name = '.'.join(parts[:i])
            if name not in sys.modules:
                mod = types.ModuleType(name)
                mod.__path__ = []                      # behaves like a package for the import machinery
                sys.modules[name] = mod
                _CREATED.append(name)
                if i > 1:
                    setattr(sys.modules['.'.join(parts[:i - 1])], parts[i - 1], mod)
        return sys.modules[path]


def install():
    """Route the reference's operator import paths to b200eg3d.  Idempotent.  Returns the list of patched module paths."""
    if _SAVED or _CREATED:
        return sorted(TABLE)
    for path, attrs in TABLE.items():
        mod = _get_or_create(path)
        for name, repl in attrs.items():
            _SAVED.append((mod, name, getattr(mod, name, _MISSING)))
            setattr(mod, name, repl)
    return sorted(TABLE)


def uninstall():
    while _SAVED:
        mod, name, old = _SAVED.pop()
        if old is _MISSING:
            if hasattr(mod, name):
                delattr(mod, name)
        else:
            setattr(mod, name, old)
    while _CREATED:
        sys.modules.pop(_CREATED.pop(), None)
