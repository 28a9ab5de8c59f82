"""b200eg3d -- Blackwell-native EG3D tri-plane generator forward+backward (the 3DGAN-Inversion hot path).

Public surface (mirrors the reference's generator API, see generator.py):
    TriPlaneGenerator, ImportanceRenderer, RaySampler, OSGDecoder, SuperresolutionHybrid8X
    ops.bias_act / ops.upfirdn2d / ops.upsample2d / ops.setup_filter      (torch_utils.ops equivalents)
    seam.convert_generator / seam.load_old_G                               (drop-in replacement of utils/models_utils.py:21-25)
    projector.calc_warping_loss / noise_regularizer / normalize_noise_     (w_projector.py:145-270 caller-side pieces, fused)
    geometry.create_samples / query_sigma / density_grid                    (single_id_coach.py:118-188 shape extraction, planes cached)
    losses.pti_loss / losses.compute_tv_norm                               (base_coach.py:101-126,294-305 as fused reductions)
"""
from . import ops  # noqa: F401
from .generator import (FullyConnectedLayer, Generator, ImportanceRenderer, MappingNetwork, OSGDecoder, RaySampler,  # noqa: F401
                        SuperresolutionHybrid8X, SynthesisBlock, SynthesisLayer, SynthesisNetwork, ToRGBLayer,
                        TriPlaneGenerator)
from . import seam  # noqa: F401
from . import losses  # noqa: F401
from . import projector  # noqa: F401
from . import geometry  # noqa: F401
from . import optim  # noqa: F401

__version__ = '0.1.0'
