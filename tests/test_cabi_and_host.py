"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200eg3d.h declares (no compute calls),
the ctypes table covers the header, the module tree mirrors the reference's parameter names, the product refuses to run
without CUDA, and the N>1 host logic works under gloo with world_size 2."""
import ctypes
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'b200eg3d.h')


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(b200_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from b200eg3d import _lib
    from b200eg3d import build
    build.build()                                   # nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/b200eg3d.h but not exported'
    assert lib.b200_version() >= 100


def test_ctypes_table_matches_header():
    from b200eg3d import _lib
    syms = set(declared_symbols())
    bound = set(_lib.SIGNATURES) | {'b200_version', 'b200_last_error', 'b200_set_pdl', 'b200_conv_tc_supported', 'b200_triplane_bwd_workspace_bytes',
                                        'b200_noise_pyramid_work_floats', 'b200_conv_tc_act_fusable'}
    assert syms == bound, (sorted(syms - bound), sorted(bound - syms))
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():      # argument counts agree with the prototypes
        m = re.search(r'\b' + name + r'\s*\((.*?)\)\s*;', src, flags=re.S)
        assert m, name
        assert len([a for a in m.group(1).split(',') if a.strip()]) == len(args), name


def test_shape_support_query_needs_no_gpu():
    from b200eg3d import _lib
    lib = _lib.load()
    assert lib.b200_conv_tc_supported(0, 64, 64, 512, 512, 3, 1) == 1
    assert lib.b200_conv_tc_supported(0, 64, 64, 12, 20, 3, 1) == 0        # channel count not a multiple of 8 -> fp32 path
    assert lib.b200_conv_tc_supported(1, 64, 64, 64, 3, 1, 1) == 0
    assert lib.b200_conv_tc_supported(0, 8, 8, 64, 64, 5, 1) == 0
    prev = lib.b200_set_pdl(0)
    assert lib.b200_set_pdl(prev) == 0 and lib.b200_set_pdl(prev) == prev


def test_module_tree_mirrors_reference_names():
    import b200eg3d
    import synth_params as sp
    from golden_util import param_shapes
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=sp.rendering_kwargs(), **sp.G_KWARGS_TINY)
    mine = {k: tuple(v.shape) for k, v in list(G.named_parameters()) + list(G.named_buffers())}
    for k, shp in param_shapes(sp.G_KWARGS_TINY).items():
        assert mine.get(k) == tuple(shp), (k, shp, mine.get(k))
    assert [n for n, _ in G.named_children()] == ['renderer', 'ray_sampler', 'backbone', 'superresolution', 'decoder']
    assert G.backbone.num_ws == 14 and G.init_kwargs['img_resolution'] == 512
    noise = [n for n, _ in G.backbone.synthesis.named_buffers() if 'noise_const' in n]       # w_projector.py:103-104
    assert len(noise) == 13
    for attr in ('synthesis', 'mapping', 'sample', 'sample_mixed', 'forward'):
        assert callable(getattr(G, attr))


def test_no_cpu_fallback():
    import b200eg3d
    import synth_params as sp
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=sp.rendering_kwargs(), **sp.G_KWARGS_TINY).eval()
    with pytest.raises(RuntimeError, match='no CPU path'):
        G.synthesis(sp.latent_ws(1), sp.camera())


def test_seam_copies_by_name():
    import b200eg3d
    import synth_params as sp
    src = b200eg3d.TriPlaneGenerator(rendering_kwargs=sp.rendering_kwargs(), **sp.G_KWARGS_TINY)
    sp.fill_params_(dict(list(src.named_parameters()) + list(src.named_buffers())), 3)
    src.neural_rendering_resolution = 96
    dst = b200eg3d.seam.convert_generator(src, device='cpu')
    for (n1, p1), (n2, p2) in zip(sorted(src.state_dict().items()), sorted(dst.state_dict().items())):
        assert n1 == n2 and torch.equal(p1, p2)
    assert dst.neural_rendering_resolution == 96 and dst.rendering_kwargs is src.rendering_kwargs


def test_shard_indices():
    from b200eg3d.shard import shard_indices
    assert shard_indices(10, 0, 4) == [0, 4, 8] and shard_indices(10, 3, 4) == [3, 7] and shard_indices(2, 3, 4) == []
    assert sorted(sum((shard_indices(17, r, 8) for r in range(8)), [])) == list(range(17))
    with pytest.raises(ValueError):
        shard_indices(4, 4, 4)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from b200eg3d.shard import reduce_run_stats, shard_indices
    mine = shard_indices(5, rank, world)
    res = reduce_run_stats(steps=len(mine) * 10, elapsed_ms=100.0 + 50.0 * rank, loss_sum=float(sum(mine)))
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def test_gloo_world_size_2_reduction():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    res = out.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert res == (50, 150.0, 10.0)


def test_tuned_generator_checkpoint_roundtrip(tmp_path):
    import b200eg3d
    import synth_params as sp
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=sp.rendering_kwargs(), **sp.G_KWARGS_TINY)
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 5)
    G.neural_rendering_resolution = 128
    path = str(tmp_path / 'tuned.pt')
    b200eg3d.seam.save_tuned_G(G, path)
    G2 = b200eg3d.seam.load_tuned_G(path, device='cpu')
    assert G2.neural_rendering_resolution == 128 and G2.rendering_kwargs == G.rendering_kwargs
    for (n1, p1), (n2, p2) in zip(sorted(G.state_dict().items()), sorted(G2.state_dict().items())):
        assert n1 == n2 and torch.equal(p1, p2)
    torch.save({'format': 'other'}, path)
    with pytest.raises(ValueError):
        b200eg3d.seam.load_tuned_G(path, device='cpu')
