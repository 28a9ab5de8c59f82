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
    assert lib.b200_version() == _lib.EXPECTED_VERSION


def test_ctypes_table_matches_header():
    from b200eg3d import _lib
    syms = set(declared_symbols())
    bound = set(_lib.SIGNATURES) | {'b200_version', 'b200_last_error', 'b200_set_pdl', 'b200_set_mlp_passes', 'b200_set_triplane_impl', 'b200_conv_tc_supported', 'b200_triplane_bwd_workspace_bytes', 'b200_triplane_fsave_bytes',
                                        'b200_noise_pyramid_work_floats', 'b200_conv_tc_act_fusable', 'b200_set_conv_pair', 'b200_conv1x1_thin_supported', 'b200_conv_tc_ksplit', 'b200_conv1x1_fwd_thin_supported'}
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
    prev = lib.b200_set_conv_pair(0)
    assert lib.b200_set_conv_pair(prev) == 0 and lib.b200_set_conv_pair(prev) == prev
    assert lib.b200_conv1x1_thin_supported(64, 3) == 1 and lib.b200_conv1x1_thin_supported(64, 96) == 0
    # the 4x4 .. 32x32 blocks split K (their outputs must start at zero), the large layers do not; unsupported shapes report 1
    assert lib.b200_conv_tc_ksplit(0, 1, 16, 16, 512, 512, 3, 1) > 1 and lib.b200_conv_tc_ksplit(1, 1, 16, 16, 512, 512, 3, 1) > 1
    assert lib.b200_conv_tc_ksplit(0, 1, 8, 8, 512, 512, 3, 2) > 1
    assert lib.b200_conv_tc_ksplit(0, 1, 256, 256, 128, 128, 3, 1) == 1 and lib.b200_conv_tc_ksplit(1, 1, 256, 256, 128, 128, 3, 1) == 1
    assert lib.b200_conv_tc_ksplit(0, 1, 16, 16, 12, 20, 3, 1) == 1
    # two-addend activation backward: channel vectors of a pixel inside one warp, a power-of-two count when the noise gradient needs
    # the shuffle sum, 32-bit element indices
    q = lib.b200_layer_act_bwd_sum2_supported
    assert q(1, 256 * 256, 128, 1, 0) == 1 and q(1, 64 * 64, 512, 1, 0) == 1 and q(4, 16, 64, 0, 16) == 1
    assert q(1, 64, 20, 1, 0) == 0 and q(1, 64, 20, 0, 0) == 1 and q(1, 64, 3, 0, 0) == 0 and q(1, 64, 1024, 0, 0) == 0
    assert q(64, 512 * 512, 512, 0, 0) == 0                                   # 2^33 elements: the generic kernels' 64-bit indices


@pytest.mark.parametrize('arch', ['tiny', 'full'])
def test_module_tree_mirrors_reference_names(arch):
    """Names and shapes of EVERY parameter and buffer (incl. backbone.mapping.*) equal the manifest recorded from the real
    reference class by oracle/make_goldens.py; the hand-written mirror used by the oracle agrees with it too."""
    import b200eg3d
    import synth_params as sp
    from golden_util import manifest, param_shapes
    man = manifest(arch)
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=sp.rendering_kwargs(), **sp.G_KWARGS[arch])
    assert {k: list(v.shape) for k, v in G.named_parameters()} == man['parameters']
    assert {k: list(v.shape) for k, v in G.named_buffers()} == man['buffers']
    assert G.backbone.num_ws == man['num_ws']
    ref = dict(man['parameters'], **man['buffers'])
    for k, shp in param_shapes(sp.G_KWARGS[arch]).items():
        assert ref.get(k) == list(shp), (k, shp, ref.get(k))
    if arch == 'full':
        return
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


class _ManifestModule(torch.nn.Module):
    """Stand-in for an un-pickled reference generator: exposes exactly the manifest's named parameters / buffers plus
    init_args / init_kwargs / rendering_kwargs -- everything seam.convert_generator reads (gen_samples.py:146-152)."""

    def __init__(self, man, gk, rk):
        super().__init__()
        self.init_args, self.init_kwargs = (), dict(gk, rendering_kwargs=rk)
        self.rendering_kwargs, self.neural_rendering_resolution = rk, 96
        for kind, reg in (('parameters', self._reg_p), ('buffers', self._reg_b)):
            for name, shape in man[kind].items():
                reg(name, torch.zeros(shape))

    def _leaf(self, name):
        mod = self
        *path, leaf = name.split('.')
        for p in path:
            if not hasattr(mod, p):
                mod.add_module(p, torch.nn.Module())
            mod = getattr(mod, p)
        return mod, leaf

    def _reg_p(self, name, t):
        mod, leaf = self._leaf(name)
        mod.register_parameter(leaf, torch.nn.Parameter(t))

    def _reg_b(self, name, t):
        mod, leaf = self._leaf(name)
        mod.register_buffer(leaf, t)


def test_seam_copies_by_name():
    """convert_generator against a module carrying the REAL reference's name/shape manifest (incl. backbone.mapping.*)."""
    import b200eg3d
    import synth_params as sp
    from golden_util import manifest
    rk = sp.rendering_kwargs()
    src = _ManifestModule(manifest('tiny'), sp.G_KWARGS_TINY, rk)
    sp.fill_params_(dict(list(src.named_parameters()) + list(src.named_buffers())), 3)
    with torch.no_grad():
        src.backbone.mapping.w_avg.normal_()
    dst = b200eg3d.seam.convert_generator(src, device='cpu')
    a, b = src.state_dict(), dst.state_dict()
    assert sorted(a) == sorted(b)
    for n in a:
        assert torch.equal(a[n], b[n]), n
    assert dst.neural_rendering_resolution == 96 and dst.rendering_kwargs is src.rendering_kwargs
    # a source lacking a tensor the destination has must fail loudly (misc.copy_params_and_buffers(require_all=True))
    man = manifest('tiny')
    man['parameters'].pop('decoder.net.2.bias')
    with pytest.raises(KeyError):
        b200eg3d.seam.convert_generator(_ManifestModule(man, sp.G_KWARGS_TINY, rk), device='cpu')


@pytest.mark.skipif(not os.path.isdir('/root/reference/training'), reason='the reference tree only exists in the build container')
def test_seam_converts_the_real_reference_class():
    """Build container only: the unmodified reference TriPlaneGenerator -> b200eg3d, misc.copy_params_and_buffers semantics."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path[:0] = [%r, %r, '/root/reference'];"
        "import synth_params as sp; import b200eg3d;"
        "from training.triplane import TriPlaneGenerator as Ref;"
        "import dnnlib;"
        "rk = dnnlib.EasyDict(sp.rendering_kwargs());"
        "G = Ref(rendering_kwargs=rk, **sp.G_KWARGS_TINY).eval();"
        "sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 3);"
        "G.neural_rendering_resolution = 128;"
        "D = b200eg3d.seam.convert_generator(G, device='cpu');"
        "a, b = G.state_dict(), D.state_dict();"
        "assert sorted(a) == sorted(b), 'names differ';"
        "assert all(torch.equal(a[k], b[k]) for k in a);"
        "b200eg3d.seam.save_tuned_G(D, sys.argv[1]);"
        "print('converted', len(a))"
    ) % (os.path.join(ROOT, '3dgan-inversion_b200'), os.path.join(ROOT, 'oracle'))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, 'g.pt')
        r = subprocess.run([sys.executable, '-c', code, path], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert 'converted' in r.stdout
        # the checkpoint written from an EasyDict-carrying generator loads WITHOUT dnnlib importable and with weights_only=True
        import b200eg3d
        G2 = b200eg3d.seam.load_tuned_G(path, device='cpu')
        assert G2.neural_rendering_resolution == 128 and type(G2.rendering_kwargs) is dict


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


def test_optimizer_has_no_cpu_path():
    """b200eg3d.optim.Adam runs only as the one-launch CUDA kernel: CPU parameters are rejected at construction (no silent fallback)."""
    import torch
    from b200eg3d.optim import Adam
    with pytest.raises(ValueError):
        Adam([torch.zeros(4, requires_grad=True)], lr=1e-3)
    with pytest.raises(ValueError):
        Adam([], lr=1e-3)
