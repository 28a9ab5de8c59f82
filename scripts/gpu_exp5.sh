timeout 500 python -m pytest tests/test_geometry.py tests/test_stage1.py -q --tb=short -p no:cacheprovider 2>&1 | tail -20 | cut -c1-220
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5 | cut -c1-250
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '3dgan-inversion_b200'); sys.path.insert(0, 'oracle')
import b200eg3d, synth_params as sp
rk = sp.rendering_kwargs()
G = b200eg3d.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_FULL).eval()
sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 7)
G = G.cuda(); ws = sp.latent_ws(1).cuda()
for res in (256, 512):
    b200eg3d.geometry.density_grid(G, ws, shape_res=res); torch.cuda.synchronize()
    t0 = time.perf_counter(); g = b200eg3d.geometry.density_grid(G, ws, shape_res=res); torch.cuda.synchronize(); t1 = time.perf_counter()
    samples, _, _ = b200eg3d.geometry.create_samples(N=res, cube_length=rk['box_warp']); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b200eg3d.geometry.query_sigma(G, ws, samples); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'density_grid {res}^3: total {1e3*(t1-t0):.1f} ms; query_sigma (1 backbone + {res**3/1e6:.1f} M points) {ms:.2f} ms = {res**3/ms/1e6:.2f} G points/s')
PY
