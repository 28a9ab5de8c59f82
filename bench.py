#!/usr/bin/env python
"""PTI steps/sec (fwd+bwd) of the EG3D tri-plane generator at 512 px out / 128 px neural render / 48+48 depth samples.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One step = G.synthesis(ws, c, noise_mode='const', force_fp32=True) -> stand-in PTI loss (MSE@512 + MSE@128 + depth TV,
base_coach.py:101-126,294-305; LPIPS weights are not available offline) -> backward to ALL generator parameters ->
Adam step (lr 3e-4, base_coach.py:96-99).  Random-init generator of the ffhqrebalanced512-128 architecture, synthetic
latent / camera / targets.  Multi-GPU: one independent image per rank (no data-path collective), NCCL only for the
final max-over-ranks time.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, '3dgan-inversion_b200'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import torch  # noqa: E402

METRIC = 'PTI steps/sec (fwd+bwd) 512px out, 128px neural render, 48+48 depth samples'
WORKLOAD = 'ffhqrebalanced512-128 EG3D (random-init), single-image PTI step, R=128, 48+48 samples, batch 1 per GPU'
R, S, S_IMP = 128, 48, 48          # BASELINE config[1]; --render-res / --depth-samples select the config[4] stress shape


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in a background thread, 5 ms period;
    falls back to polling nvidia-smi when the NVML python binding is unavailable)."""
    REASONS = {'hw_slowdown': 0x8, 'sw_thermal_slowdown': 0x20, 'hw_thermal_slowdown': 0x40, 'sw_power_cap': 0x4}

    def __init__(self, index):
        self.index, self.samples, self.mask, self.max_mhz, self.run, self.th, self.how = index, [], 0, None, False, None, 'nvml'
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv, self.how = None, 'nvidia-smi'

    def _loop(self):
        while self.run:
            try:
                if self.nv is not None:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    self.mask |= self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    time.sleep(0.005)
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=clocks.sm,clocks.max.sm,'
                                          'clocks_event_reasons.active', '--format=csv,noheader,nounits'],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    self.mask |= int(out[2].strip(), 16)
            except Exception:
                time.sleep(0.02)

    def start(self):
        self.run = True
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.run = False
        if self.th is not None:
            self.th.join(timeout=6)
        sm = sorted(self.samples)
        reasons = [n for n, bit in self.REASONS.items() if self.mask & bit]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': reasons, 'samples': len(sm),
                'source': self.how}


def conv_flops_per_step():
    """Algorithmic conv FLOPs of one PTI step: fwd + dgrad + wgrad = 3 x 2 x 71.387 GMAC."""
    return 3 * 2 * 71.387e9


def make_problem(seed, device, r=None, s=None):
    import synth_params as sp
    import b200eg3d
    r, s = r or R, s or S
    rk = sp.rendering_kwargs(depth_resolution=s, depth_resolution_importance=s)
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_FULL).eval()
    named = dict(list(G.named_parameters()) + list(G.named_buffers()))
    sp.fill_params_(named, 7 + seed)
    G = G.to(device).float().requires_grad_(True)
    G.neural_rendering_resolution = r
    ws = sp.latent_ws(1 + seed)
    c = sp.camera(0.3, -0.2)
    t512, t_raw = sp.targets(2 + seed, r)
    return G, ws, c, t512.contiguous(), t_raw.contiguous()


def stats_ms(times):
    """median / mean / p10 / p90 of a list of per-step milliseconds."""
    t = sorted(times)
    n = len(t)
    return {'steps': n, 'median_ms': round(t[n // 2], 4), 'mean_ms': round(sum(t) / n, 4), 'p10_ms': round(t[n // 10], 4),
            'p90_ms': round(t[(9 * n) // 10], 4)}


def per_step_times(fn, n_steps, warm=3):
    """Per-step device times of fn() from CUDA event pairs on the current stream."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
    evs[0].record()
    for i in range(n_steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    return [evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)]


def rank_max(ms, world, dev):
    if world > 1:
        t = torch.tensor([ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = t.item()
    return ms


def extra_render_only(G, resident, world, dev, steps=100):
    """Forward only (eval / preview calls: single_id_coach.py:90, base_coach.py:137): G.synthesis under no_grad, replayed as a graph."""
    ws, c = resident[0], resident[1]
    with torch.no_grad():
        for _ in range(3):
            G.synthesis(ws, c, noise_mode='const', force_fp32=True)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = G.synthesis(ws, c, noise_mode='const', force_fp32=True)
    t = per_step_times(g.replay, steps)
    ms = rank_max(sorted(t)[len(t) // 2], world, dev)
    return {'metric': 'render-only (forward) images/sec, same shape', 'value': round(world / (ms * 1e-3), 2), 'unit': 'images/s', 'median_ms': round(ms, 4),
            'steps': steps, 'image_sum': round(float(out['image'].double().sum()), 3)}


def extra_cfg5(rank, world, dev, steps=30):
    """BASELINE config 5: R = 256, 96 + 96 samples (SR input antialias-resized 256 -> 128), the same PTI step."""
    from b200eg3d.coach import PTIStep
    G, ws, c, t512, _ = make_problem(100 + rank if world > 1 else 0, dev, 256, 96)
    inputs = [t.to(dev) for t in (ws, c, t512)]
    st = PTIStep(G, graphed=True, example=inputs)
    t = per_step_times(lambda: st.step(*inputs), steps)
    ms = rank_max(sorted(t)[len(t) // 2], world, dev)
    res = {'workload': 'PTI step at R=256, 96+96 samples (BASELINE config 5)', 'value': round(world / (ms * 1e-3), 2), 'unit': 'steps/s',
           'median_ms': round(ms, 3), 'steps': steps, 'loss': round(float(st.step(*inputs)), 5)}
    del st, G
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def extra_two_in_flight(rank, world, dev, steps=100):
    """Throughput mode: TWO independent single-image inversions per GPU, their graph-replayed PTI steps issued on two streams.  Not the
    headline (that stays one image per GPU, the reference's loop): it shows how much of a step is latency- rather than
    throughput-bound -- the 4x4 .. 32x32 blocks, kernel tails and the optimiser leave SMs idle that a second image fills."""
    from b200eg3d.coach import PTIStep
    probs = []
    for k in range(2):
        G, ws, c, t512, _ = make_problem(200 + 2 * rank + k, dev)
        inputs = [t.to(dev) for t in (ws, c, t512)]
        probs.append((PTIStep(G, graphed=True, example=inputs), inputs))
    streams = [torch.cuda.Stream() for _ in range(2)]
    main = torch.cuda.current_stream()

    def both():
        for (st, inp), s in zip(probs, streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                st.step(*inp)
        for s in streams:
            main.wait_stream(s)
    t = per_step_times(both, steps)
    ms = rank_max(sorted(t)[len(t) // 2], world, dev)
    res = {'workload': 'two independent single-image PTI steps in flight per GPU (two CUDA graphs on two streams)', 'value': round(2 * world / (ms * 1e-3), 2),
           'unit': 'steps/s', 'median_ms_per_pair': round(ms, 3), 'steps': steps, 'loss': [round(float(st.step(*inp)), 5) for st, inp in probs]}
    del probs
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def extra_stage1(rank, world, dev, steps=50):
    """One w-projection iteration (w_projector.py:160-268; BASELINE config 4 shape: pose + latent + noise optimised, warping loss)."""
    st, noise = make_stage1(dev, graphed=True, seed=rank)
    t = per_step_times(lambda: st.step(noise), steps)
    ms = rank_max(sorted(t)[len(t) // 2], world, dev)
    res = {'workload': 'stage-1 w-projection iteration: 2 synthesis calls fwd+bwd (grads to w, 17 noise buffers, pose), warping loss, noise regulariser; '
                       'seeded stand-in feature networks', 'value': round(world / (ms * 1e-3), 2), 'unit': 'steps/s', 'median_ms': round(ms, 3),
           'steps': steps, 'loss': float(st.step(noise))}
    del st
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def make_stage1(dev, graphed=True, seed=0):
    """ProjectionStep (b200eg3d.coach) on a random-init generator with seeded stand-in feature networks (VGG16 weights are unavailable
    offline) and a 6-D rotation leaf instead of the camera encoder."""
    import b200eg3d
    from b200eg3d import projector
    from b200eg3d.coach import ProjectionStep
    import synth_params as sp
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_IMP)
    G = b200eg3d.TriPlaneGenerator(rendering_kwargs=rk, **sp.G_KWARGS_FULL).eval()
    sp.fill_params_(dict(list(G.named_parameters()) + list(G.named_buffers())), 7 + seed)
    G = G.to(dev).float().requires_grad_(False)
    G.neural_rendering_resolution = R
    for m in G.modules():                                   # noise enters only where noise_strength != 0
        if hasattr(m, 'noise_strength'):
            m.noise_strength.data.fill_(0.05)

    def feat_net(sd):
        chans = [(3, 16), (16, 16), None, (16, 32), (32, 32), None, (32, 64), (64, 64), (64, 64), None, (64, 64), (64, 64), (64, 64)]
        layers = []
        for c in chans:
            layers += [torch.nn.MaxPool2d(2)] if c is None else [torch.nn.Conv2d(c[0], c[1], 3, padding=1), torch.nn.ReLU()]
        net = torch.nn.Sequential(*layers)
        gg = torch.Generator().manual_seed(sd)
        with torch.no_grad():
            for p_ in net.parameters():
                p_.copy_(torch.randn(p_.shape, generator=gg) * (0.1 if p_.ndim > 1 else 0.01))
        return net.to(dev).eval().requires_grad_(False)

    g = torch.Generator().manual_seed(3 + seed)
    c0 = sp.camera(0.0, 0.0).to(dev)
    target = (torch.rand(1, 3, 512, 512, generator=g) * 2 - 1).to(dev)
    pose6d = torch.tensor([[1.0, 0.05, 0.0, 0.02, -1.0, 0.03]], device=dev, requires_grad=True)
    st = ProjectionStep(G, sp.latent_ws(1 + seed)[:, :1].to(dev), [pose6d], lambda: projector.rot6d_to_rotmat(pose6d),
                        c0[:, :16].reshape(1, 4, 4).contiguous(), c0[0, 16:25].contiguous(), target, feat_net(1), feat_net(2), graphed=graphed, seed=3 + seed)
    st.pose6d = pose6d
    return st, torch.randn(1, 1, 512, device=dev) * 0.01


def gpu_aten_baseline(dev, steps=20):
    """The comparator the reference itself would be on this GPU: the same PTI step through plain ATen / cuDNN ops --
    the oracle port (oracle/eg3d_oracle.py, a statement-by-statement restatement of the reference's PyTorch path) with every tensor on
    the B200.  Two figures: strict fp32 (TF32 off, the numerics of force_fp32=True) and TF32 allowed (PyTorch's conv default)."""
    import eg3d_oracle as oracle
    import synth_params as sp
    from golden_util import param_shapes
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_IMP)
    out = {'what': 'same PTI step via ATen/cuDNN ops on this GPU (oracle port on cuda), eager, median of per-step CUDA-event times'}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    for tag, tf32 in (('fp32', False), ('tf32', True)):
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
        P = {k: torch.zeros(v) for k, v in param_shapes(sp.G_KWARGS_FULL).items()}
        sp.fill_params_(P, 7)
        P = {k: v.to(dev).requires_grad_(True) for k, v in P.items()}
        opt = torch.optim.Adam(list(P.values()), lr=3e-4, fused=True)
        ws, c = sp.latent_ws(1).to(dev), sp.camera(0.3, -0.2).to(dev)
        t512, traw = (t.to(dev) for t in sp.targets(2, R))

        def step():
            u1, u2 = torch.rand(1, R * R, S, 1, device=dev), torch.rand(R * R, S_IMP, device=dev)
            o = oracle.synthesis(P, ws, c, rk, R, u1, u2)
            loss = oracle.pti_loss(o, t512, traw)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        t = per_step_times(step, steps, warm=3)
        st = stats_ms(t)
        out[tag] = {'value': round(1e3 / st['median_ms'], 3), 'unit': 'steps/s', **st}
        del P, opt
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    return out


def fast_mode(G, resident, dev, steps=50):
    """Non-parity single-pass mode (bf16 operands once instead of the split-bf16 3-pass scheme; the counterpart of the reference's own
    fp16 blocks, networks_stylegan2.py:421-423): throughput and its measured deviation from the parity mode on the same generator."""
    import copy
    from b200eg3d import ops, _lib
    from b200eg3d.coach import PTIStep
    Gf = copy.deepcopy(G)
    ws, c, t512 = resident
    keys = ('fwd_passes', 'dgrad_passes')
    saved = {k: ops.CONFIG[k] for k in keys}
    with torch.no_grad():
        ref = Gf.synthesis(ws, c, noise_mode='const', force_fp32=True)
        ref = {k: v.clone() for k, v in ref.items()}
    prev = _lib.load().b200_set_mlp_passes(1)
    try:
        for k in keys:
            ops.CONFIG[k] = 1
        with torch.no_grad():
            out = Gf.synthesis(ws, c, noise_mode='const', force_fp32=True)
        dev_abs = {k: round(float((out[k] - ref[k]).abs().max()), 6) for k in ref}
        st = PTIStep(Gf, graphed=True, example=list(resident))
        t = per_step_times(lambda: st.step(*resident), steps)
        res = {'what': 'single-pass operands (B200EG3D_FWD_PASSES=1, DGRAD_PASSES=1, MLP_PASSES=1): NOT parity mode', 'unit': 'steps/s',
               'value': round(1e3 / sorted(t)[len(t) // 2], 2), **stats_ms(t), 'max_abs_vs_parity_mode': dev_abs}
        del st
    finally:
        for k in keys:
            ops.CONFIG[k] = saved[k]
        _lib.load().b200_set_mlp_passes(prev)
    del Gf
    torch.cuda.empty_cache()
    return res


def pti_loss(out, t512):
    """calc_loss of base_coach.py:101-126 without LPIPS, as the fused CUDA reduction (t_raw = area(t512) is formed in the kernel)."""
    from b200eg3d import losses
    return losses.pti_loss(out, t512)


def run_b200(args):
    import b200eg3d
    from b200eg3d import _lib
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    b200eg3d.ops.library_info()
    G, ws_h, c_h, t512_h, traw_h = make_problem(100 + rank if world > 1 else 0, dev)
    params = [p for n, p in G.named_parameters() if '.mapping.' not in n]
    from b200eg3d.optim import Adam
    opt = Adam(params, lr=3e-4)                      # b200_adam_step: the whole parameter list in one launch
    host = [t.pin_memory() for t in (ws_h, c_h, t512_h)]       # the raw-resolution target is derived from t512 inside the loss kernel
    resident = [t.to(dev) for t in host]

    def eager_step(ws, c, t512):
        out = G.synthesis(ws, c, noise_mode='const', force_fp32=True)
        loss = pti_loss(out, t512)
        if args.eager:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    if args.eager:
        def step(inputs):
            return eager_step(*inputs)
    else:
        from b200eg3d.graphs import GraphedStep
        graphed = GraphedStep(eager_step, resident, optimizer=opt, warmup=3)
        _lib.LAUNCHES = 0
        eager_step(*resident)                 # counts the C-ABI kernel calls one step makes (the graph replays exactly these)
        calls_per_step = _lib.LAUNCHES

        def step(inputs):
            return graphed(*inputs)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream()
    stage = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    ready, consumed, loss_ev = ([torch.cuda.Event() for _ in range(2)] for _ in range(3))
    for ev in consumed:
        ev.record()
    loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    e2e_losses = []

    def timed(n_steps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if e2e:
            # Every step: host->device copy of its inputs from pinned memory and a device->host read of its loss.  Double-buffered:
            # the copy for step i+1 runs on a second stream under step i's kernels, and step i's loss is read on the host while
            # step i+1 is already queued (one step of lag), so neither the PCIe transfer nor the host sync sits on the critical path.
            main = torch.cuda.current_stream()
            for i in range(n_steps):
                b = i & 1
                if i == 0:
                    with torch.cuda.stream(copy_stream):
                        for dst, src in zip(stage[0], host):
                            dst.copy_(src, non_blocking=True)
                        ready[0].record(copy_stream)
                main.wait_event(ready[b])
                loss = step(stage[b])
                consumed[b].record(main)
                loss_host[b].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_ev[b].record(main)
                if i + 1 < n_steps:
                    nb = b ^ 1
                    copy_stream.wait_event(consumed[nb])
                    with torch.cuda.stream(copy_stream):
                        for dst, src in zip(stage[nb], host):
                            dst.copy_(src, non_blocking=True)
                        ready[nb].record(copy_stream)
                if i >= 1:
                    loss_ev[b ^ 1].synchronize()
                    e2e_losses.append(float(loss_host[b ^ 1]))
            loss_ev[(n_steps - 1) & 1].synchronize()
            e2e_losses.append(float(loss_host[(n_steps - 1) & 1]))
        else:
            for _ in range(n_steps):
                step(resident)
        e1.record()
        barrier()
        # whole-job time = slowest rank (b200eg3d.shard.reduce_run_stats: SUM of steps, MAX of elapsed time over NCCL)
        from b200eg3d.shard import reduce_run_stats
        total_steps, ms, _ = reduce_run_stats(n_steps, e0.elapsed_time(e1), 0.0, device=dev)
        assert total_steps == n_steps * world
        return ms

    for _ in range(max(args.warmup, 3)):
        step(resident)
    if args.ncu_step:
        # exactly one eager step between cudaProfilerStart/Stop:  ncu --profile-from-start off ... bench.py --eager --ncu-step
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.LAUNCHES = 0
    ms = timed(args.steps, e2e=False)
    launches = _lib.LAUNCHES if args.eager else calls_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(args.steps, e2e=True)
    final_loss = float(step(resident).item())        # sanity value: the optimisation must behave the same in every launch mode

    # long run: >= 400 steps (>= ~1.7 s of device time) with one CUDA-event pair per step -> median / spread next to the K-step mean
    long_t = per_step_times(lambda: step(resident), max(400, args.steps), warm=0)
    long_run = stats_ms(long_t)
    long_run['median_ms'] = round(rank_max(long_run['median_ms'], world, dev), 4)
    long_run['seconds'] = round(sum(long_t) * 1e-3, 3)
    long_run['value_from_median'] = round(world / (long_run['median_ms'] * 1e-3), 3)

    # secondary workloads, measured on every rank (whole-job value = world / slowest rank's median step)
    extra = {}
    if not args.no_extra:
        for name, fn in (('render_only', lambda: extra_render_only(G, resident, world, dev)),
                         ('stage1', lambda: extra_stage1(rank, world, dev)),
                         ('cfg5', lambda: extra_cfg5(rank, world, dev)),
                         ('two_in_flight', lambda: extra_two_in_flight(rank, world, dev))):
            if name == 'two_in_flight' and world > 1:
                continue                                # a single-GPU throughput figure; the scaling runs carry the three workloads above
            try:
                extra[name] = fn()
            except Exception as e:                      # a secondary leg must never cost the headline number
                import traceback
                traceback.print_exc()
                extra[name] = {'error': f'{type(e).__name__}: {e}'[:300]}
                if world > 1:
                    raise

    # per-kernel device time of the same step, instrumented with CUDA events on the launching stream (separate steps)
    roof = roof_conv = per_call = None
    cpu = aten = fast = None
    if rank == 0:
        _lib.PROFILE = {}
        b200eg3d.ops.CONFIG['overlap'] = False        # per-kernel times are taken with the two graph branches serialised (one stream)
        for _ in range(2):
            opt.zero_grad(set_to_none=True)
            eager_step(*resident)
        torch.cuda.synchronize()
        prof = {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in _lib.PROFILE.items()}
        _lib.PROFILE = None
        try:
            def one_step():
                opt.zero_grad(set_to_none=True)
                eager_step(*resident)
            pdl_prev = _lib.load().b200_set_pdl(0)        # with PDL a traced duration would include the wait for the predecessor
            try:
                kt = cupti_kernel_times(one_step)
            finally:
                _lib.load().b200_set_pdl(pdl_prev)
        except Exception as e:                        # CUPTI unavailable: keep the event brackets only
            kt = None
            print(f'[bench] CUPTI kernel trace unavailable ({type(e).__name__}: {e}); conv-stack time from CUDA-event brackets', file=sys.stderr)
        pk, how = peaks()
        nprof = 2
        per_call = {k: round(v[0] / nprof, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        tot_ms = sum(v[0] for v in prof.values()) / nprof
        traffic = {}
        tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tp):
            traffic = json.load(open(tp))
        # (1) the dominant single kernel of the step: the backward of the fused tri-plane sampler + decoder (2 launches / step).
        #     Algorithmic bytes per launch: P1 points x (132 B incoming gradients + 4 B depth)
        #     + 25.17 MB planes read + 25.17 MB plane gradient written.
        P1 = R * R * S
        plane_bytes = 3 * 32 * 256 * 256 * 4
        bwd_bytes = P1 * 136 + 2 * plane_bytes
        bwd_ms, bwd_n = prof.get('b200_triplane_mlp_bwd', (0.0, 1))
        bwd_eager_ms = bwd_ms / max(bwd_n, 1)
        bwd_launch_ms = time_triplane_bwd(G, resident, dev)
        bwd_key = next((k for k in (kt or {}) if k.startswith('triplane_bwd_tc_kernel') or k.startswith('triplane_mlp_bwd')), None)
        ach = bwd_bytes / (bwd_launch_ms * 1e-3) / 1e9 if bwd_launch_ms > 0 else 0.0
        roof = {'kernel': 'triplane_bwd_tc_kernel (fused tri-plane sample + OSG decoder, backward; tcgen05)', 'bound': 'hbm',
                'achieved': round(ach, 1), 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': round(ach / pk['hbm_gbs'], 4),
                'traffic': traffic.get('triplane_bwd_tc_kernel', {}).get('bytes'), 'peak_source': how,
                'algorithmic_bytes_per_launch': bwd_bytes, 'launch_ms': round(bwd_launch_ms, 4), 'launches_per_step': bwd_n / nprof,
                'launch_ms_is': 'microbenchmark: 10 back-to-back launches of the kernel alone on the step\'s planes / rays, CUDA events',
                'launch_ms_in_step_cupti': round(kt[bwd_key][0] / max(kt[bwd_key][1], 1), 4) if (kt and bwd_key) else None,
                'launch_ms_eager_bracket': round(bwd_eager_ms, 4),
                'l2_scatter_gbs': round(P1 * 1536 / (bwd_launch_ms * 1e-3) / 1e9, 1) if bwd_launch_ms > 0 else None,
                'l2_atomic_peak_gbs': 6300.0,
                'share_of_kernel_time': round(bwd_launch_ms * (bwd_n / nprof) / (sum(v[0] for v in kt.values()) if kt else tot_ms), 3),
                'note': 'HBM-bound only by the compulsory-byte definition.  The binding limit is L2: every point adds 12 weighted 128-byte texel lines '
                        'into the plane gradient (1.2 GB of red.global.add.v4 per launch) and the measured L2 atomic throughput of this part is '
                        '~6.3 TB/s (scripts/microbench_red.cu: red.v4, scalar red and cp.reduce.async.bulk all saturate there) -> >= 0.19 ms per launch; '
                        'l2_scatter_gbs / l2_atomic_peak_gbs is the fraction of THAT roofline.  The features are not gathered again: the forward '
                        'leaves them as finished tensor-core operand tiles (100 MB, bulk-copied back); decoder weight gradients accumulate in tensor memory'}
        # (2) the conv stack (all tcgen05 / SIMT conv launches of the step) against the tensor roofline
        conv_ms = sum(prof[k][0] for k in prof if k.startswith('b200_conv_')) / nprof
        n_conv = sum(prof[k][1] for k in prof if k.startswith('b200_conv_')) / nprof
        conv_bracket_ms, conv_timing = conv_ms, 'CUDA-event brackets around the C-ABI calls (include launch latency)'
        kernel_ms = None
        if kt:
            is_conv = lambda k: k.startswith(('conv_tc_', 'conv_pix_kernel', 'conv_wgrad', 'conv_dgrad'))
            conv_ms = sum(v[0] for k, v in kt.items() if is_conv(k))
            tot_ms = sum(v[0] for v in kt.values())
            conv_timing = 'CUPTI kernel trace of eager steps (kernel device time)'
            kernel_ms = {k: round(v[0], 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])[:30]}
        achieved = conv_flops_per_step() / (conv_ms * 1e-3) / 1e12
        peak = pk['bf16_tflops_sustained']
        roof_conv = {'kernel': 'modulated-conv stack (b200_conv_fwd/dgrad/wgrad[_tc])', 'bound': 'tensor', 'achieved': round(achieved, 2),
                     'peak': peak, 'unit': 'TFLOP/s', 'frac': round(achieved / peak, 4), 'traffic': None, 'peak_source': how,
                     'launches_per_step': n_conv, 'ms_per_step_in_kernel': round(conv_ms, 3), 'share_of_kernel_time': round(conv_ms / tot_ms, 3),
                     'timing': conv_timing, 'ms_per_step_event_brackets': round(conv_bracket_ms, 3),
                     'issued_tflops': round(achieved * 7 / 3, 1), 'issued_frac': round(achieved * 7 / 3 / peak, 4),
                     'note': 'algorithmic FLOPs (fwd+dgrad+wgrad = 428.3 GFLOP); forward and dgrad issue 3 MMAs per product (split-bf16 parity mode), '
                             'wgrad 1: the tensor pipe executes 7/3 of the algorithmic FLOPs (issued_*)'}
        b200eg3d.ops.CONFIG['overlap'] = True
        if world == 1 and not args.no_extra:
            for name, fn in (('aten', lambda: gpu_aten_baseline(dev)), ('fast', lambda: fast_mode(G, resident, dev))):
                try:
                    res = fn()
                except Exception as e:
                    res = {'error': f'{type(e).__name__}: {e}'[:300]}
                if name == 'aten':
                    aten = res
                else:
                    fast = res
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline(1, 5)          # ~10-12 s of host work on the GPU box (1 warm-up + 5 timed PTI steps)
    if rank == 0:
        n_in = sum(t.numel() * t.element_size() for t in host)
        line = {
            'metric': METRIC, 'value': round(world * args.steps / (ms * 1e-3), 3), 'unit': 'steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (random-init generator, random targets)',
            'config': workload_config(world),
            'impl_notes': {'loss': 'fused kernel b200eg3d.losses.pti_loss', 'optimizer': 'b200eg3d.optim.Adam: one b200_adam_step launch per step',
                           'launch': 'eager (one Python-driven launch per kernel)' if args.eager else
                                     'whole step captured once in a CUDA graph (b200eg3d.graphs.GraphedStep) and replayed'},
            'e2e': {'value': round(world * args.steps / (ms_e2e * 1e-3), 3), 'unit': 'steps/s', 'h2d_bytes_per_step': n_in,
                    'd2h_bytes_per_step': 4, 'host_reads': len(e2e_losses),
                    'pipelining': 'double-buffered: H2D of step i+1 on a copy stream under step i, loss of step i read on the host one step later'},
            'timed_region_s': round(ms * 1e-3, 4), 'long_run': long_run, 'extra': extra, 'gpu_aten_baseline': aten, 'fast_mode': fast,
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof, 'roofline_conv_stack': roof_conv if rank == 0 else None,
            'kernel_ms': kernel_ms if rank == 0 else None, 'per_call_ms_event_brackets': per_call if rank == 0 else None, 'cpu_baseline': cpu, 'final_loss': round(final_loss, 6),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def cupti_kernel_times(fn, nsteps=2):
    """kernel name -> (ms per step, launches per step) from a CUPTI trace (torch.profiler) of `nsteps` eager steps.  Eager steps
    are CPU-bound, so kernels never overlap and the traced durations are the kernels' own device times (the CUDA-event
    brackets around the C-ABI calls additionally contain ~9 us of launch latency per call)."""
    import collections
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(nsteps):
            fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        name = e.name.replace('(anonymous namespace)::', '').replace('void ', '').split('(')[0][:60]
        agg[name][0] += (e.device_time if hasattr(e, 'device_time') else e.cuda_time) / 1e3
        agg[name][1] += 1
    return {k: (v[0] / nsteps, v[1] / nsteps) for k, v in agg.items()}


def b200eg3d_lib_mod():
    return __import__('b200eg3d')._lib


def time_triplane_bwd(G, resident, dev, iters=10):
    """Average duration of the dominant kernel (backward of the fused tri-plane sampler + decoder, with decoder-parameter
    gradients) launched on its own stream position: the step's real planes, rays and coarse depths, CUDA events around
    `iters` back-to-back launches after two warm-ups.  (The eager per-call brackets of the profile leg also contain launch
    latency and whatever the caching allocator does inside the call; ncu's launch list agrees with this number.)"""
    from b200eg3d._lib import call, ptr, stream
    from b200eg3d import ops
    ws, c = resident[0], resident[1]
    with torch.no_grad():
        planes = G.backbone.synthesis(ws, noise_mode='const')
        pl = G.renderer._planes_nhwc(planes.view(1, 3, 32, planes.shape[-2], planes.shape[-1])).contiguous()
        ro, rd = G.ray_sampler(c[:, :16].view(-1, 4, 4), c[:, 16:25].view(-1, 3, 3), R)
        rk = G.rendering_kwargs
        t = (torch.linspace(rk['ray_start'], rk['ray_end'], S, device=dev)[None, None]
             + torch.rand(1, R * R, S, device=dev) * ((rk['ray_end'] - rk['ray_start']) / (S - 1))).contiguous()
    W1, b1, W2, b2, lr_mul = ops._decoder_params(G.decoder)
    w = [x.detach().float().contiguous() for x in (W1, b1, W2, b2)]
    P = R * R * S
    d_rgb, d_sig = torch.randn(1, P, 32, device=dev) * 1e-3, torch.randn(1, P, device=dev) * 1e-3
    d_pl = torch.zeros_like(pl)
    dws = [torch.zeros_like(x) for x in w]
    lib = b200eg3d_lib_mod().load()
    work = torch.empty([lib.b200_triplane_bwd_workspace_bytes(1, P)], device=dev, dtype=torch.uint8)
    fsave = torch.empty([lib.b200_triplane_fsave_bytes(1, P)], device=dev, dtype=torch.uint8)
    rgb, sig = torch.empty(1, P, 32, device=dev), torch.empty(1, P, device=dev)
    call('b200_triplane_mlp_fwd', ptr(pl), 1, pl.shape[1], pl.shape[2], None, ptr(ro.contiguous()), ptr(rd.contiguous()), ptr(t), S, 0, P,
         float(rk['box_warp']), *map(ptr, w), float(lr_mul), ptr(rgb), ptr(sig), ptr(fsave), stream())     # the forward leaves its feature tiles for the backward

    def launch():
        call('b200_triplane_mlp_bwd', ptr(pl), 1, pl.shape[1], pl.shape[2], None, ptr(ro.contiguous()), ptr(rd.contiguous()), ptr(t), S, 0, P,
             float(rk['box_warp']), *map(ptr, w), float(lr_mul), ptr(d_rgb), ptr(d_sig), ptr(fsave), ptr(d_pl), None, None, None, *map(ptr, dws),
             ptr(work), work.numel(), stream())

    saved, b200eg3d_lib = None, __import__('b200eg3d')._lib
    saved, b200eg3d_lib.PROFILE = b200eg3d_lib.PROFILE, None
    for _ in range(2):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        launch()
    e1.record()
    torch.cuda.synchronize()
    b200eg3d_lib.PROFILE = saved
    return e0.elapsed_time(e1) / iters


def run_stage1(args):
    """Secondary measurement: one w-projection iteration (w_projector.py:145-270) -- two synthesis calls
    (predicted + canonical camera), warping loss, feature distance, noise regulariser, backward to w / noise buffers / pose,
    three Adam steps, noise normalisation.  Same JSON contract, different workload name.  The default run reports the same
    measurement as extra.stage1."""
    torch.cuda.set_device(0)
    dev = torch.device('cuda', 0)
    st, noise = make_stage1(dev, graphed=True)
    t = per_step_times(lambda: st.step(noise), args.steps, warm=max(args.warmup, 3))
    ms = sum(t)
    h_noise = noise.cpu().pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st.step(h_noise.to(dev, non_blocking=True)).item()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    line = {'metric': 'w-projection steps/sec (2 synthesis calls fwd+bwd + warping loss + noise regulariser) 512px out, 128px neural render, 48+48 depth samples',
            'value': round(args.steps / (ms * 1e-3), 3), 'unit': 'steps/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': round(ms / args.steps, 3), 'median_ms': stats_ms(t)['median_ms'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic (random-init generator, random target, seeded stand-in feature networks)',
            'config': {'workload': 'stage-1 w-projection iteration (w_projector.py:145-270) at R=%d, %d+%d samples; secondary measurement, not BASELINE.json\'s metric' % (R, S, S_IMP),
                       'launch': 'whole iteration captured in a CUDA graph'},
            'e2e': {'value': round(args.steps / (ms_e2e * 1e-3), 3), 'unit': 'steps/s', 'h2d_bytes_per_step': 2048, 'd2h_bytes_per_step': 4},
            'loss': float(st.step(noise).item())}
    print(json.dumps(line), flush=True)


def cpu_baseline(warm, steps):
    """The oracle port (CPU restatement of the reference path) timed on this host's cores: full PTI steps."""
    import eg3d_oracle as oracle
    import synth_params as sp
    from golden_util import param_shapes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rk = sp.rendering_kwargs(depth_resolution=S, depth_resolution_importance=S_IMP)
    P = {k: torch.zeros(v) for k, v in param_shapes(sp.G_KWARGS_FULL).items()}
    sp.fill_params_(P, 7)
    for v in P.values():
        v.requires_grad_(True)
    opt = torch.optim.Adam(list(P.values()), lr=3e-4)
    ws, c = sp.latent_ws(1), sp.camera(0.3, -0.2)
    t512, traw = sp.targets(2, R)
    times = []
    for i in range(warm + steps):
        t0 = time.perf_counter()
        u1, u2 = oracle.draw_depth_noise(11 + i, 1, R * R, S, S_IMP)
        out = oracle.synthesis(P, ws, c, rk, R, u1, u2)
        loss = oracle.pti_loss(out, t512, traw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    t = sum(times[warm:]) / steps
    return {'value': round(1.0 / t, 4), 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
            'sample': f'{steps} full PTI step(s) after {warm} warm-up, same workload, torch CPU fp32 oracle (oracle/eg3d_oracle.py)',
            'seconds_per_step': round(t, 2)}


def workload_config(world):
    """The workload both arms (--impl b200 / reference) run: the same dict in both lines; what an arm does to run it goes to `impl_notes`."""
    return {'workload': WORKLOAD, 'parallelism': f'independent images x{world}',
            'l2': 'per-step working set (>1 GB of activations and gradients) exceeds the 126 MB L2; no explicit flush',
            'loss': 'mse512 + mse128 + depth TV (LPIPS weights unavailable offline)', 'optimizer': 'Adam lr 3e-4'}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    first = cpu_baseline(0, 1)
    per = first['seconds_per_step']
    budget = 90.0
    k = max(1, min(args.steps, int(budget / per) - 1))
    w = 1
    res = cpu_baseline(0, k) if k > 0 else first
    line = {'impl': 'reference', 'metric': METRIC, 'value': res['value'], 'unit': 'steps/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', 1)),
            'steps': k, 'warmup': w, 'ms_per_step': round(1e3 / res['value'], 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic (random-init generator, random targets)',
            'config': workload_config(int(os.environ.get('WORLD_SIZE', 1))),
            'impl_notes': {'what': 'reference CPU path restated by the oracle port, all host threads, rank 0 only; steps bounded to ~90 s'},
            'cpu_baseline': dict(res, value=res['value']),
            'e2e': {'value': res['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--no-extra', action='store_true', help='skip the secondary legs (render-only, stage 1, config 5, ATen-on-GPU comparator, fast mode)')
    ap.add_argument('--render-res', type=int, default=None, help='neural rendering resolution (default 128; 256 = BASELINE config 5)')
    ap.add_argument('--depth-samples', type=int, default=None, help='coarse = fine depth samples per ray (default 48; 96 = config 5)')
    ap.add_argument('--ncu-step', action='store_true', help='profile exactly one step (cudaProfilerStart/Stop) and exit')
    ap.add_argument('--eager', action='store_true', help='launch every kernel from Python instead of replaying a CUDA graph')
    ap.add_argument('--stage', default='pti', choices=['pti', 'w_projection'], help="pti = BASELINE.json's metric (default); w_projection = stage-1 iteration (secondary)")
    a = ap.parse_args()
    if a.render_res or a.depth_samples:
        R = a.render_res or R
        S = S_IMP = a.depth_samples or S
        WORKLOAD = f'ffhqrebalanced512-128 EG3D (random-init), single-image PTI step, R={R}, {S}+{S_IMP} samples, batch 1 per GPU (non-default shape)'
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    if a.impl == 'reference':
        run_reference(a)
    elif a.stage == 'w_projection':
        run_stage1(a)
    else:
        run_b200(a)
