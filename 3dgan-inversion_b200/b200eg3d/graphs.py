"""CUDA-graph capture of a whole optimisation step (synthesis -> loss -> backward -> optimizer).

The per-step inversion loops of the reference (w_projector.py:145-270, single_id_coach.py:64-77) launch several hundred
small kernels per step from Python; at B200 kernel speeds the step becomes launch-bound.  `GraphedStep` records the exact
kernel sequence of one eager step once (same kernels, same numerics) and replays it with a single launch per step.
Inputs are copied into static device buffers before each replay; `rendering_kwargs` and shapes must stay fixed while a
graph is alive (re-capture after changing them).
"""
import gc

import torch


class GraphedStep:
    def __init__(self, step_fn, example_inputs, optimizer=None, warmup=3):
        """step_fn(*inputs) -> scalar loss tensor; it must run forward, backward and (if any) optimizer.step() itself,
        with gradients released via optimizer.zero_grad(set_to_none=True) handled here."""
        self.static_inputs = [t.detach().clone() for t in example_inputs]
        self.optimizer = optimizer
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                if optimizer is not None:
                    optimizer.zero_grad(set_to_none=True)
                step_fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if optimizer is not None:
            optimizer.zero_grad(set_to_none=True)
        # No cyclic garbage collection while capturing: finalising an older CUDA graph (or freeing its private pool) from inside the
        # capture is a prohibited call that silently invalidates it ("operation failed due to a previous error during capture").
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = step_fn(*self.static_inputs)
        finally:
            if gc_was_on:
                gc.enable()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.loss
