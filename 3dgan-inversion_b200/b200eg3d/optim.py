"""Adam as ONE kernel launch per step over the whole parameter list (csrc/optim.cu `b200_adam_step`).

The inversion loops of the reference update every generator parameter with torch.optim.Adam each step
(training/coaches/base_coach.py:96-99, training/projectors/w_projector.py:134-140).  This class keeps that interface
(`step()`, `zero_grad(set_to_none)`, `param_groups`, `state_dict()`) for fp32 CUDA parameters; the step counter is a
device scalar, so a captured CUDA graph (graphs.GraphedStep) replays the very same launch every step, and `lr` may be a
device scalar that follows a schedule.  Same update rule as torch.optim.Adam(amsgrad=False, maximize=False).
"""
import ctypes

import torch

from ._lib import call, on_device, stream


class _AdamTensor(ctypes.Structure):
    """B200AdamTensor of include/b200eg3d.h."""
    _fields_ = [('p', ctypes.c_void_p), ('g', ctypes.c_void_p), ('m', ctypes.c_void_p), ('v', ctypes.c_void_p), ('n', ctypes.c_long)]


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError('optimizer got an empty parameter list')
        for p in self.params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise ValueError('b200eg3d.optim.Adam: parameters must be contiguous fp32 CUDA tensors (there is no CPU path)')
        self.device = self.params[0].device
        if any(p.device != self.device for p in self.params):
            raise ValueError('b200eg3d.optim.Adam: all parameters must live on one device')
        self.param_groups = [{'params': self.params, 'lr': lr, 'betas': tuple(betas), 'eps': eps, 'weight_decay': weight_decay}]
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.step_count = torch.zeros([], device=self.device, dtype=torch.float32)
        self._ticket = torch.zeros([1], device=self.device, dtype=torch.int32)
        self._table = (_AdamTensor * len(self.params))()
        self._keep = None

    @property
    def state(self):
        """torch.optim-style view: parameter -> {'step', 'exp_avg', 'exp_avg_sq'} (the step counter is shared by all parameters)."""
        return {p: {'step': self.step_count, 'exp_avg': m, 'exp_avg_sq': v} for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq)}

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_()
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        lr = g['lr']
        keep = []
        for i, p in enumerate(self.params):
            e = self._table[i]
            gr = p.grad
            if gr is None:
                e.n = 0
                continue
            if gr.dtype != torch.float32 or not gr.is_contiguous():
                gr = gr.to(torch.float32).contiguous()
            keep.append(gr)
            e.p, e.g, e.m, e.v, e.n = p.data_ptr(), gr.data_ptr(), self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr(), p.numel()
        self._keep = keep              # converted gradient copies must outlive the (asynchronous) launch
        lr_dev, lr_val = (lr.data_ptr(), 0.0) if isinstance(lr, torch.Tensor) else (None, float(lr))
        with on_device(self.params[0]):
            call('b200_adam_step', ctypes.addressof(self._table), len(self.params), lr_dev, lr_val, float(g['betas'][0]),
                 float(g['betas'][1]), float(g['eps']), float(g['weight_decay']), self.step_count.data_ptr(),
                 self._ticket.data_ptr(), stream(self.device))

    def state_dict(self):
        return {'step': self.step_count.clone(), 'exp_avg': [t.clone() for t in self.exp_avg],
                'exp_avg_sq': [t.clone() for t in self.exp_avg_sq],
                'param_groups': [{k: v for k, v in self.param_groups[0].items() if k != 'params'}]}

    def load_state_dict(self, sd):
        self.step_count.copy_(sd['step'])
        for dst, src in zip(self.exp_avg, sd['exp_avg']):
            dst.copy_(src)
        for dst, src in zip(self.exp_avg_sq, sd['exp_avg_sq']):
            dst.copy_(src)
        self.param_groups[0].update(sd['param_groups'][0])
