"""Fused momentum SGD for the masked retrain step.

The reference steps ``torch.optim.SGD(model.parameters(), lr=1e-5, momentum=0.9, weight_decay=5e-4*batch)``
(src/train.py:144-147, 233-235): 68 parameter tensors, which stock PyTorch walks with a dozen multi-tensor launches
(0.44 ms per step on a B200).  ``MaskedSGD`` is a drop-in ``torch.optim.Optimizer`` with the same state layout
(``momentum_buffer`` per parameter, same ``state_dict``) whose ``step()`` is ONE libmcb200 kernel over all parameters
(mc_sgd_momentum_step), in PyTorch's fp32 operation order.  Pruned weights receive exactly zero gradients from the
B200 backward, so they stay exactly zero (``are_masks_consistent`` holds after every step).  CUDA float32 only: anything
else raises (no fallback)."""
import torch

from . import _lib


class MaskedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            fresh, warm = [], []
            for p in group['params']:
                if p.grad is None:
                    continue
                _lib.require_cuda(p, "MaskedSGD.step")
                g = p.grad
                if p.dtype != torch.float32 or g.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("MaskedSGD expects contiguous float32 parameters and gradients")
                if not g.is_contiguous():
                    g = g.contiguous()
                st = self.state[p]
                first = 'momentum_buffer' not in st or st['momentum_buffer'] is None
                if first:
                    st['momentum_buffer'] = torch.empty_like(p, memory_format=torch.contiguous_format)
                (fresh if first else warm).append((p, g, st['momentum_buffer']))
            for items, first in ((fresh, 1), (warm, 0)):
                if not items:
                    continue
                dev = items[0][0].device
                with torch.cuda.device(dev):
                    _lib.check(lib.mc_sgd_momentum_step(_lib.ptr_array([i[0] for i in items]),
                                                        _lib.ptr_array([i[1] for i in items]),
                                                        _lib.ptr_array([i[2] for i in items]),
                                                        _lib.int64_array([i[0].numel() for i in items]), len(items),
                                                        float(group['lr']), float(group['momentum']),
                                                        float(group['weight_decay']), first, _lib.stream_ptr()),
                               "mc_sgd_momentum_step")
        return loss
