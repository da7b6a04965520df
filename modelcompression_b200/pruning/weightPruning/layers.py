"""Drop-in for src/pruning/weightPruning/layers.py of the reference (MaskedLinear :8-30, MaskedConv2d :33-64).

Same attributes and state_dict keys (``weight``, ``bias``, ``mask`` buffer; ``name``, ``mask_flag``).  The
difference is where the arithmetic happens: the reference multiplies ``weight * mask`` on every forward
(layers.py:59); here the mask is applied once by ``set_mask`` (one CUDA kernel) and the forward of a whole
``Darknet`` runs through the tcgen05 engine (modelcompression_b200/engine.py), which packs the already-masked
weights.  A stand-alone ``MaskedConv2d.forward`` call runs the same conv kernel for that single layer.
"""
import torch
import torch.nn as nn

from ... import _lib
from .utils import to_var


def _apply_mask_inplace(weight, mask):
    """weight.data *= mask — layers.py:46 — as one libmcb200 kernel on CUDA tensors."""
    lib = _lib.load()
    _lib.require_cuda(weight, "set_mask")
    w = weight.data
    if not w.is_contiguous():
        raise RuntimeError("set_mask: weight must be contiguous")
    m = mask.to(device=w.device, dtype=torch.float32).contiguous()
    if m.numel() != w.numel():
        raise ValueError("mask shape %s does not match weight shape %s" % (tuple(m.shape), tuple(w.shape)))
    with torch.cuda.device(w.device):
        _lib.check(lib.mc_apply_masks(_lib.ptr_array([w]), _lib.ptr_array([m]), _lib.int64_array([w.numel()]), 1,
                                      _lib.stream_ptr()), "mc_apply_masks")


def _normalised_mask(mask, weight):
    """Mask as a contiguous float32 tensor of the weight's shape on the weight's device (bool / uint8 / float64 /
    strided masks are converted; a shape mismatch raises)."""
    if not torch.is_tensor(mask):
        mask = torch.as_tensor(mask)
    if mask.numel() != weight.numel():
        raise ValueError("mask shape %s does not match weight shape %s" % (tuple(mask.shape), tuple(weight.shape)))
    return mask.detach().to(device=weight.device, dtype=torch.float32).reshape(weight.shape).contiguous()


class MaskedConv2d(nn.Conv2d):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1,
                 padding=0, dilation=1, groups=1, bias=True):
        super(MaskedConv2d, self).__init__(in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                           bias)
        self.mask_flag = False
        self.name = 'MaskedConv2d'

    def set_mask(self, mask):
        # layers.py:41-47: register_buffer('mask'), weight.data *= mask, mask_flag = True
        # The reference's `weight * mask` accepts any dtype / layout; the kernels read the buffer as dense float32 on the
        # weight's device (engine_train passes mask.data_ptr() to the wgrad / dgrad packers), so it is normalised here.
        mask = _normalised_mask(mask, self.weight)
        self.register_buffer('mask', mask)
        mask_var = self.get_mask()
        if mask_var.device != self.mask.device:
            self.mask = mask_var  # keep the buffer where the weights live (to_var moved it to the GPU)
        _apply_mask_inplace(self.weight, mask_var)
        self.mask_flag = True
        # the kernel above writes weight.data behind autograd's version counter: tell the engine to re-pack
        self._mask_epoch = getattr(self, '_mask_epoch', 0) + 1

    def get_mask(self):
        return to_var(self.mask, requires_grad=False)

    def forward(self, x):
        from ...engine import single_conv_forward
        return single_conv_forward(self, x)


class MaskedLinear(nn.Linear):
    """layers.py:8-30.  Only used by the reference's YOLOv1 experiments (out of scope): the
    constructor and mask bookkeeping are kept so checkpoints load; the forward is not on the B200 path."""

    def __init__(self, in_features, out_features, bias=True):
        super(MaskedLinear, self).__init__(in_features, out_features, bias)
        self.mask_flag = False
        self.name = 'MaskedLinear'

    def set_mask(self, mask):
        mask = _normalised_mask(mask, self.weight)
        self.register_buffer('mask', mask)
        mask_var = self.get_mask()
        if mask_var.device != self.mask.device:
            self.mask = mask_var
        _apply_mask_inplace(self.weight, mask_var)
        self.mask_flag = True

    def get_mask(self):
        return to_var(self.mask, requires_grad=False)

    def forward(self, x):
        raise NotImplementedError("MaskedLinear.forward is outside the B200 hot path (YOLOv1 only, SURVEY.md §2 #6)")
