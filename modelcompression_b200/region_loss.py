"""YOLOv2 region loss on the GPU — replaces ``RegionLoss.forward`` + ``build_targets`` of the
reference (src/nets.py:282-636), which copy the predictions to the CPU and run three nested Python loops per batch
(image x ground-truth box x anchor) before copying ten target tensors back, and the autograd backward of the loss.

One libmcb200 call (csrc/region_loss.cu, two kernels + a one-thread finalize): targets per ground-truth box, then one
thread per (image, anchor, cell) computes its loss terms and d loss / d head; nothing leaves the GPU.  The reference's
quirks are kept because they define the loss value (w, h exp()-ed twice for the IoU boxes, tw = gw / anchor_w, the
label list ends at the first x == 0, a later box overwrites an earlier one on the same (anchor, cell), no positive
anchor IoU -> the LAST anchor, anchor areas are double products rounded once).  No CPU / PyTorch fallback: a CPU tensor
raises."""
import ctypes

import torch

from . import _lib

MAX_BBOX = 50


class _RegionLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, anchors, nA, nC, scales, thresh):
        lib = _lib.load()
        nB, ch, nH, nW = output.shape
        if ch != nA * (5 + nC):
            raise ValueError("region_loss: head has %d channels, expected %d anchors x (5 + %d classes)" % (ch, nA, nC))
        out = output.detach().float().contiguous()
        tgt = target.detach().to(out.device, torch.float32).contiguous().view(nB, -1)
        if tgt.shape[1] != MAX_BBOX * 5:
            raise ValueError("region_loss: target must be [nB, %d] rows of (cls, x, y, w, h)" % (MAX_BBOX * 5))
        grad = torch.empty_like(out)
        loss = torch.empty((), dtype=torch.float32, device=out.device)
        counts = torch.empty(2, dtype=torch.int32, device=out.device)
        nbytes = int(lib.mc_workspace_bytes_region_loss(nB))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=out.device)
        anc = (ctypes.c_double * (2 * nA))(*[float(a) for a in anchors[:2 * nA]])
        cs, ns, os_, cl = scales
        with torch.cuda.device(out.device):
            _lib.check(lib.mc_region_loss(out.data_ptr(), tgt.data_ptr(), nB, nA, nC, nH, nW, anc, float(cs), float(ns),
                                          float(os_), float(cl), float(thresh), grad.data_ptr(), loss.data_ptr(),
                                          counts.data_ptr(), ws.data_ptr(), nbytes, _lib.stream_ptr()), "mc_region_loss")
        ctx.save_for_backward(grad)
        ctx.counts = counts
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


def region_loss(output, target, anchors, num_anchors, num_classes, coord_scale=1, noobject_scale=1, object_scale=1,
                class_scale=1, thresh=0.6):
    """RegionLoss.forward, src/nets.py:468-610.  output [nB, nA*(5+nC), nH, nW] on the GPU (grad flows), target
    [nB, 250] rows of (cls, x, y, w, h) normalised, zero padded (dataloader.py:82-96).  Returns the scalar loss."""
    _lib.require_cuda(output, "RegionLoss.forward")
    anchor_step = int(len(anchors) / num_anchors)
    if anchor_step != 2:
        raise NotImplementedError("region_loss: anchors must be (w, h) pairs")
    return _RegionLossFn.apply(output, target, list(anchors), int(num_anchors), int(num_classes),
                               (coord_scale, noobject_scale, object_scale, class_scale), thresh)
