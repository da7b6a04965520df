"""Drop-in for src/pruning/weightPruning/utils.py of the reference (to_var :8-14, prune_rate :59-93,
arg_nonzero_min :96-120, are_masks_consistent :122-133).  Full-tensor scans run as one multi-tensor CUDA
kernel launch (libmcb200: mc_count_zeros / mc_masked_residual) instead of 23 device->host copies + NumPy."""
import numpy as np
import torch

from ... import _lib


def to_var(x, requires_grad=False, volatile=False):
    """utils.py:8-14 — tensor on the GPU when one is available (``Variable`` is a no-op since torch 0.4)."""
    if torch.cuda.is_available():
        x = x.cuda()
    if requires_grad:
        x = x.requires_grad_(True)
    return x


def _segments(tensors):
    out = []
    for t in tensors:
        _lib.require_cuda(t, "modelcompression_b200 pruning utils")
        if t.dtype != torch.float32:
            raise TypeError("expected float32 parameters, got %s" % t.dtype)
        out.append(t if t.is_contiguous() else t.contiguous())
    return out


def _groups(n, size=_lib.MC_MAX_SEGMENTS):
    return [(i, min(i + size, n)) for i in range(0, n, size)]


def count_zeros(tensors):
    """Exact-zero count of every tensor, as a list of Python ints (one kernel launch per 64 tensors)."""
    lib = _lib.load()
    segs = _segments(tensors)
    counts = []
    for lo, hi in _groups(len(segs)):
        grp = segs[lo:hi]
        d_counts = torch.zeros(len(grp), dtype=torch.int64, device=grp[0].device)
        with torch.cuda.device(grp[0].device):
            _lib.check(lib.mc_count_zeros(_lib.ptr_array(grp), _lib.int64_array([t.numel() for t in grp]), len(grp),
                                          d_counts.data_ptr(), _lib.stream_ptr()), "mc_count_zeros")
        counts += d_counts.tolist()
    return counts


def prune_rate(model, verbose=True):
    """utils.py:59-93 — 100 * (#zeros in params with dim != 1) / (#all params, BN and biases included)."""
    total_nb_param = 0
    prunable = []
    for parameter in model.parameters():
        total_nb_param += parameter.numel()
        if parameter.dim() != 1:
            prunable.append(parameter.data)
    zeros = count_zeros(prunable) if prunable else []
    nb_zero_param = 0
    for layer_id, (p, z) in enumerate(zip(prunable, zeros), start=1):
        nb_zero_param += z
        if verbose:
            print("Layer {} | {} layer | {:.2f}% parameters pruned".format(
                layer_id, 'Conv' if p.dim() == 4 else 'Linear', 100. * z / p.numel()))
    pruning_perc = 100. * nb_zero_param / total_nb_param
    if verbose:
        print("Final pruning rate: {:.2f}%".format(pruning_perc))
    return pruning_perc


def arg_nonzero_min(a):
    """utils.py:96-120 — nonzero argmin of a non-negative list (kept with the reference's quirk: the start value is
    the LAST nonzero entry and index 0 as the only nonzero entry reports 'all zero')."""
    if not a:
        return
    min_ix, min_v = None, None
    for i, e in enumerate(a):
        if e != 0:
            min_ix = i
            min_v = e
    if not min_ix:
        print('Warning: all zero')
        return np.inf, np.inf
    for i, e in enumerate(a):
        if e < min_v and e != 0:
            min_v = e
            min_ix = i
    return min_v, min_ix


def are_masks_consistent(model, masks):
    """utils.py:122-133 — True iff sum(w * |mask-1|) == 0 over the conv (dim==4) parameters."""
    lib = _lib.load()
    conv_params = [p.data for p in model.parameters() if p.dim() == 4]
    assert len(conv_params) == len(masks)
    if not conv_params:
        return True
    dev = conv_params[0].device
    w = _segments(conv_params)
    m = _segments([mk.to(dev) for mk in masks])
    for a, b in zip(w, m):
        if a.numel() != b.numel():
            raise ValueError("mask shape %s does not match weight shape %s" % (tuple(b.shape), tuple(a.shape)))
    total = 0.0
    for lo, hi in _groups(len(w)):
        d_out = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mc_masked_residual(_lib.ptr_array(w[lo:hi]), _lib.ptr_array(m[lo:hi]),
                                              _lib.int64_array([t.numel() for t in w[lo:hi]]), hi - lo,
                                              d_out.data_ptr(), _lib.stream_ptr()), "mc_masked_residual")
        total += float(d_out.sum().item())
    return total == 0
