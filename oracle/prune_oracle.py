"""CPU restatement of the reference pruners (TEST INFRASTRUCTURE — see oracle/__init__.py).

weight_prune_np       <- src/pruning/weightPruning/methods.py:9-26
quick_filter_prune_np <- src/pruning/weightPruning/methods.py:28-78
prune_rate_np         <- src/pruning/weightPruning/utils.py:59-93
np_pairwise_sum / filter_values_explicit: NumPy's float32 summation order written out (the order the CUDA kernels
in csrc/prune_filter.cu reproduce), checked here against NumPy itself.
"""
import numpy as np


def weight_prune_np(weights, pruning_perc):
    """weights: list of float32 arrays = the parameters with ndim != 1, in parameters() order.
    The reference builds a Python list of 50.6 M np.float32 scalars and np.array()s it (methods.py:14-18); the
    resulting float32 array equals the concatenation below, so the percentile is the same call on the same data."""
    all_w = np.concatenate([np.abs(w).ravel() for w in weights]).astype(np.float32, copy=False)
    threshold = np.percentile(all_w, pruning_perc)
    masks = [(np.abs(w) > threshold).astype(np.float32) for w in weights]
    return threshold, masks


def filter_values_np(w):
    """methods.py:43-51 for one conv weight [O,C,kh,kw] (float32)."""
    v = np.square(w).sum(axis=1).sum(axis=1).sum(axis=1) / (w.shape[1] * w.shape[2] * w.shape[3])
    v = v / np.sqrt(np.square(v).sum())
    v = v / np.max(v)
    return v


def quick_filter_prune_np(conv_weights, pruning_perc, want_masks=True):
    """conv_weights: list of float32 arrays with ndim == 4 in parameters() order.
    Returns (values float64 [sum O], threshold float64, keep list of bool arrays, masks or None)."""
    values = []
    per_layer = []
    for w in conv_weights:
        v = filter_values_np(w)
        per_layer.append(v)
        values = np.concatenate((values, v))  # starts from [] -> float64, as in methods.py:34,53
    threshold = np.percentile(values, pruning_perc)
    keep = [~(v < threshold) for v in per_layer]
    masks = None
    if want_masks:
        masks = []
        for w, kp in zip(conv_weights, keep):
            m = np.ones(w.shape, dtype=np.float32)
            m[~kp] = 0.
            masks.append(m)
    return values, threshold, keep, masks


def prune_rate_np(all_params):
    """utils.py:59-93: zeros of params with ndim != 1 over the number of ALL params."""
    total = sum(p.size for p in all_params)
    zeros = sum(int(np.count_nonzero(p == 0)) for p in all_params if p.ndim != 1)
    return 100. * zeros / total


# ---- NumPy's summation order, spelled out -------------------------------------------------------------------
def np_pairwise_sum(a):
    """numpy/_core/src/umath/loops_utils.h.src (pairwise_sum) for a contiguous float32 vector."""
    a = np.asarray(a, dtype=np.float32)
    n = a.shape[0]
    if n < 8:
        r = np.float32(0)
        for x in a:
            r = np.float32(r + x)
        return r
    if n <= 128:
        r = a[:8].copy()
        i = 8
        while i < n - (n % 8):
            r = (r + a[i:i + 8]).astype(np.float32)
            i += 8
        res = np.float32(np.float32(np.float32(r[0] + r[1]) + np.float32(r[2] + r[3])) +
                         np.float32(np.float32(r[4] + r[5]) + np.float32(r[6] + r[7])))
        while i < n:
            res = np.float32(res + a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return np.float32(np_pairwise_sum(a[:n2]) + np_pairwise_sum(a[n2:]))


def filter_values_explicit(w):
    """Same result as filter_values_np with every float32 operation in explicit order (what the GPU kernel does)."""
    O, C, kh, kw = w.shape
    sq = np.square(w)
    if kh * kw > 1:
        s = np.zeros((O, kh, kw), np.float32)
        for c in range(C):
            s = (s + sq[:, c]).astype(np.float32)      # sequential over c
        t = np.zeros((O, kw), np.float32)
        for h in range(kh):
            t = (t + s[:, h]).astype(np.float32)       # then over h
        tot = np.zeros((O,), np.float32)
        for x in range(kw):
            tot = (tot + t[:, x]).astype(np.float32)   # then over w
    else:
        tot = np.array([np_pairwise_sum(sq[o, :, 0, 0]) for o in range(O)], np.float32)
    v = (tot / np.float32(C * kh * kw)).astype(np.float32)
    nrm = np.sqrt(np_pairwise_sum(np.square(v)))
    v = (v / nrm).astype(np.float32)
    return (v / v.max()).astype(np.float32)
