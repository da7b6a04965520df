"""PASCAL-VOC scorer on device-resident detections — replaces the file-based path of the reference:
the per-class text files written by the eval loop (src/predict.py:157-173), ``voc_eval`` (:265-395) and ``voc_ap``
(:239-263), which go through '%f' text, pickled xml annotations, pandas and a Python loop per detection.

The detection table and the matching are libmcb200 kernels (csrc/voc_eval.cu) over ALL classes at once: rows -> scores
and corner boxes after the reference's float32 arithmetic and 6-decimal text round trip; detections in (class,
descending score) order are matched to the ground-truth boxes of their image and class in float64, and the greedy
"first detection to reach a ground-truth box is the true positive" rule of voc_eval — order-dependent only per
ground-truth box — is an atomicMin of the sorted position.  Sorting, the cumulative sums and the 11-point / envelope AP
are a fixed number of tensor ops on the same device (no per-class loop for the VOC07 metric).  Ties in confidence keep
file order (stable sort); the reference's np.argsort leaves them unspecified.  No CPU fallback: CPU tensors raise."""
import numpy as np
import torch

from . import _lib

DET_COLS = 8  # image index, x, y, w, h, det_conf, cls_conf, cls_id


def detection_table(dets, image_sizes=None, default_size=(416, 416)):
    """src/predict.py:157-173 for all rows at once (mc_voc_table).  dets [n, 8] float32 (image, x, y, w, h, box_conf,
    cls_conf, cls_id) on the GPU, image_sizes [n_images, 2] (width, height) or None.  Returns (image int64 [n], cls int64
    [n], conf float64 [n], corners float64 [n, 4]) after the float32 arithmetic of the reference and the text round trip."""
    _lib.require_cuda(dets, "voc_eval.detection_table")
    lib = _lib.load()
    dev = dets.device
    d = dets.detach().to(torch.float32).contiguous().view(-1, DET_COLS)
    n = d.shape[0]
    img = torch.empty(n, dtype=torch.long, device=dev)
    cls = torch.empty(n, dtype=torch.long, device=dev)
    conf = torch.empty(n, dtype=torch.float64, device=dev)
    corners = torch.empty(n, 4, dtype=torch.float64, device=dev)
    sz = None
    if image_sizes is not None:
        sz = torch.as_tensor(image_sizes, dtype=torch.float32, device=dev).contiguous()
    if n:
        with torch.cuda.device(dev):
            _lib.check(lib.mc_voc_table(d.data_ptr(), n, sz.data_ptr() if sz is not None else None,
                                        float(default_size[0]), float(default_size[1]), img.data_ptr(), cls.data_ptr(),
                                        conf.data_ptr(), corners.data_ptr(), _lib.stream_ptr()), "mc_voc_table")
    return img, cls, conf, corners


def voc_ap(rec, prec, use_07_metric=False):
    """src/predict.py:239-263 on float64 tensors."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            sel = rec >= float(t)
            p = float(prec[sel].max()) if bool(sel.any()) else 0
            ap = ap + p / 11.
        return ap
    one = rec.new_ones(1)
    zero = rec.new_zeros(1)
    mrec = torch.cat([zero, rec, one])
    mpre = torch.cat([zero, prec, zero])
    mpre = torch.flip(torch.cummax(torch.flip(mpre, [0]), 0).values, [0])  # monotone envelope
    i = torch.nonzero(mrec[1:] != mrec[:-1]).flatten()
    # the final sum in NumPy (pairwise summation), so the value is bit-equal to the reference's np.sum
    return float(np.sum(((mrec[i + 1] - mrec[i]) * mpre[i + 1]).cpu().numpy()))


def match_detections(dets, gts, num_classes=20, image_sizes=None, ovthresh=0.5):
    """src/predict.py:305-380 for every class at once.  Returns (cls [n] of the detections in (class, descending score)
    order, tp [n], fp [n] float64 flags in that order, bounds int64 [num_classes + 1] with class c in
    [bounds[c], bounds[c+1]), npos float64 [num_classes] non-difficult ground-truth boxes per class)."""
    lib = _lib.load()
    dev = dets.device
    gts = torch.as_tensor(gts, dtype=torch.long, device=dev).view(-1, 7)
    img, cls, conf, corners = detection_table(dets, image_sizes)
    n, m = int(img.numel()), int(gts.shape[0])
    o1 = torch.sort(-conf, stable=True).indices
    order = o1[torch.sort(cls[o1], stable=True).indices]
    cls_s, img_s = cls[order], img[order]
    corners_s = corners[order].contiguous()
    if m > 1:
        gts = gts[torch.sort(gts[:, 0], stable=True).indices]   # by image (row order kept inside an image) ...
        gts = gts[torch.sort(gts[:, 1], stable=True).indices]   # ... then by class
    hi = 0
    if n:
        hi = max(hi, int(img.max()))
    if m:
        hi = max(hi, int(gts[:, 0].max()))
    K = hi + 1
    key = (cls_s * K + img_s).contiguous()
    gkey = (gts[:, 1] * K + gts[:, 0]).contiguous()
    gbox = gts[:, 2:6].to(torch.float64).contiguous()
    gdiff = (gts[:, 6] != 0).to(torch.uint8).contiguous()
    tp = torch.zeros(n, dtype=torch.float64, device=dev)
    fp = torch.zeros(n, dtype=torch.float64, device=dev)
    if n:
        nbytes = int(lib.mc_workspace_bytes_voc_match(n, m))
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mc_voc_match(key.data_ptr(), corners_s.data_ptr(), n, gkey.data_ptr() if m else None,
                                        gbox.data_ptr() if m else None, gdiff.data_ptr() if m else None, m, float(ovthresh),
                                        tp.data_ptr(), fp.data_ptr(), ws.data_ptr(), ws.numel() * 8, _lib.stream_ptr()),
                       "mc_voc_match")
    bounds = torch.searchsorted(cls_s, torch.arange(num_classes + 1, device=dev))
    gc = gts[:, 1][(gts[:, 6] == 0) & (gts[:, 1] >= 0) & (gts[:, 1] < num_classes)]
    npos = torch.bincount(gc, minlength=num_classes).to(torch.float64)
    return cls_s, tp, fp, bounds, npos


def mean_ap(dets, gts, num_classes=20, image_sizes=None, ovthresh=0.5, use_07_metric=True):
    """src/predict.py:397-437.  dets [n, 8] float32 rows on the GPU (eval.evaluate_sharded(..., validation=True)), gts
    int64 [m, 7] (image, class, xmin, ymin, xmax, ymax, difficult).  Returns (list of AP per class, mAP)."""
    _lib.require_cuda(dets, "voc_eval.mean_ap")
    dev = dets.device
    C = int(num_classes)
    cls_s, tp, fp, bounds, npos = match_detections(dets, gts, C, image_sizes, ovthresh)
    n = int(cls_s.numel())
    ctp, cfp = torch.cumsum(tp, 0), torch.cumsum(fp, 0)
    zero = torch.zeros(1, dtype=torch.float64, device=dev)
    base_tp = torch.cat([zero, ctp])[bounds[:C]]   # cumulative counts before each class's first detection
    base_fp = torch.cat([zero, cfp])[bounds[:C]]
    valid = (cls_s >= 0) & (cls_s < C)
    ci = cls_s.clamp(0, C - 1)
    ctp = ctp - base_tp[ci]
    cfp = cfp - base_fp[ci]
    rec = ctp / npos[ci]                           # npos == 0: x / 0 like the reference's float division of arrays
    prec = ctp / torch.clamp(ctp + cfp, min=float(np.finfo(np.float64).eps))
    if not use_07_metric:
        b = bounds.tolist()
        aps = [voc_ap(rec[b[c]:b[c + 1]], prec[b[c]:b[c + 1]], False) for c in range(C)]
        return aps, float(np.mean(aps))
    # VOC07 11-point metric for all classes at once: p_k = max precision where recall >= t_k (0 if none); the eleven
    # per-class maxima come back in ONE transfer and are summed like the reference does (ap = ap + p / 11. in float64 —
    # a tensor / scalar division on the device would multiply by the rounded reciprocal instead)
    ninf = float('-inf')
    civ = ci[valid]
    pm = torch.full((11, C), ninf, dtype=torch.float64, device=dev)
    for k, t in enumerate(np.arange(0., 1.1, 0.1)):
        if n:
            val = torch.where(rec >= float(t), prec, torch.full_like(prec, ninf))[valid]
            pm[k].scatter_reduce_(0, civ, val, reduce='amax', include_self=True)
    pm = pm.cpu().numpy()
    aps = []
    for c in range(C):
        ap = 0.
        for k in range(11):
            p = float(pm[k, c]) if pm[k, c] != ninf else 0
            ap = ap + p / 11.
        aps.append(ap)
    return aps, float(np.mean(aps))
