"""PASCAL-VOC scorer on device-resident detections — replaces the file-based path of the reference:
the per-class text files written by the eval loop (src/predict.py:157-173), ``voc_eval`` (:265-395) and ``voc_ap``
(:239-263), which go through '%f' text, pickled xml annotations, pandas and a Python loop per detection.

Everything here is tensor code on the device the detections live on (the gathered output of
``eval.evaluate_sharded``): the greedy "first detection to reach a ground-truth box is the true positive" rule of
voc_eval is order-dependent only per ground-truth box, so it becomes a scatter-min of the sorted position.
The reference's text round trip is part of its arithmetic ('%f' keeps 6 decimals, values are read back as float64) and
is reproduced numerically.  Ties in confidence keep file order (stable sort); the reference's np.argsort leaves them
unspecified."""
import numpy as np
import torch

DET_COLS = 8  # image index, x, y, w, h, det_conf, cls_conf, cls_id


def _text_round(t):
    """float64 tensor -> what the reference reads back from '%f' (6 decimals, round-half-even like printf)."""
    return torch.round(t * 1e6) / 1e6


def detection_table(dets, image_sizes=None, default_size=(416, 416)):
    """src/predict.py:157-173 for all rows at once.  dets [n, 8] float32 (image, x, y, w, h, box_conf, cls_conf, cls_id),
    image_sizes [n_images, 2] (width, height) or None.  Returns (image int64 [n], cls int64 [n], conf float64 [n],
    corners float64 [n, 4]) after the float32 arithmetic of the reference and the text round trip."""
    img = dets[:, 0].long()
    if image_sizes is None:
        width = torch.full_like(dets[:, 1], float(default_size[0]))
        height = torch.full_like(dets[:, 1], float(default_size[1]))
    else:
        sz = torch.as_tensor(image_sizes, dtype=torch.float32, device=dets.device)
        width, height = sz[img, 0], sz[img, 1]
    x, y, w, h = dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    x1 = (x - w / 2.0) * width
    y1 = (y - h / 2.0) * height
    x2 = (x + w / 2.0) * width
    y2 = (y + h / 2.0) * height
    prob = dets[:, 5] * dets[:, 6]
    corners = _text_round(torch.stack([x1, y1, x2, y2], dim=1).double())
    return img, dets[:, 7].long(), _text_round(prob.double()), corners


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


def voc_eval_class(img, conf, corners, gts, cls_id, ovthresh=0.5, use_07_metric=False):
    """src/predict.py:305-395 for one class.  img/conf/corners: the rows of this class in file order; gts int64
    [m, 7] (image, class, xmin, ymin, xmax, ymax, difficult) sorted by image.  Returns (rec, prec, ap)."""
    dev = conf.device
    g = gts[gts[:, 1] == cls_id]
    npos = int((g[:, 6] == 0).sum())
    order = torch.sort(-conf, stable=True).indices
    img, corners = img[order], corners[order]
    nd = int(img.numel())
    tp = torch.zeros(nd, dtype=torch.float64, device=dev)
    fp = torch.zeros(nd, dtype=torch.float64, device=dev)
    if nd > 0:
        gimg = g[:, 0].contiguous()
        start = torch.searchsorted(gimg, img, right=False)
        end = torch.searchsorted(gimg, img, right=True)
        G = int((end - start).max()) if g.shape[0] > 0 else 0
        ovmax = torch.full((nd,), float('-inf'), dtype=torch.float64, device=dev)
        jglob = torch.zeros(nd, dtype=torch.long, device=dev)
        if G > 0:
            j = torch.arange(G, device=dev).unsqueeze(0)
            valid = j < (end - start).unsqueeze(1)
            gi = (start.unsqueeze(1) + j).clamp_(max=g.shape[0] - 1)
            bb = corners.unsqueeze(1)
            gb = g[gi][:, :, 2:6].double()
            ixmin = torch.maximum(gb[..., 0], bb[..., 0])
            iymin = torch.maximum(gb[..., 1], bb[..., 1])
            ixmax = torch.minimum(gb[..., 2], bb[..., 2])
            iymax = torch.minimum(gb[..., 3], bb[..., 3])
            iw = torch.clamp(ixmax - ixmin + 1., min=0.)
            ih = torch.clamp(iymax - iymin + 1., min=0.)
            inters = iw * ih
            uni = ((bb[..., 2] - bb[..., 0] + 1.) * (bb[..., 3] - bb[..., 1] + 1.) +
                   (gb[..., 2] - gb[..., 0] + 1.) * (gb[..., 3] - gb[..., 1] + 1.) - inters)
            ov = torch.where(valid, inters / uni, torch.full_like(inters, float('-inf')))
            ovmax, jloc = ov.max(dim=1)  # first maximum, like np.argmax
            jglob = start + jloc
        hit = ovmax > ovthresh
        difficult = torch.zeros(nd, dtype=torch.bool, device=dev)
        if g.shape[0] > 0:
            difficult = hit & (g[jglob.clamp(max=g.shape[0] - 1), 6] != 0)
        cand = hit & ~difficult
        pos = torch.arange(nd, device=dev)
        first = torch.full((max(int(g.shape[0]), 1),), nd, dtype=torch.long, device=dev)
        first.scatter_reduce_(0, jglob[cand], pos[cand], reduce='amin', include_self=True)
        is_tp = cand & (first[jglob.clamp(max=first.numel() - 1)] == pos)
        tp[is_tp] = 1.
        fp[(~hit) | (cand & ~is_tp)] = 1.
    fp = torch.cumsum(fp, 0)
    tp = torch.cumsum(tp, 0)
    rec = tp / float(npos) if npos > 0 else tp / torch.zeros((), dtype=torch.float64, device=dev)
    prec = tp / torch.clamp(tp + fp, min=float(np.finfo(np.float64).eps))
    return rec, prec, voc_ap(rec, prec, use_07_metric)


def mean_ap(dets, gts, num_classes=20, image_sizes=None, ovthresh=0.5, use_07_metric=True):
    """src/predict.py:397-437.  dets [n, 8] float32 rows (eval.evaluate_sharded(..., validation=True)), gts int64
    [m, 7] (image, class, xmin, ymin, xmax, ymax, difficult).  Returns (list of AP per class, mAP)."""
    gts = torch.as_tensor(gts, dtype=torch.long, device=dets.device)
    if gts.shape[0] > 1:
        gts = gts[torch.sort(gts[:, 0], stable=True).indices]
    img, cls, conf, corners = detection_table(dets, image_sizes)
    aps = []
    for c in range(num_classes):
        sel = cls == c
        _, _, ap = voc_eval_class(img[sel], conf[sel], corners[sel], gts, c, ovthresh, use_07_metric)
        aps.append(ap)
    return aps, float(np.mean(aps))
