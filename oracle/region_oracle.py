"""ORACLE (test infrastructure, never imported by the product): YOLOv2 region loss restated with PyTorch tensor ops —
``RegionLoss.forward`` + ``build_targets`` of the reference (src/nets.py:282-636), which run three nested Python loops
per batch (image x ground-truth box x anchor).  Pinned against the UNMODIFIED reference by oracle/make_golden_region.py
(loss equal, gradient <= 3e-9 relative); the product's kernels (csrc/region_loss.cu) are checked against it and against
the golden vectors.

Same arithmetic, vectorised over (image, ground-truth box, anchor, cell).  The reference's quirks are kept because they
define the loss value:
  * w, h are exp()-ed once for the loss and a SECOND time for the predicted boxes used in the IoUs (:511-512, :546-547);
  * tw, th are gw/anchor_w, gh/anchor_h, not their logarithms (:429-430);
  * the confidence target is the IoU of the ground-truth box with the prediction at its cell (:435-436);
  * a ground-truth list ends at the first box whose x is 0 (:328, :370); when two boxes land on the same
    (anchor, cell) the later one overwrites the earlier (sequential assignment);
  * if no anchor has a positive IoU with the box, ``best_n`` stays -1 and Python indexing addresses the LAST anchor;
  * anchor areas are products of Python floats (double) rounded once to float32 when they meet a tensor.
"""
import torch
import torch.nn.functional as F

MAX_BBOX = 50


def _ious_center(x1, y1, w1, h1, x2, y2, w2, h2):
    """bbox_ious(..., x1y1x2y2=False), src/nets2_utils.py:100-131, same operation order (float32)."""
    mx = torch.min(x1 - w1 / 2.0, x2 - w2 / 2.0)
    Mx = torch.max(x1 + w1 / 2.0, x2 + w2 / 2.0)
    my = torch.min(y1 - h1 / 2.0, y2 - h2 / 2.0)
    My = torch.max(y1 + h1 / 2.0, y2 + h2 / 2.0)
    uw = Mx - mx
    uh = My - my
    cw = w1 + w2 - uw
    ch = h1 + h2 - uh
    bad = (cw <= 0) | (ch <= 0)
    area1 = w1 * h1
    area2 = w2 * h2
    carea = torch.where(bad, torch.zeros_like(cw), cw * ch)
    uarea = area1 + area2 - carea
    return carea / uarea


def build_targets(pred, target, anchors, nA, nC, nH, nW, noobject_scale, object_scale, sil_thresh):
    """src/nets.py:282-440, vectorised.  pred: (px, py, pw, ph) each [nB, nA, nH, nW] (no grad); target [nB, 250].
    Returns (nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls), masks/targets [nB,nA,nH,nW]."""
    px, py, pw, ph = pred
    dev = px.device
    nB = target.shape[0]
    anchor_step = int(len(anchors) / nA)
    tg = target.float().view(nB, MAX_BBOX, 5)
    valid = torch.cumprod((tg[:, :, 1] != 0).to(torch.int32), dim=1).bool()          # list ends at the first x == 0
    gx, gy = tg[:, :, 1] * nW, tg[:, :, 2] * nH
    gw, gh = tg[:, :, 3] * nW, tg[:, :, 4] * nH
    # ---- step 1: predictions whose best IoU with any ground-truth box exceeds sil_thresh are not penalised
    P = nA * nH * nW
    iou = _ious_center(px.reshape(nB, P, 1), py.reshape(nB, P, 1), pw.reshape(nB, P, 1), ph.reshape(nB, P, 1),
                       gx.view(nB, 1, -1), gy.view(nB, 1, -1), gw.view(nB, 1, -1), gh.view(nB, 1, -1))
    iou = torch.where(valid.view(nB, 1, -1), iou, torch.zeros_like(iou))
    cur_ious = torch.clamp(iou.max(dim=2).values, min=0.)                               # torch.max(zeros, tmp)
    conf_mask = torch.full((nB, nA, nH, nW), float(noobject_scale), device=dev)
    conf_mask[cur_ious.view(nB, nA, nH, nW) > sil_thresh] = 0
    coord_mask = torch.zeros(nB, nA, nH, nW, device=dev)
    cls_mask = torch.zeros(nB, nA, nH, nW, device=dev)
    tx, ty, tw, th = (torch.zeros(nB, nA, nH, nW, device=dev) for _ in range(4))
    tconf, tcls = torch.zeros(nB, nA, nH, nW, device=dev), torch.zeros(nB, nA, nH, nW, device=dev)
    b_idx, t_idx = torch.nonzero(valid, as_tuple=True)
    nGT = int(b_idx.numel())
    if nGT == 0:
        return 0, 0, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls
    g = tg[b_idx, t_idx]
    gxv, gyv, gwv, ghv = g[:, 1] * nW, g[:, 2] * nH, g[:, 3] * nW, g[:, 4] * nH
    gi, gj = gxv.long(), gyv.long()                                                     # int(): truncation
    # ---- best anchor: IoU of (0,0,aw,ah) with (0,0,gw,gh); the anchors are Python floats in the reference
    aw = torch.tensor([anchors[anchor_step * n] for n in range(nA)], dtype=torch.float32, device=dev)
    ah = torch.tensor([anchors[anchor_step * n + 1] for n in range(nA)], dtype=torch.float32, device=dev)
    a_area = torch.tensor([anchors[anchor_step * n] * anchors[anchor_step * n + 1] for n in range(nA)],
                          dtype=torch.float32, device=dev)                              # double product, rounded once
    zero = torch.zeros(nGT, 1, device=dev)
    awb, ahb, gwb, ghb = aw.view(1, nA), ah.view(1, nA), gwv.view(-1, 1), ghv.view(-1, 1)
    mx = torch.min(zero - awb / 2.0, zero - gwb / 2.0)
    Mx = torch.max(zero + awb / 2.0, zero + gwb / 2.0)
    my = torch.min(zero - ahb / 2.0, zero - ghb / 2.0)
    My = torch.max(zero + ahb / 2.0, zero + ghb / 2.0)
    cw = awb + gwb - (Mx - mx)
    ch = ahb + ghb - (My - my)
    carea = cw * ch
    a_iou = torch.where((cw <= 0) | (ch <= 0), torch.zeros_like(carea), carea / (a_area.view(1, nA) + gwb * ghb - carea))
    best_iou = a_iou.max(dim=1).values
    ar = torch.arange(nA, device=dev).view(1, nA)
    best_n = torch.where(a_iou == best_iou.view(-1, 1), ar, torch.full_like(ar, nA)).min(dim=1).values  # first maximum
    best_n = torch.where(best_iou > 0, best_n, torch.full_like(best_n, nA - 1))         # best_n == -1 -> last anchor
    # ---- sequential assignment: the LAST box written to an (image, anchor, cell) wins
    lin = ((b_idx * nA + best_n) * nH + gj) * nW + gi
    order = torch.arange(nGT, device=dev)
    last = torch.full((nB * nA * nH * nW,), -1, dtype=torch.long, device=dev)
    last.scatter_reduce_(0, lin, order, reduce='amax', include_self=True)
    win = last[lin] == order
    lw = lin[win]
    coord_mask.view(-1)[lw] = 1
    conf_mask.view(-1)[lw] = float(object_scale)
    cls_mask.view(-1)[lw] = 1
    tx.view(-1)[lw] = (gxv - gi.float())[win]
    ty.view(-1)[lw] = (gyv - gj.float())[win]
    tw.view(-1)[lw] = (gwv / aw[best_n])[win]
    th.view(-1)[lw] = (ghv / ah[best_n])[win]
    pbx, pby = px.reshape(-1)[lin], py.reshape(-1)[lin]
    pbw, pbh = pw.reshape(-1)[lin], ph.reshape(-1)[lin]
    iou_gt = _ious_center(gxv, gyv, gwv, ghv, pbx, pby, pbw, pbh)
    tconf.view(-1)[lw] = iou_gt[win]
    tcls.view(-1)[lw] = g[:, 0][win]
    nCorrect = int((iou_gt > 0.5).sum())
    return nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls


def region_loss(output, target, anchors, num_anchors, num_classes, coord_scale=1, noobject_scale=1, object_scale=1,
                class_scale=1, thresh=0.6):
    """RegionLoss.forward, src/nets.py:468-610.  output [nB, nA*(5+nC), nH, nW] (grad flows), target [nB, 250] rows of
    (cls, x, y, w, h) normalised, zero padded (dataloader.py:82-96).  Returns the scalar loss."""
    nB, _, nH, nW = output.shape
    nA, nC = num_anchors, num_classes
    dev = output.device
    out = output.view(nB, nA, 5 + nC, nH, nW)
    x = torch.sigmoid(out[:, :, 0])
    y = torch.sigmoid(out[:, :, 1])
    w = torch.exp(out[:, :, 2])
    h = torch.exp(out[:, :, 3])
    conf = torch.sigmoid(out[:, :, 4])
    cls = out[:, :, 5:].reshape(nB * nA, nC, nH * nW).transpose(1, 2).reshape(nB * nA * nH * nW, nC)
    with torch.no_grad():
        anchor_step = int(len(anchors) / nA)
        grid_x = torch.arange(nW, dtype=torch.float32, device=dev).view(1, 1, 1, nW)
        grid_y = torch.arange(nH, dtype=torch.float32, device=dev).view(1, 1, nH, 1)
        aw = torch.tensor([anchors[anchor_step * n] for n in range(nA)], dtype=torch.float32, device=dev).view(1, nA, 1, 1)
        ah = torch.tensor([anchors[anchor_step * n + 1] for n in range(nA)], dtype=torch.float32, device=dev).view(1, nA, 1, 1)
        pred = (x + grid_x, y + grid_y, torch.exp(w) * aw, torch.exp(h) * ah)       # exp of the already exp-ed w, h
        nGT, nCorrect, coord_mask, conf_mask, cls_mask, tx, ty, tw, th, tconf, tcls = build_targets(
            pred, target.to(dev), anchors, nA, nC, nH, nW, noobject_scale, object_scale, thresh)
        cls_sel = cls_mask.view(-1) == 1
        tcls_sel = tcls.view(-1)[cls_sel].long()
        conf_mask = conf_mask.sqrt()

    def half_sse(a, b):
        return ((a - b) ** 2).sum() / 2.0

    loss_x = coord_scale * half_sse(x * coord_mask, tx * coord_mask)
    loss_y = coord_scale * half_sse(y * coord_mask, ty * coord_mask)
    loss_w = coord_scale * half_sse(w * coord_mask, tw * coord_mask)
    loss_h = coord_scale * half_sse(h * coord_mask, th * coord_mask)
    loss_conf = half_sse(conf * conf_mask, tconf * conf_mask)
    loss_cls = class_scale * F.cross_entropy(cls[cls_sel], tcls_sel, reduction='sum')
    return (loss_x + loss_y + loss_w + loss_h + loss_conf + loss_cls) / nB
