"""CPU restatement of region decode + NMS (TEST INFRASTRUCTURE — see oracle/__init__.py).

decode_np <- get_region_boxes, src/nets2_utils.py:141-234 (same torch ops, then a vectorised candidate filter in the
             reference's (cy, cx, anchor) list order)
nms_np    <- nms + bbox_iou(x1y1x2y2=False), src/nets2_utils.py:236-259, 63-98: float32, reference operation order,
             no FMA; sort key fl32(1-conf) ascending, ties by ascending list index (stable).
"""
import numpy as np
import torch


def decode_np(output, conf_thresh, num_classes, anchors_list, anchors_cell, only_objectness=1):
    """output: torch float32 [B, A*(5+nc), H, W] (CPU).  Returns per image: dict(pos, box [n,7] float32, cls [n,nc])."""
    if output.dim() == 3:
        output = output.unsqueeze(0)
    B, _, h, w = output.shape
    A, nc = anchors_cell, num_classes
    out = output.view(B * A, 5 + nc, h * w).transpose(0, 1).contiguous().view(5 + nc, B * A * h * w)
    grid_x = torch.linspace(0, w - 1, w).repeat(h, 1).repeat(B * A, 1, 1).view(B * A * h * w)
    grid_y = torch.linspace(0, h - 1, h).repeat(w, 1).t().repeat(B * A, 1, 1).view(B * A * h * w)
    xs = torch.sigmoid(out[0]) + grid_x
    ys = torch.sigmoid(out[1]) + grid_y
    anc = torch.tensor(anchors_list, dtype=torch.float32).view(A, 2)
    anchor_w = anc[:, 0:1].repeat(B, 1).repeat(1, 1, h * w).view(B * A * h * w)
    anchor_h = anc[:, 1:2].repeat(B, 1).repeat(1, 1, h * w).view(B * A * h * w)
    ws = torch.exp(out[2]) * anchor_w
    hs = torch.exp(out[3]) * anchor_h
    det = torch.sigmoid(out[4])
    cls = torch.nn.Softmax(dim=1)(out[5:5 + nc].transpose(0, 1))
    cmax, cid = torch.max(cls, 1)
    score = det if only_objectness else det * cmax
    res = []
    thr = torch.tensor(float(conf_thresh), dtype=torch.float32)
    # list order: cy, cx, anchor  -> flat index b*A*h*w + a*h*w + cy*w + cx
    a_idx = torch.arange(A).view(1, A).expand(h * w, A).reshape(-1)
    cell = torch.arange(h * w).view(h * w, 1).expand(h * w, A).reshape(-1)
    for b in range(B):
        ind = b * A * h * w + a_idx * (h * w) + cell
        sel = score[ind] > thr
        ind = ind[sel]
        pos = torch.nonzero(sel).flatten()
        box = torch.stack([xs[ind] / w, ys[ind] / h, ws[ind] / w, hs[ind] / h, det[ind], cmax[ind],
                           cid[ind].float()], dim=1)
        res.append(dict(pos=pos.numpy().astype(np.int32), box=box.numpy().astype(np.float32),
                        cls=cls[ind].numpy().astype(np.float32)))
    return res


def nms_np(boxes5, nms_thresh):
    """boxes5: float32 [n,5] = x, y, w, h, det_conf.  Returns (kept indices in output order, conf after mutation)."""
    b = np.asarray(boxes5, dtype=np.float32)
    n = b.shape[0]
    if n == 0:
        return [], b[:, 4].copy() if b.ndim == 2 else np.zeros(0, np.float32)
    x, y, w, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    conf = b[:, 4].copy()
    two = np.float32(2.0)
    l, r = (x - w / two).astype(np.float32), (x + w / two).astype(np.float32)
    t, bt = (y - h / two).astype(np.float32), (y + h / two).astype(np.float32)
    area = (w * h).astype(np.float32)
    key = (np.float32(1) - conf).astype(np.float32)
    order = np.argsort(key, kind='stable')
    thr = np.float32(nms_thresh)
    keep = []
    for i in range(n):
        bi = order[i]
        if not conf[bi] > 0:
            continue
        keep.append(int(bi))
        js = order[i + 1:]
        if js.size == 0:
            continue
        mx = np.minimum(l[bi], l[js])
        Mx = np.maximum(r[bi], r[js])
        my = np.minimum(t[bi], t[js])
        My = np.maximum(bt[bi], bt[js])
        uw = (Mx - mx).astype(np.float32)
        uh = (My - my).astype(np.float32)
        cw = ((w[bi] + w[js]).astype(np.float32) - uw).astype(np.float32)
        ch = ((h[bi] + h[js]).astype(np.float32) - uh).astype(np.float32)
        empty = (cw <= 0) | (ch <= 0)
        carea = (cw * ch).astype(np.float32)
        uarea = ((area[bi] + area[js]).astype(np.float32) - carea).astype(np.float32)
        with np.errstate(divide='ignore', invalid='ignore'):
            iou = (carea / uarea).astype(np.float32)
        sup = np.where(empty, np.float32(0.0) > thr, iou > thr)
        conf[js[sup]] = 0
    return keep, conf
