"""Drop-in for the detection post-processing of src/nets2_utils.py of the reference: bbox_iou (:63-98),
get_region_boxes (:141-234), nms (:236-259), do_detect (:334-386).

The list-of-lists API of the reference is kept (boxes are ``[x, y, w, h, det_conf, cls_max_conf, cls_max_id, ...]``
with 0-dim float32 tensors and an int64 class id), but the work is one decode kernel + one NMS kernel per batch
(libmcb200: mc_decode_region / mc_nms_batched).  ``decode_device`` / ``nms_device`` / ``detect_batch`` expose the
device-resident form the batched evaluator uses (no per-box Python objects).
"""
import torch

from . import _lib


def bbox_iou(box1, box2, x1y1x2y2=True):
    """nets2_utils.py:63-98 — scalar IoU with the reference's operation order (host helper, not a hot path)."""
    if x1y1x2y2:
        mx = min(box1[0], box2[0])
        Mx = max(box1[2], box2[2])
        my = min(box1[1], box2[1])
        My = max(box1[3], box2[3])
        w1 = box1[2] - box1[0]
        h1 = box1[3] - box1[1]
        w2 = box2[2] - box2[0]
        h2 = box2[3] - box2[1]
    else:
        mx = min(box1[0] - box1[2] / 2.0, box2[0] - box2[2] / 2.0)
        Mx = max(box1[0] + box1[2] / 2.0, box2[0] + box2[2] / 2.0)
        my = min(box1[1] - box1[3] / 2.0, box2[1] - box2[3] / 2.0)
        My = max(box1[1] + box1[3] / 2.0, box2[1] + box2[3] / 2.0)
        w1 = box1[2]
        h1 = box1[3]
        w2 = box2[2]
        h2 = box2[3]
    uw = Mx - mx
    uh = My - my
    cw = w1 + w2 - uw
    ch = h1 + h2 - uh
    if cw <= 0 or ch <= 0:
        return 0.0
    area1 = w1 * h1
    area2 = w2 * h2
    carea = cw * ch
    uarea = area1 + area2 - carea
    return carea / uarea


def bbox_ious(boxes1, boxes2, x1y1x2y2=True):
    """nets2_utils.py:100-131 — element-wise IoU of two box sets given as [4, n] tensors (row i = coordinate i); returns
    a float32 tensor [n] on the inputs' device.  One libmcb200 kernel with the reference's float32 operation order."""
    lib = _lib.load()
    _lib.require_cuda(boxes1, "bbox_ious")
    _lib.require_cuda(boxes2, "bbox_ious")
    if boxes1.shape[0] != 4 or boxes2.shape != boxes1.shape:
        raise ValueError("bbox_ious expects two [4, n] tensors of the same shape")
    shape = boxes1.shape[1:]
    b1 = boxes1.detach().float().reshape(4, -1).contiguous()
    b2 = boxes2.detach().float().reshape(4, -1).contiguous()
    out = torch.empty(b1.shape[1], dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        _lib.check(lib.mc_bbox_ious(b1.data_ptr(), b2.data_ptr(), b1.shape[1], 1 if x1y1x2y2 else 0, out.data_ptr(),
                                    _lib.stream_ptr()), "mc_bbox_ious")
    return out.reshape(shape)


# ------------------------------------------------------------------------------------------------ device forms
def decode_device(output, conf_thresh, num_classes, anchors_list, anchors_cell, only_objectness=1, want_cls=False):
    """Region decode on the GPU.  Returns (boxes [B,P,8] float32, counts [B] int32, cls [B,P,nc] or None) where
    row r < counts[b] of boxes[b] is the r-th candidate in the reference's (cy, cx, anchor) order:
    x, y, w, h, det_conf, cls_max_conf, cls_max_id, source position."""
    lib = _lib.load()
    if output.dim() == 3:
        output = output.unsqueeze(0)
    _lib.require_cuda(output, "get_region_boxes")
    assert output.size(1) == (5 + num_classes) * anchors_cell
    anchor_step = int(len(anchors_list) / anchors_cell)
    if anchor_step != 2:
        raise NotImplementedError("anchor_step %d (only (w,h) anchors are supported)" % anchor_step)
    head = output.detach().float().contiguous()
    B, _, H, W = head.shape
    P = H * W * anchors_cell
    dev = head.device
    boxes = torch.empty(B, P, 8, dtype=torch.float32, device=dev)
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    cls = torch.empty(B, P, num_classes, dtype=torch.float32, device=dev) if want_cls else None
    anchors = (_lib.c_float * len(anchors_list))(*[float(a) for a in anchors_list])
    with torch.cuda.device(dev):
        _lib.check(lib.mc_decode_region(head.data_ptr(), B, H, W, anchors_cell, num_classes, anchors,
                                        float(conf_thresh), 1 if only_objectness else 0, boxes.data_ptr(),
                                        None if cls is None else cls.data_ptr(), counts.data_ptr(),
                                        _lib.stream_ptr()), "mc_decode_region")
    return boxes, counts, cls


def nms_device(boxes, counts, nms_thresh, cls=None, conf_thresh=0.0, want_rows=False):
    """Greedy NMS per image on a box table.  Returns (keep [B,P] int32, keep_counts [B] int32) — keep[b,:kc] are row
    indices of the table in the reference's output order — and, with ``want_rows``, a third tensor [B] int32 with the
    number of detection rows each image contributes (one per kept box; with ``cls`` also one per other class passing
    conf*cls > conf_thresh, nets2_utils.py:223-228).  ``boxes[...,4]`` of suppressed candidates is zeroed in place, as
    the reference mutates its input (nets2_utils.py:258).

    counts [B] int32: rows [0, counts[b]) are image b's candidates (decode_device output).  counts=None: ``boxes`` is
    the DENSE slot table of the fused decode epilogue (element 7 = slot index, -1 for a non-candidate)."""
    lib = _lib.load()
    _lib.require_cuda(boxes, "nms")
    B, P, S = boxes.shape
    assert S == 8 and boxes.is_contiguous() and boxes.dtype == torch.float32
    dev = boxes.device
    keep = torch.empty(B, P, dtype=torch.int32, device=dev)
    keep_counts = torch.empty(B, dtype=torch.int32, device=dev)
    rows = torch.empty(B, dtype=torch.int32, device=dev) if want_rows else None
    nc = 0
    if cls is not None:
        assert cls.is_contiguous() and cls.dtype == torch.float32 and cls.shape[:2] == (B, P)
        nc = cls.shape[2]
    with torch.cuda.device(dev):
        _lib.check(lib.mc_nms_detect(boxes.data_ptr(), None if counts is None else counts.data_ptr(), B, P,
                                     float(nms_thresh), keep.data_ptr(), keep_counts.data_ptr(),
                                     None if cls is None else cls.data_ptr(), nc, float(conf_thresh),
                                     None if rows is None else rows.data_ptr(), None, _lib.stream_ptr()),
                   "mc_nms_detect")
    if want_rows:
        return keep, keep_counts, rows
    return keep, keep_counts


def detect_batch(output, conf_thresh, nms_thresh, num_classes, anchors_list, anchors_cell, only_objectness=1,
                 want_cls=False):
    """decode + NMS, everything left on the device: (boxes, counts, keep, keep_counts, cls)."""
    boxes, counts, cls = decode_device(output, conf_thresh, num_classes, anchors_list, anchors_cell, only_objectness,
                                       want_cls)
    keep, keep_counts = nms_device(boxes, counts, nms_thresh)
    return boxes, counts, keep, keep_counts, cls


# ------------------------------------------------------------------------------------------------ reference API
def get_region_boxes(output, CONF_THRESH, num_classes, anchors_list, anchors_cell, only_objectness=1,
                     validation=False):
    """nets2_utils.py:141-234 — list (per image) of lists of boxes, in (cy, cx, anchor) order."""
    want_cls = bool(validation) and not only_objectness
    boxes, counts, cls = decode_device(output, CONF_THRESH, num_classes, anchors_list, anchors_cell, only_objectness,
                                       want_cls)
    boxes_h = boxes.cpu()
    counts_h = counts.cpu().tolist()
    cls_h = cls.cpu() if cls is not None else None
    thr = torch.tensor(float(CONF_THRESH), dtype=torch.float32)
    all_boxes = []
    for b, n in enumerate(counts_h):
        rows = boxes_h[b, :n]
        ids = rows[:, 6].to(torch.int64)
        img_boxes = []
        for r in range(n):
            row = rows[r]
            box = [row[0], row[1], row[2], row[3], row[4], row[5], ids[r]]
            if want_cls:
                # nets2_utils.py:223-228: every other class whose conf*cls_conf passes the threshold
                probs = cls_h[b, r]
                sel = torch.nonzero((row[4] * probs) > thr).flatten().tolist()
                cid = int(ids[r])
                for c in sel:
                    if c != cid:
                        box.append(probs[c])
                        box.append(c)
            img_boxes.append(box)
        all_boxes.append(img_boxes)
    return all_boxes


def nms(boxes, NMS_THRESH):
    """nets2_utils.py:236-259 — greedy class-agnostic NMS over centre-format boxes; returns the kept boxes in sorted
    order and sets ``box[4] = 0`` on suppressed boxes of the caller's list.  Arithmetic is float32, as it is in the
    reference when boxes come from get_region_boxes (0-dim float32 tensors).  Equal sort keys keep their list
    order (the reference's torch.sort is not stable; SURVEY.md §8c shim 3)."""
    if len(boxes) == 0:
        return boxes
    if not torch.cuda.is_available():
        raise RuntimeError("nms: no CUDA device. modelcompression_b200 has no CPU fallback for this path.")
    n = len(boxes)
    host = torch.zeros(1, n, 8, dtype=torch.float32)
    host[0, :, :5] = torch.tensor([[float(b[j]) for j in range(5)] for b in boxes], dtype=torch.float32)
    dev_boxes = host.cuda()
    counts = torch.tensor([n], dtype=torch.int32, device=dev_boxes.device)
    keep, keep_counts = nms_device(dev_boxes, counts, NMS_THRESH)
    kc = int(keep_counts.item())
    keep_h = keep[0, :kc].cpu().tolist()
    conf_after = dev_boxes[0, :, 4].cpu()
    out_boxes = [boxes[i] for i in keep_h]
    kept = set(keep_h)
    for j in range(n):
        if j not in kept and float(conf_after[j]) == 0.0 and float(boxes[j][4]) != 0.0:
            boxes[j][4] = 0  # nets2_utils.py:258
    return out_boxes


def do_detect(model, img, conf_thresh, nms_thresh, use_cuda=1, verbose=0):
    """nets2_utils.py:334-386 — single image: to-tensor -> model -> get_region_boxes(...)[0] -> nms."""
    import numpy as np
    model.eval()
    # The reference converts to float and divides by 255 on the CPU before .cuda() (nets2_utils.py:346-352).  Here the
    # uint8 pixels are shipped as they are (4x less host->device traffic) and Darknet.forward applies the same
    # x/255 inside its first-layer kernel.
    if isinstance(img, np.ndarray):
        img = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).unsqueeze(0)
        if img.dtype != torch.uint8:
            img = img.float().div(255.0)
    elif hasattr(img, 'tobytes') and hasattr(img, 'width'):  # PIL image
        width, height = img.width, img.height
        buf = torch.frombuffer(bytearray(img.tobytes()), dtype=torch.uint8)
        img = buf.view(height, width, 3).permute(2, 0, 1).contiguous().view(1, 3, height, width)
    elif not torch.is_tensor(img):
        raise TypeError("unknown image type")
    img = img.cuda()
    output = model(img).data
    boxes = get_region_boxes(output, conf_thresh, model.num_classes, model.anchors, model.num_anchors)[0]
    boxes = nms(boxes, nms_thresh)
    if verbose:
        print('  -- [do_detect] boxes (post-nms) :', len(boxes))
    return boxes
