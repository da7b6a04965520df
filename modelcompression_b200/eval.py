"""Batch-sharded evaluation: forward -> region decode -> NMS per image shard, detections gathered once at the end.

Reference: the eval loop of src/predict.py:116-179 / src/valid.py:61-87 is single-process; it decodes and runs NMS in
Python per image and writes per-class text files.  Images are independent, so here image indices are split into
contiguous ranges, one per rank (one process per GPU, weights replicated); the only exchange is one all_gather of
detection counts and one of the padded detection rows at the end of the evaluation.  Rank-major
concatenation of contiguous ranges == image order, so the gathered result is identical to a 1-GPU run.
"""
import torch
import torch.distributed as dist

from .nets2_utils import decode_device, nms_device

DET_COLS = 8  # image index, x, y, w, h, det_conf, cls_max_conf, cls_max_id


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) of ``n_items`` owned by ``rank``: ceil(n/world) per rank, the tail ranks may be short/empty."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi


def compact_detections(boxes, keep, keep_counts, first_image_index):
    """Kept boxes of a batch as rows [img, x, y, w, h, det_conf, cls_max_conf, cls_max_id], image-major, each image's
    rows in NMS output order.  boxes [B,P,8], keep [B,P] int32, keep_counts [B] int32 (device tensors)."""
    B, P, _ = boxes.shape
    ar = torch.arange(P, device=boxes.device).unsqueeze(0)
    valid = ar < keep_counts.unsqueeze(1).long()
    b_idx, slot = torch.nonzero(valid, as_tuple=True)
    cand = keep[b_idx, slot].long()
    rows = boxes[b_idx, cand]
    out = torch.empty(rows.shape[0], DET_COLS, dtype=torch.float32, device=boxes.device)
    out[:, 0] = (b_idx + first_image_index).float()
    out[:, 1:8] = rows[:, :7]
    return out


def compact_detections_validation(boxes, keep, keep_counts, cls, conf_thresh, first_image_index):
    """Validation-mode rows (get_region_boxes(..., only_objectness=0, validation=True), src/nets2_utils.py:223-228, as
    consumed by src/predict.py:167-172): every kept box yields one row for its arg-max class and one for every other
    class c with box_conf*cls_conf[c] > conf_thresh.  Rows [img, x, y, w, h, box_conf, cls_conf, cls_id], image-major,
    NMS order, arg-max class first.  cls [B,P,nc] = softmax probabilities from decode_device(want_cls=True)."""
    B, P, _ = boxes.shape
    nc = cls.shape[2]
    ar = torch.arange(P, device=boxes.device).unsqueeze(0)
    valid = ar < keep_counts.unsqueeze(1).long()
    b_idx, slot = torch.nonzero(valid, as_tuple=True)
    cand = keep[b_idx, slot].long()
    rows = boxes[b_idx, cand]                       # [K, 8]
    probs = cls[b_idx, cand]                        # [K, nc]
    cmax = rows[:, 6].long()
    thr = torch.tensor(float(conf_thresh), dtype=torch.float32, device=boxes.device)
    extra = (rows[:, 4:5] * probs) > thr            # float32 product, strict, like the reference
    extra[torch.arange(rows.shape[0], device=boxes.device), cmax] = False
    # column 0 = the arg-max pair, columns 1..nc = the other classes in ascending order
    sel = torch.cat([torch.ones(rows.shape[0], 1, dtype=torch.bool, device=boxes.device), extra], dim=1)
    k_idx, col = torch.nonzero(sel, as_tuple=True)
    out = torch.empty(k_idx.shape[0], DET_COLS, dtype=torch.float32, device=boxes.device)
    out[:, 0] = (b_idx[k_idx] + first_image_index).float()
    out[:, 1:6] = rows[k_idx, :5]
    is_max = col == 0
    cid = torch.where(is_max, cmax[k_idx], col - 1)
    out[:, 6] = torch.where(is_max, rows[k_idx, 5], probs[k_idx, cid])
    out[:, 7] = cid.float()
    return out


def gather_detections(local, group=None):
    """All ranks receive the detections of every rank concatenated in rank order.  ``local``: [n, DET_COLS] float32 on
    the backend's device (CUDA for nccl, CPU for gloo).  Two collectives: counts, then rows padded to the max count."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
    n_max = max(max(counts), 1)
    padded = torch.zeros(n_max, local.shape[1], dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


@torch.no_grad()
def evaluate_sharded(model, get_batch, n_images, batch_size, conf_thresh=0.005, nms_thresh=0.45, only_objectness=0,
                     rank=0, world_size=1, group=None, gather=True, validation=False):
    """Run detection over images [0, n_images) split across ranks.

    get_batch(lo, hi) -> float32 CUDA tensor [hi-lo, 3, H, W] for global image indices [lo, hi).
    Returns [n_det, 8] detections (all ranks' when ``gather``), rows ordered by image index then NMS order.
    validation=True (with only_objectness=0) emits the multi-class rows the reference's scorer consumes
    (compact_detections_validation); feed them to voc_eval.mean_ap."""
    lo, hi = shard_range(n_images, rank, world_size)
    model.eval()
    want_cls = bool(validation) and not only_objectness
    raw = []  # per batch (boxes, keep, keep_counts, cls): everything stays queued on the stream, no host sync per batch
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        x = get_batch(b0, b1)
        head = model(x)
        boxes, counts, cls = decode_device(head, conf_thresh, model.num_classes, model.anchors, model.num_anchors,
                                           only_objectness, want_cls)
        keep, keep_counts = nms_device(boxes, counts, nms_thresh)
        raw.append((boxes, keep, keep_counts, cls))
    dev = next(model.parameters()).device
    if raw:
        # one compaction for the whole shard (torch.nonzero synchronises: once, not once per batch); images of the shard
        # are consecutive, so the row index of the concatenation + lo is the global image index
        boxes = torch.cat([r[0] for r in raw], dim=0)
        keep = torch.cat([r[1] for r in raw], dim=0)
        keep_counts = torch.cat([r[2] for r in raw], dim=0)
        if want_cls:
            local = compact_detections_validation(boxes, keep, keep_counts, torch.cat([r[3] for r in raw], dim=0),
                                                  conf_thresh, lo)
        else:
            local = compact_detections(boxes, keep, keep_counts, lo)
    else:
        local = torch.zeros(0, DET_COLS, dtype=torch.float32, device=dev)
    return gather_detections(local, group) if gather else local
