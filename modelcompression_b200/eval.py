"""Batch-sharded evaluation: forward -> region decode -> NMS per image shard, detections gathered once at the end.

Reference: the eval loop of src/predict.py:116-179 / src/valid.py:61-87 is single-process; it decodes and runs NMS in
Python per image and writes per-class text files.  Images are independent, so here image indices are split into
contiguous ranges, one per rank (one process per GPU, weights replicated); the only exchange is one all_gather of
detection counts and one grouped send/recv of the (unpadded) detection rows to the destination rank at the end of the
evaluation.  Rank-major concatenation of contiguous ranges == image order, so the gathered result is
identical to a 1-GPU run.

Per batch the device work is: forward with the region decode fused into the head convolution's epilogue
(engine.CompiledDarknet.run_detect -> dense slot table), one NMS kernel that also counts each image's detection rows,
and — after ONE host synchronisation per shard to size the output — one compaction kernel per batch that writes the rows
[img, x, y, w, h, box_conf, cls_conf, cls_id] straight from the NMS output (libmcb200 mc_compact_detections).
"""
import torch
import torch.distributed as dist

from . import _lib
from .nets2_utils import decode_device, nms_device

DET_COLS = 8  # image index, x, y, w, h, det_conf, cls_max_conf, cls_max_id


def shard_range(n_items, rank, world_size):
    """Contiguous [lo, hi) of ``n_items`` owned by ``rank``: ceil(n/world) per rank, the tail ranks may be short/empty."""
    per = (n_items + world_size - 1) // world_size
    lo = min(rank * per, n_items)
    hi = min(lo + per, n_items)
    return lo, hi


def _compact_into(out, row_offsets, boxes, keep, keep_counts, cls, conf_thresh, first_image_index):
    lib = _lib.load()
    B, P, _ = boxes.shape
    with torch.cuda.device(boxes.device):
        _lib.check(lib.mc_compact_detections(boxes.data_ptr(), keep.data_ptr(), keep_counts.data_ptr(),
                                             None if cls is None else cls.data_ptr(), B, P,
                                             0 if cls is None else cls.shape[2], float(conf_thresh),
                                             int(first_image_index), row_offsets.data_ptr(), out.data_ptr(),
                                             _lib.stream_ptr()), "mc_compact_detections")


def _row_counts(boxes, keep, keep_counts, cls, conf_thresh):
    """Rows per image for tables whose NMS pass did not count them (host-level helpers below)."""
    B, P, _ = boxes.shape
    ar = torch.arange(P, device=boxes.device).unsqueeze(0)
    valid = ar < keep_counts.unsqueeze(1).long()
    if cls is None:
        return valid.sum(dim=1)
    kidx = keep.long().clamp_(0, P - 1)
    conf = torch.gather(boxes[:, :, 4], 1, kidx)
    cid = torch.gather(boxes[:, :, 6], 1, kidx).long()
    probs = torch.gather(cls, 1, kidx.unsqueeze(2).expand(B, P, cls.shape[2]))
    thr = torch.tensor(float(conf_thresh), dtype=torch.float32, device=boxes.device)
    extra = (conf.unsqueeze(2) * probs) > thr
    extra.scatter_(2, cid.unsqueeze(2), False)
    return ((1 + extra.sum(dim=2)) * valid).sum(dim=1)


def compact_detections(boxes, keep, keep_counts, first_image_index):
    """Kept boxes of a batch as rows [img, x, y, w, h, det_conf, cls_max_conf, cls_max_id], image-major, each image's
    rows in NMS output order.  boxes [B,P,8], keep [B,P] int32, keep_counts [B] int32 (device tensors)."""
    return compact_detections_validation(boxes, keep, keep_counts, None, 0.0, first_image_index)


def compact_detections_validation(boxes, keep, keep_counts, cls, conf_thresh, first_image_index):
    """Validation-mode rows (get_region_boxes(..., only_objectness=0, validation=True), src/nets2_utils.py:223-228, as
    consumed by src/predict.py:167-172): every kept box yields one row for its arg-max class and one for every other
    class c with box_conf*cls_conf[c] > conf_thresh.  Rows [img, x, y, w, h, box_conf, cls_conf, cls_id], image-major,
    NMS order, arg-max class first.  cls [B,P,nc] = softmax probabilities from decode_device(want_cls=True), or None
    for one row per kept box."""
    _lib.require_cuda(boxes, "compact_detections")
    rows = _row_counts(boxes, keep, keep_counts, cls, conf_thresh).to(torch.int64)
    offs = torch.cumsum(rows, 0) - rows
    total = int(rows.sum().item())
    out = torch.empty(total, DET_COLS, dtype=torch.float32, device=boxes.device)
    if total:
        _compact_into(out, offs, boxes.contiguous(), keep.contiguous(), keep_counts.contiguous(),
                      None if cls is None else cls.contiguous(), conf_thresh, first_image_index)
    return out


def gather_detections(local, group=None, dst=0):
    """The detections of every rank concatenated in rank order, delivered to rank ``dst`` (the other ranks get an empty
    [0, DET_COLS] tensor); dst=None delivers to every rank.  ``local``: [n, DET_COLS] float32 on the backend's device
    (CUDA for nccl, CPU for gloo).  Exchange: one all_gather of the row counts, then every rank sends its rows ONCE,
    unpadded, and the destination receives them directly into its slice of the result (grouped isend/irecv) — at 8 GPUs
    the destination receives 7/8 of the table instead of every rank receiving 8 padded copies."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c) for c in torch.cat(counts).tolist()]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    local = local.contiguous()
    if dst is None:
        n_max = max(max(counts), 1)
        padded = torch.zeros(n_max, local.shape[1], dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    ops = []
    if rank == dst:
        out = torch.empty(offs[-1], local.shape[1], dtype=local.dtype, device=local.device)
        out[offs[rank]:offs[rank + 1]] = local
        for r in range(world):
            if r != dst and counts[r]:
                ops.append(dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], dist.get_global_rank(group, r) if group else r,
                                      group=group))
    else:
        out = local.new_zeros(0, local.shape[1])
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, local, dist.get_global_rank(group, dst) if group else dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def _detect_batch_device(model, x, conf_thresh, nms_thresh, only_objectness, want_cls, fused):
    """forward -> decode -> NMS for one batch, everything left queued on the stream.
    Returns (boxes, keep, keep_counts, cls, rows)."""
    boxes = cls = counts = None
    if fused:
        from .engine import darknet_detect_forward
        got = darknet_detect_forward(model, x, conf_thresh, only_objectness, want_cls)
        if got is not None:
            boxes, cls = got
    if boxes is None:
        head = model(x)
        boxes, counts, cls = decode_device(head, conf_thresh, model.num_classes, model.anchors, model.num_anchors,
                                           only_objectness, want_cls)
    keep, keep_counts, rows = nms_device(boxes, counts, nms_thresh, cls, conf_thresh, want_rows=True)
    return boxes, keep, keep_counts, cls, rows


@torch.no_grad()
def evaluate_sharded(model, get_batch, n_images, batch_size, conf_thresh=0.005, nms_thresh=0.45, only_objectness=0,
                     rank=0, world_size=1, group=None, gather=True, validation=False, fused=True, phases=None):
    """Run detection over images [0, n_images) split across ranks.

    get_batch(lo, hi) -> float32 or uint8 CUDA tensor [hi-lo, 3, H, W] for global image indices [lo, hi).
    Returns [n_det, 8] detections, rows ordered by image index then NMS order: with ``gather`` the detections of all
    ranks on rank 0 (an empty tensor elsewhere; gather='all' delivers to every rank), else this rank's own.
    validation=True (with only_objectness=0) emits the multi-class rows the reference's scorer consumes
    (compact_detections_validation); feed them to voc_eval.mean_ap.
    fused=True uses the head convolution's decode epilogue when the model's plan supports it (same results).
    phases: optional dict; when given, the wall-clock seconds of the three phases (batches / compaction / gather) are
    stored in it — this adds a device synchronisation after each phase, so it is a diagnostic, not the fast path."""
    import time as _time

    def _mark(name, t0):
        if phases is not None:
            torch.cuda.synchronize()
            phases[name] = phases.get(name, 0.0) + _time.perf_counter() - t0
        return _time.perf_counter()

    t0 = _time.perf_counter()
    lo, hi = shard_range(n_images, rank, world_size)
    model.eval()
    want_cls = bool(validation) and not only_objectness
    raw = []  # per batch (first image, boxes, keep, keep_counts, cls, rows): no host sync per batch
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        raw.append((b0,) + _detect_batch_device(model, get_batch(b0, b1), conf_thresh, nms_thresh, only_objectness,
                                                want_cls, fused))
    dev = next(model.parameters()).device
    t0 = _mark('batches_s', t0)
    if raw:
        # ONE synchronisation per shard: the table size.  Offsets stay on the device.
        rows = torch.cat([r[5] for r in raw]).to(torch.int64)
        ends = torch.cumsum(rows, 0)
        offs = ends - rows
        total = int(ends[-1].item())
        local = torch.empty(total, DET_COLS, dtype=torch.float32, device=dev)
        pos = 0
        for b0, boxes, keep, keep_counts, cls, r in raw:
            nb = boxes.shape[0]
            if total:
                _compact_into(local, offs[pos:pos + nb], boxes, keep, keep_counts, cls, conf_thresh, b0)
            pos += nb
    else:
        local = torch.zeros(0, DET_COLS, dtype=torch.float32, device=dev)
    t0 = _mark('compaction_s', t0)
    if not gather:
        return local
    out = gather_detections(local, group, None if gather == 'all' else 0)
    _mark('gather_s', t0)
    return out
