"""Stage timing of the eval pipeline (forward / decode / NMS / compaction) for one batch of 64."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200.nets2_utils import decode_device, nms_device
from modelcompression_b200.eval import compact_detections_validation, compact_detections
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
model.set_masks(mc.quick_filter_prune(model, 40.))
x = torch.randint(0, 256, (64, 3, 416, 416), dtype=torch.uint8, device=dev)
def timed(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): r = f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, r
with torch.no_grad():
    t_f, head = timed(lambda: model(x))
    for (thr, oo, val) in ((0.005, 0, True), (0.25, 1, False)):
        t_d, (boxes, counts, cls) = timed(lambda: decode_device(head, thr, 20, model.anchors, model.num_anchors, oo, val))
        t_n, (keep, kc) = timed(lambda: nms_device(boxes.clone(), counts, 0.45))
        if val:
            t_c, rows = timed(lambda: compact_detections_validation(boxes, keep, kc, cls, thr, 0))
        else:
            t_c, rows = timed(lambda: compact_detections(boxes, keep, kc, 0))
        print("thr %.3f only_obj %d: forward %.3f ms, decode %.3f ms, nms %.3f ms (incl. 1.7 MB clone), compaction %.3f ms; "
              "candidates/img %.0f, kept/img %.0f, rows %d" % (thr, oo, t_f, t_d, t_n, t_c, counts.float().mean().item(),
                                                              kc.float().mean().item(), rows.shape[0]))
