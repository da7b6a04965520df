import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import modelcompression_b200 as mc
from conftest import make_darknet
from modelcompression_b200.eval import evaluate_sharded, compact_detections_validation
from modelcompression_b200.nets2_utils import decode_device, nms_device
from oracle import forward_oracle
DEV = 'cuda:0'
model = make_darknet(mc.write_yolov2_voc_cfg(), seed=0, kn=True, device=DEV)
n, B, conf_t = 32, 16, 0.005
g = torch.Generator(device=DEV).manual_seed(12)
images = torch.rand(n, 3, 416, 416, device=DEV, generator=g)
get_batch = lambda lo, hi: images[lo:hi].contiguous()
dets = evaluate_sharded(model, get_batch, n, B, conf_t, 0.45, 0, validation=True)
dets_unfused = evaluate_sharded(model, get_batch, n, B, conf_t, 0.45, 0, validation=True, fused=False)
print("fused == unfused:", dets.shape, dets_unfused.shape, torch.equal(dets, dets_unfused))
parts, heads_o, heads_m = [], [], []
with torch.no_grad():
    for lo in range(0, n, B):
        x = images[lo:lo + B]
        head, _ = forward_oracle.darknet_forward_fp32(model.blocks, model.state_dict(), x)
        hm = model(x)
        heads_o.append(head); heads_m.append(hm)
        boxes, counts, cls = decode_device(head.contiguous(), conf_t, 20, model.anchors, model.num_anchors, 0, True)
        keep, kc = nms_device(boxes, counts, 0.45)
        parts.append(compact_detections_validation(boxes, keep, kc, cls, conf_t, lo))
        if lo == 0:
            bm, cm, clm = decode_device(hm.contiguous(), conf_t, 20, model.anchors, model.num_anchors, 0, True)
            print("candidates oracle/model", counts[:4].tolist(), cm[:4].tolist())
            d = (boxes[0, :int(counts[0]), :6] - bm[0, :int(cm[0]), :6]).abs() if int(counts[0]) == int(cm[0]) else None
            if d is not None:
                print("decoded box diff max per column", d.max(0).values.tolist())
            km, kcm = nms_device(bm.clone(), cm, 0.45)
            ko = keep[0, :int(kc[0])].tolist(); kmm = km[0, :int(kcm[0])].tolist()
            print("kept oracle %d model %d common %d" % (len(ko), len(kmm), len(set(ko) & set(kmm))))
ho, hm = torch.cat(heads_o), torch.cat(heads_m)
print("head rel-L2 %.4g max-rel %.4g" % (float((ho - hm).norm() / ho.norm()), float((ho - hm).abs().max() / ho.abs().max())))
print("head stats: std %.3f absmax %.3f" % (float(ho.std()), float(ho.abs().max())))
ref = torch.cat(parts)
print("rows", ref.shape[0], dets.shape[0])
