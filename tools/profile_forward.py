"""One eager forward of the bench workload (shrunk or dense) for ncu captures.  Usage: profile_forward.py [dense] [B]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200.engine import compile_darknet

dense = 'dense' in sys.argv
B = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 64
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
if not dense:
    model.set_masks(mc.quick_filter_prune(model, 40.))
plan = compile_darknet(model)
plan.use_graph = False
x = torch.rand(B, 3, 416, 416, device=dev)
with torch.no_grad():
    for _ in range(2):
        y = model(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape), [(op['name'], op.get('N'), op.get('Cin')) for op in plan.ops])
