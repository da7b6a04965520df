import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
for _ in range(3):
    mc.quick_filter_prune(model, 40.)
    mc.weight_prune(model, 70.)
torch.cuda.synchronize()
