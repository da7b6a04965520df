"""Masked retrain step (BASELINE config 3: 90 % weight pruning, batch 64): event timing of forward / backward / SGD.
Usage: profile_train.py [batch] [reps]"""
import os
import statistics
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev)
model.set_masks(mc.weight_prune(model, 90.))
model.train()
opt = mc.MaskedSGD(model.parameters(), lr=1e-5, momentum=0.9, weight_decay=5e-4 * B)
torch.manual_seed(1)
x = torch.rand(B, 3, 416, 416, device=dev)
torch.manual_seed(3)
g = torch.randn(B, 125, 13, 13, device=dev)
tf, tb, ts = [], [], []
for it in range(reps + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    y = model(x)
    loss = (y * g).sum()
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        tf.append(ev[0].elapsed_time(ev[1])); tb.append(ev[1].elapsed_time(ev[2])); ts.append(ev[2].elapsed_time(ev[3]))
f, b, s = statistics.median(tf), statistics.median(tb), statistics.median(ts)
flops = 3 * 29.360e9 - 0.299e9
print("batch %d: forward %.2f ms, backward %.2f ms, SGD step %.2f ms -> %.1f images/s, %.0f TFLOP/s (87.8 GFLOP/img)" %
      (B, f, b, s, B / (f + b + s) * 1e3, flops * B / (f + b + s) / 1e9))
print("masks consistent:", mc.are_masks_consistent(model, [c.mask for c in model.masked_convs()]),
      " peak memory %.1f GB" % (torch.cuda.max_memory_allocated() / 1e9))
