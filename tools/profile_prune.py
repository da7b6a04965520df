"""weight_prune / quick_filter_prune on the 50.6 M-weight Darknet: event timings (plain run) or a few calls for an ncu
launch list.  Usage: profile_prune.py [reps]"""
import os
import statistics
import sys
import time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
n = sum(p.numel() for p in model.parameters() if p.dim() != 1)
for name, fn, nbytes in (("weight_prune_70", lambda: mc.weight_prune(model, 70.), 12 * n),
                         ("weight_prune_90", lambda: mc.weight_prune(model, 90.), 12 * n),
                         ("quick_filter_prune_40", lambda: mc.quick_filter_prune(model, 40.), 8 * n)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        fn()
        hs.append((time.perf_counter() - t0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    print("%s: %.4f ms device (min %.4f), host call %.4f ms, %.0f GB/s algorithmic = %.3f of 6542" %
          (name, ms, min(ts), statistics.median(hs), nbytes / ms / 1e6, nbytes / ms / 1e6 / 6542.4))

import ctypes
from modelcompression_b200 import _lib
from modelcompression_b200.pruning.weightPruning import methods
mc.weight_prune(model, 70.)
buf = (ctypes.c_ulonglong * 12)()
with torch.cuda.device(0):
    _lib.load().mc_debug_select_tstamps(methods._WS[('cuda', 0)].data_ptr(), ctypes.addressof(buf), _lib.stream_ptr())
t = list(buf)
print("phase stamps (us since kernel start) [start, sample loaded, pivots, pass end(block 0), pass end(all), resolved, "
      "fixed up]:", [round((x - t[0]) / 1e3, 1) if x else None for x in t[:7]])

# quick_filter_prune: device time of the call's kernels (launch queue primed), 9 calls
import statistics
ts = []
for _ in range(9):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(3000000)
    e0.record()
    mc.quick_filter_prune(model, 40.)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("quick_filter_prune(40): median %.1f us (min %.1f) -> %.3f of the 8n roofline at 6542 GB/s" %
      (statistics.median(ts), min(ts), 405076736 / (statistics.median(ts) * 1e-6) / 6542.4e9))

# fused filter pruner: phase stamps (the 512-byte state block sits behind values | thr in the call's flat buffer)
import numpy as np
plan, params, parr = methods._plan_and_tensors(model, conv_only=True)
n = sum(plan.O)
n4 = (n + 1) // 2 * 2
masks = mc.quick_filter_prune(model, 40.)
torch.cuda.synchronize()
flat_bytes = 4 * (plan.flat_len + n4 + 2)
whole = torch.empty(0, dtype=torch.uint8, device=dev).set_(masks[0].untyped_storage())
st = whole[flat_bytes + 384:flat_bytes + 384 + 64].view(torch.int64).cpu().tolist()
t = [int(x) for x in st]
print("fused filter pruner stamps (us since block 0 start): last block out of work %.1f, finisher starts %.1f, normalised %.1f, flag %.1f, done %.1f" %
      tuple((t[i] - t[0]) / 1e3 for i in (1, 2, 5, 3, 4)))
