"""Host-side cost of weight_prune, section by section (perf_counter, median of 200 calls)."""
import os, sys, time, statistics
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200 import _lib
from modelcompression_b200.pruning.weightPruning import methods as M
dev = torch.device('cuda:0')
model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
lib = _lib.load()
mc.weight_prune(model, 70.)
torch.cuda.synchronize()
T = {}
def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e6)
for _ in range(200):
    t = time.perf_counter(); idx, flat = M._index(model); tick('index+validate', t)
    t = time.perf_counter(); plan, params = M._plan_and_tensors(model, False); tick('plan_and_tensors(total)', t)
    t = time.perf_counter(); k, gamma = M._rank_cached(plan.n, 70., np.float32); tick('rank', t)
    t = time.perf_counter(); ws_bytes = lib.mc_workspace_bytes_kth_abs_select(plan.n); ws = M._workspace(dev, ws_bytes); tick('workspace', t)
    t = time.perf_counter(); fl = torch.empty(plan.flat_len + 4, dtype=torch.float32, device=dev); base = fl.data_ptr(); tick('alloc', t)
    t = time.perf_counter(); mp = (_lib.c_void_p * len(params))(*[base + 4 * o for o in plan.offs]); tick('mask_ptrs', t)
    t = time.perf_counter(); pa = _lib.ptr_array(params); tick('ptr_array', t)
    t = time.perf_counter()
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr()
    tick('device ctx + stream_ptr', t)
    t = time.perf_counter()
    rc = lib.mc_weight_prune_masks(pa, mp, plan.sizes64, len(params), k, gamma, base + 4 * plan.flat_len, ws.data_ptr(), ws_bytes, sp)
    tick('C call (tables + memset + cooperative launch)', t)
    t = time.perf_counter(); views = [fl[o:o + ne].view(sh) for o, ne, sh in zip(plan.offs, plan.numels, plan.shapes)]; tick('views (after launch)', t)
    torch.cuda.synchronize()
for k_, v in T.items():
    print("%-50s %7.1f us" % (k_, statistics.median(v)))
