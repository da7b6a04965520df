"""Micro-benchmark of mc_conv_window_fwd vs mc_conv_im2col_fwd on the stem shapes (ncu-free: CUDA events, primed queue)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import modelcompression_b200 as mc  # noqa: E402
from modelcompression_b200.engine import compile_darknet  # noqa: E402

dev = torch.device('cuda:0')
B = 64
for dense in (False, True):
    for window in (True, False):
        torch.manual_seed(0)
        model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(dev).eval()
        if not dense:
            model.set_masks(mc.quick_filter_prune(model, 40.))
        model.b200_window = window
        plan = compile_darknet(model)
        gen = torch.Generator(device=dev).manual_seed(1)
        xs = [torch.randint(0, 256, (B, 3, 416, 416), dtype=torch.uint8, device=dev, generator=gen) for _ in range(6)]
        with torch.no_grad():
            for i in range(24):
                y = model(xs[i % 6])
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(3000000)
            a.record()
            for i in range(60):
                y = model(xs[i % 6])
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 60
            per = {}
            for i in range(10):
                evs = []
                plan.run(xs[i % 6], events=evs)
                torch.cuda.synchronize()
                for op, e0, e1 in evs:
                    per.setdefault(op['name'], []).append(e0.elapsed_time(e1) * 1e3)
        import statistics
        print("dense" if dense else "shrunk", "window" if window else "im2col", "ms/step %.4f" % ms, "img/s %.0f" % (B / ms * 1e3),
              "checksum %.6e" % float(y.double().abs().sum()),
              {k: round(statistics.median(v), 1) for k, v in list(per.items())[:8]}, flush=True)
