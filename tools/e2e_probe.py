"""Where does the end-to-end loop of bench.py spend its step?  Variants: full, no D2H, no H2D, neither."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import modelcompression_b200 as mc

dev = torch.device('cuda:0')
print('bound cpus', bench.bind_to_gpu_numa(0))
model, masks, keep = bench.build_pruned_model(dev)
B, IMG = 64, 416
host_in = [torch.randint(0, 256, (B, 3, IMG, IMG), dtype=torch.uint8).pin_memory() for _ in range(2)]
dev_in = [torch.empty(B, 3, IMG, IMG, dtype=torch.uint8, device=dev) for _ in range(2)]
with torch.no_grad():
    y = model(dev_in[0])
host_out = [torch.empty_like(y, device='cpu').pin_memory() for _ in range(2)]
copy_stream, out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream()
ready = [torch.cuda.Event() for _ in range(2)]
consumed = [torch.cuda.Event() for _ in range(2)]
produced = [torch.cuda.Event() for _ in range(2)]
d2h_done = [torch.cuda.Event() for _ in range(2)]
keep = [None, None]


def loop(n, h2d=True, d2h=True):
    host_t = 0.0
    if h2d:
        with torch.cuda.stream(copy_stream):
            dev_in[0].copy_(host_in[0], non_blocking=True)
            ready[0].record(copy_stream)
    for i in range(n):
        t0 = time.perf_counter()
        cur, nxt = i % 2, (i + 1) % 2
        if h2d and i + 1 < n:
            with torch.cuda.stream(copy_stream):
                if i >= 1:
                    copy_stream.wait_event(consumed[nxt])
                dev_in[nxt].copy_(host_in[nxt], non_blocking=True)
                ready[nxt].record(copy_stream)
        if h2d:
            main.wait_event(ready[cur])
        out = model(dev_in[cur])
        consumed[cur].record(main)
        produced[cur].record(main)
        if d2h == 'main':
            host_out[cur].copy_(out, non_blocking=True)
        elif d2h == 'keep':
            with torch.cuda.stream(out_stream):
                out_stream.wait_event(produced[cur])
                host_out[cur].copy_(out, non_blocking=True)
                d2h_done[cur].record(out_stream)
            if keep[cur] is not None:
                pass
            keep[cur] = out  # the tensor of two steps ago is dropped here; its copy finished long ago
        elif d2h:
            with torch.cuda.stream(out_stream):
                out_stream.wait_event(produced[cur])
                host_out[cur].copy_(out, non_blocking=True)
                out.record_stream(out_stream)
        host_t += time.perf_counter() - t0
    torch.cuda.synchronize()
    return host_t


with torch.no_grad():
    for name, kw in (('full', {}), ('no_d2h', dict(d2h=False)), ('d2h_main', dict(d2h='main')), ('d2h_keep', dict(d2h='keep')),
                     ('full', {})):
        loop(6, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ht = loop(100, **kw)
        dt = time.perf_counter() - t0
        print('%-8s %.3f ms/step wall, %.3f ms/step host-side issue time' % (name, dt * 10, ht * 10))
    # single copies
    for nm, fn in (('h2d 33MB', lambda: dev_in[0].copy_(host_in[0], non_blocking=True)),
                   ('d2h 5.4MB', lambda: host_out[0].copy_(y, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        print(nm, '%.3f ms' % ((time.perf_counter() - t0) * 50))
