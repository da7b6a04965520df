#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json: "YOLOv2-416 pruned-fwd images/sec ...; prune-mask ms vs
HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): YOLOv2-VOC, seed-0 default init, 40 % global filter pruning
(quick_filter_prune) with the pruned filters PHYSICALLY removed, bf16 forward, batch 64 per GPU, 416x416 synthetic
images.  One step = one forward over one batch.  Under torchrun each rank runs the same per-GPU batch on its own
images (weak scaling, weights replicated, no data-path collective); the timed region is bracketed by
barrier + synchronize and the elapsed time is the max over ranks.

`--impl reference` times the CPU path (oracle port of the reference's PyTorch forward: same F.conv2d / batch_norm /
leaky_relu / max_pool2d calls, same masked weights) on the host cores, on a bounded sample per step.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 64
IMG = 416
PRUNE_PERC = 40.0
METRIC = "yolov2_416_pruned_fwd_images_per_sec"
CPU_SAMPLE_BATCH = 4


def profiled_traffic():
    """dram__bytes_read+write per launch of the dominant kernel, from the committed ncu --set full capture of this
    same workload (profiles/r1_ncu_full_conv_gemm.json, tools/run_ncu_full.sh); None when no capture is committed."""
    path = os.path.join(ROOT, 'profiles', 'r1_ncu_full_conv_gemm.json')
    try:
        with open(path) as f:
            return float(json.load(f)['dram_bytes_per_launch'])
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p.get('hbm_gbs', 6650.0), bf16=p.get('bf16_tflops', 1590.0),
                    bf16_sustained=p.get('bf16_tflops_sustained', 1400.0), source='measured')
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source='fallback')


def nvml_handle(pynvml, index):
    """NVML handle of CUDA device `index` of this process (NVML numbers physical GPUs, CUDA_VISIBLE_DEVICES renumbers
    CUDA's): matched by UUID, falling back to the index."""
    try:
        import torch
        uuid = str(torch.cuda.get_device_properties(index).uuid)
        if not uuid.startswith('GPU-'):
            uuid = 'GPU-' + uuid
        return pynvml.nvmlDeviceGetHandleByUUID(uuid)
    except Exception:
        return pynvml.nvmlDeviceGetHandleByIndex(index)


class ClockSampler(object):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = nvml_handle(pynvml, index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
                 'hw_power_brake': 0x80, 'sync_boost': 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root) BEFORE the pinned
    host buffers are allocated, so they land in memory next to the GPU: on a two-socket host a pinned buffer on the far
    socket halves the host->device rate of the end-to-end loop.  Returns the number of CPUs bound to (None if NVML or
    the affinity call is unavailable — the run then proceeds unbound)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = nvml_handle(pynvml, index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def build_pruned_model(device):
    """seed-0 default-init Darknet, 40 % filter pruning applied with set_masks (reference flow: src/train.py:167-174)."""
    import torch
    import modelcompression_b200 as mc
    torch.manual_seed(0)
    model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(device).eval()
    masks, keep = mc.quick_filter_prune(model, PRUNE_PERC, return_keep=True)
    model.set_masks(masks)
    model.b200_shrink = True
    return model, masks, keep


def cpu_forward_sample(state, blocks, batch, steps, warmup=1):
    """Oracle port of the reference forward on the host cores.  Returns (images/s, seconds per step)."""
    import torch
    from oracle import forward_oracle
    torch.manual_seed(1)
    x = torch.rand(batch, 3, IMG, IMG)
    with torch.no_grad():
        for _ in range(warmup):
            forward_oracle.darknet_forward_fp32(blocks, state, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            forward_oracle.darknet_forward_fp32(blocks, state, x)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference_arm(args, rank):
    """CPU implementation of the path (oracle port; the reference itself is Python and cannot travel to the box)."""
    if rank != 0:
        return
    import torch
    import modelcompression_b200 as mc
    from oracle import prune_oracle
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 for its workers)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    torch.manual_seed(0)
    model = mc.Darknet(mc.write_yolov2_voc_cfg()).eval()
    cw = [p.data.numpy() for p in model.parameters() if p.dim() == 4]
    _, _, _, masks = prune_oracle.quick_filter_prune_np(cw, PRUNE_PERC)
    state = dict(model.state_dict())
    for (name, p), m in zip([(n, p) for n, p in model.named_parameters() if p.dim() == 4], masks):
        state[name] = p.data * torch.from_numpy(m)  # set_mask: weight.data *= mask (layers.py:46)
    cores = torch.get_num_threads()
    ips, sps = cpu_forward_sample(state, model.blocks, CPU_SAMPLE_BATCH, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "yolov2-voc-416 40% filter-pruned (masked) forward, CPU oracle port of the reference "
                               "PyTorch path", "batch_per_step": CPU_SAMPLE_BATCH, "prune": "quick_filter_prune 40%"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps of batch %d (each step is a bounded sample of the batch-64 workload)" %
                                   (args.steps, CPU_SAMPLE_BATCH)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_masks(model_dense, peaks):
    """prune-mask ms vs the HBM roofline: weight_prune = 12n bytes, quick_filter_prune = 8n bytes."""
    import torch
    import modelcompression_b200 as mc
    n = sum(p.numel() for p in model_dense.parameters() if p.dim() != 1)
    out = {}
    for name, fn, nbytes in (("weight_prune_70", lambda: mc.weight_prune(model_dense, 70.), 12 * n),
                             ("quick_filter_prune_40", lambda: mc.quick_filter_prune(model_dense, 40.), 8 * n)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # (1) call on an IDLE GPU: event -> host prelude of the Python call -> launches -> event.  Includes the launch
        #     latency of the first kernel (the host side of the call is ~0.13 ms, the kernels ~0.12 ms).
        # (2) launch queue primed: a ~1.5 ms spin kernel runs first, so the call's launches are already queued when the
        #     start event fires and the interval is the device time of the call's kernels alone (what a pruning call
        #     costs inside a busy stream, and the kernel duration the HBM roofline is quoted against).
        idle, ts = [], []
        for primed in (False, True):
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if primed:
                    torch.cuda._sleep(3000000)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                (ts if primed else idle).append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "ms_call_on_idle_gpu": statistics.median(idle), "algorithmic_bytes": nbytes,
                     "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks['hbm_gbs'],
                     "timing": "CUDA events around the public call with the launch queue primed (device time of the "
                               "call's kernels); ms_call_on_idle_gpu adds the first launch's host latency"}
    return out


def time_eval_pipeline(model, device, B, rank, world, n_images=4952):
    """BASELINE.json configs[4]: batch-sharded evaluation of 4952 synthetic VOC2007-test-shaped images — forward,
    region decode (conf 0.005, validation mode: multi-class rows), per-image NMS (0.45), compaction, one gather of the
    detections at the end (src/predict.py:116-179 restated in eval.evaluate_sharded).  Images are generated on the
    device per batch (uint8), so the number is the device pipeline; returns images/s over all ranks."""
    import torch
    import torch.distributed as dist
    from modelcompression_b200.eval import evaluate_sharded

    g = torch.Generator(device=device).manual_seed(7 + rank)
    pool = [torch.randint(0, 256, (B, 3, IMG, IMG), dtype=torch.uint8, device=device, generator=g) for _ in range(3)]

    def get_batch(lo, hi):  # images resident in HBM: three rotating uint8 batches (33 MB each)
        return pool[(lo // B) % len(pool)][:hi - lo]

    evaluate_sharded(model, get_batch, 4 * B * world, B, 0.005, 0.45, 0, rank, world, validation=True)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    dets = evaluate_sharded(model, get_batch, n_images, B, 0.005, 0.45, 0, rank, world, validation=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {"workload": "%d synthetic 416x416 images, batch %d per GPU: forward + decode (0.005, validation) + NMS (0.45) + "
                        "detection gather" % (n_images, B), "images_per_s": n_images / dt, "seconds": dt,
            "detection_rows": int(dets.shape[0]), "includes": "images resident in HBM (3 rotating uint8 batches); no "
            "host->device copies; random-init weights make every one of the 845 boxes a candidate (worst case)"}


def time_retrain(device, peaks, B, steps=5):
    """BASELINE.json configs[2]: 90 % weight pruning, masked forward + backward + SGD step at batch B on one B200
    (src/train.py:214-235 with the synthetic loss of SURVEY.md §8d: loss = (y*g).sum())."""
    import torch
    import modelcompression_b200 as mc
    torch.manual_seed(0)
    model = mc.Darknet(mc.write_yolov2_voc_cfg()).to(device)
    model.set_masks(mc.weight_prune(model, 90.))
    model.train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-5, momentum=0.9, weight_decay=5e-4 * B)
    gen = torch.Generator(device=device).manual_seed(1)
    xs = [torch.rand(B, 3, IMG, IMG, device=device, generator=gen) for _ in range(2)]
    g = torch.randn(B, 125, 13, 13, device=device, generator=gen)
    tf, tb, ts = [], [], []
    for it in range(steps + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        loss = (model(xs[it % 2]) * g).sum()
        ev[1].record()
        loss.backward()
        ev[2].record()
        opt.step()
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 2:
            tf.append(ev[0].elapsed_time(ev[1]))
            tb.append(ev[1].elapsed_time(ev[2]))
            ts.append(ev[2].elapsed_time(ev[3]))
    f, b, s = statistics.median(tf), statistics.median(tb), statistics.median(ts)
    gflop_img = 3 * 29.360 - 0.299  # fwd + dgrad + wgrad, no dgrad for conv1
    ok = bool(mc.are_masks_consistent(model, [c.mask for c in model.masked_convs()]))
    tfl = gflop_img * B / (f + b + s)
    return {"workload": "yolov2-voc-416 90%% weight-pruned, masked forward+backward+SGD step, batch %d" % B,
            "forward_ms": f, "backward_ms": b, "sgd_ms": s, "ms_per_step": f + b + s,
            "images_per_s": B / (f + b + s) * 1e3, "algorithmic_gflop_per_image": gflop_img, "tflops": tfl,
            "frac_of_bf16_sustained": tfl / peaks['bf16_sustained'], "masks_consistent_after_steps": ok,
            "optimizer": "torch.optim.SGD(lr 1e-5, momentum 0.9, wd 5e-4*B) as src/train.py:144-147"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-dense', action='store_true', help='skip the extra un-pruned dense-network timing')
    ap.add_argument('--no-retrain', action='store_true', help='skip the extra retrain-step timing (configs[2])')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import modelcompression_b200 as mc
    from modelcompression_b200.engine import compile_darknet

    assert torch.cuda.is_available(), "bench.py (b200 arm) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    bound_cpus = bind_to_gpu_numa(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    peaks = measured_peaks()
    B = args.batch

    model, masks, keep = build_pruned_model(device)
    plan = compile_darknet(model)
    flops_img = plan.flops_per_image

    # inputs resident in HBM: 3 rotating batches (133 MB each at B=64 > 126 MB L2)
    gen = torch.Generator(device=device).manual_seed(1 + rank)
    xs = [torch.rand(B, 3, IMG, IMG, device=device, generator=gen) for _ in range(3)]

    def step(i):
        return model(xs[i % len(xs)])

    with torch.no_grad():
        # warm-up: at least W steps and at least two passes over every input buffer (the engine captures its CUDA
        # graph the second time it sees a buffer; captures must not land in the timed region)
        for i in range(max(args.warmup, 2 * len(xs))):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            y = step(i)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        elapsed_ms = e0.elapsed_time(e1)
        per_rank = None
        if world > 1:
            # every rank's own step time and clocks (diagnostic: which rank sets the max, and whether it was clocked down)
            mine = {"rank": rank, "ms_per_step": elapsed_ms / args.steps, "sm_mhz": clocks.get("sm_mhz"),
                    "reasons": clocks.get("reasons")}
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            per_rank = gathered
            t = torch.tensor([elapsed_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())
        value = B * args.steps * world / (elapsed_ms * 1e-3)

        # ---- per-kernel timing (CUDA events around every launch, same stream), K steps
        conv_ms, conv_flops, other_ms = 0.0, 0.0, 0.0
        per_layer = {}
        ksteps = min(args.steps, 20)
        for i in range(ksteps):
            events = []
            plan.run(xs[i % len(xs)], events=events)
            torch.cuda.synchronize()
            for op, a, b in events:
                ms = a.elapsed_time(b)
                if op['kind'] == 'conv':
                    conv_ms += ms
                    per_layer.setdefault(op['name'], []).append(ms)
                else:
                    other_ms += ms
                    per_layer.setdefault(op['name'], []).append(ms)
        conv_flops_step = 0.0
        conv_launches = 0
        for op in plan.ops:
            if op['kind'] == 'conv':
                conv_launches += 1
                conv_flops_step += float(op['flops_per_image']) * B
        conv_ms_per_launch = conv_ms / (ksteps * max(conv_launches, 1))
        achieved_tflops = conv_flops_step * ksteps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0

        # ---- end to end through the public API with HOST buffers (H2D of the batch + D2H of the head every step),
        #      double-buffered on a copy stream so the transfer of step i+1 overlaps the forward of step i
        # host images are uint8 NCHW, what the reference's do_detect receives from PIL/cv2 before its CPU-side
        # float().div(255) (src/nets2_utils.py:346-352); Darknet.forward scales them inside the first-layer kernel
        host_in = [torch.randint(0, 256, (B, 3, IMG, IMG), dtype=torch.uint8).pin_memory() for _ in range(2)]
        host_out = [torch.empty(B, y.shape[1], y.shape[2], y.shape[3]).pin_memory() for _ in range(2)]
        dev_in = [torch.empty(B, 3, IMG, IMG, dtype=torch.uint8, device=device) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=device)
        out_stream = torch.cuda.Stream(device=device)
        main_stream = torch.cuda.current_stream()
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        produced = [torch.cuda.Event() for _ in range(2)]

        in_flight = [None, None]
        d2h_done = [torch.cuda.Event() for _ in range(2)]

        def e2e_loop(nsteps):
            with torch.cuda.stream(copy_stream):
                dev_in[0].copy_(host_in[0], non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(nsteps):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < nsteps:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(consumed[nxt])
                        dev_in[nxt].copy_(host_in[nxt], non_blocking=True)
                        ready[nxt].record(copy_stream)
                main_stream.wait_event(ready[cur])
                if in_flight[cur] is not None:
                    d2h_done[cur].synchronize()  # the copy of two steps ago (finished long since): its block may be reused
                    in_flight[cur] = None
                out = model(dev_in[cur])
                consumed[cur].record(main_stream)
                produced[cur].record(main_stream)
                with torch.cuda.stream(out_stream):  # device->host read of the head, off the compute stream
                    out_stream.wait_event(produced[cur])
                    host_out[cur].copy_(out, non_blocking=True)
                    d2h_done[cur].record(out_stream)
                # keep the head alive until the slot is reused two steps later (its copy has long finished by then)
                # instead of Tensor.record_stream: the allocator's per-malloc event bookkeeping for stream-recorded
                # blocks cost 0.2-0.9 ms of HOST time per step on some hosts (tools/e2e_probe.py)
                in_flight[cur] = out
            torch.cuda.synchronize()

        # raw host->device rate of this host for the same pinned buffer (the end-to-end loop is bound by it whenever
        # 33 MB / rate exceeds the forward time): reported next to the end-to-end value
        with torch.cuda.stream(copy_stream):
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dev_in[0].copy_(host_in[0], non_blocking=True)
            h0.record(copy_stream)
            for _ in range(5):
                dev_in[0].copy_(host_in[0], non_blocking=True)
            h1.record(copy_stream)
        torch.cuda.synchronize()
        h2d_gbs = 5 * host_in[0].numel() / (h0.elapsed_time(h1) * 1e-3) / 1e9

        e2e_loop(4)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        e2e_value = B * args.steps * world / e2e_s

    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "yolov2-voc-416 (seed-0 default init) 40% filter-pruned, filters physically removed, "
                               "forward", "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "prune": "quick_filter_prune 40%% -> %d/%d filters kept" % (sum(int(k.numel()) for k in keep),
                                                                            sum(m.shape[0] for m in masks)),
                   "algorithmic_gflop_per_image": flops_img / 1e9,
                   "l2": "inputs larger than L2: %d rotating %.0f MB batches" % (len(xs), B * 3 * IMG * IMG * 4 / 1e6)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * IMG * IMG,
                "d2h_bytes_per_step": int(y.numel() * 4),
                "input": "uint8 NCHW images in pinned host memory (do_detect's input type), x/255 on the device",
                "host_cpus_bound_to_gpu_numa_node": bound_cpus, "h2d_gbs_measured": h2d_gbs,
                "h2d_bound_images_per_s": h2d_gbs * 1e9 / (3 * IMG * IMG),
                "note": "double-buffered: step time = max(H2D of the next batch, forward, D2H); on this host the H2D "
                        "of the 33 MB batch is the longest of the three when h2d_bound_images_per_s < value"},
        "gpu_launches": plan.num_launches * args.steps,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_tcgen05_kernel + conv_gemm_tcgen05_pair_kernel (the conv GEMM launches of a forward)", "achieved": achieved_tflops,
                     "peak": peaks['bf16_sustained'], "unit": "TFLOP/s",
                     "frac": achieved_tflops / peaks['bf16_sustained'] if peaks['bf16_sustained'] else None,
                     "traffic": profiled_traffic(), "traffic_unit": "bytes of DRAM traffic per launch (ncu --set full, "
                     "profiles/r1_ncu_full_conv_gemm.json)", "peak_source": "%s bf16_tflops_sustained (kernel timed inside a long step)" % peaks['source'],
                     "avg_launch_ms": conv_ms_per_launch, "launches_per_step": conv_launches,
                     "algorithmic_gflop_per_step": conv_flops_step / 1e9,
                     "kernel_share_of_step": conv_ms / max(conv_ms + other_ms, 1e-9)},
        "whole_net_tflops": flops_img * value / world / 1e12,
    }
    if per_rank is not None:
        line["per_rank"] = per_rank
    with torch.no_grad():
        line["eval_pipeline"] = time_eval_pipeline(model, device, B, rank, world)
    if rank == 0:
        line["per_op_ms"] = {k: round(statistics.median(v), 4) for k, v in per_layer.items()}
        if world == 1:
            torch.manual_seed(0)
            dense = mc.Darknet(mc.write_yolov2_voc_cfg()).to(device).eval()
            line["mask"] = time_masks(dense, peaks)
            if not args.no_dense:
                with torch.no_grad():
                    for i in range(6):
                        dense(xs[i % 3])
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for i in range(20):
                        dense(xs[i % 3])
                    b.record()
                    torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / 20
                    dplan = compile_darknet(dense)
                    dper = {}
                    for i in range(5):
                        evs = []
                        dplan.run(xs[i % 3], events=evs)
                        torch.cuda.synchronize()
                        for op, e0_, e1_ in evs:
                            dper.setdefault(op['name'], []).append(e0_.elapsed_time(e1_))
                line["dense"] = {"images_per_s": B / (ms * 1e-3), "tflops": 29.36e9 * B / (ms * 1e-3) / 1e12,
                                 "ms_per_step": ms,
                                 "per_op_ms": {k: round(statistics.median(v), 4) for k, v in dper.items()}}
            del dense
            if not args.no_retrain:
                torch.cuda.empty_cache()
                line["retrain"] = time_retrain(device, peaks, B)
            if not args.no_cpu_baseline:
                state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
                ips, sps = cpu_forward_sample(state, model.blocks, CPU_SAMPLE_BATCH, 6, 1)
                line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(),
                                        "kind": "port", "sample": "6 forwards of batch %d of the same pruned network "
                                        "(oracle port of the reference PyTorch CPU path)" % CPU_SAMPLE_BATCH}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
