"""clock64 timeline of one CTA of conv_window_kernel (tuning build: make -C modelcompression_b200/csrc TUNING=1)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from modelcompression_b200 import _lib  # noqa: E402
import test_gpu_conv as T  # noqa: E402

lib = _lib.load()
DEV = 'cuda:0'
buf = torch.zeros(128, dtype=torch.int64, device=DEV)


def trace(tag, fn, slots):
    buf.zero_()
    fn()  # warm
    lib.mc_debug_window_trace(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    lib.mc_debug_window_trace(None)
    t = buf.cpu().view(8, 16)
    first = list(slots.keys())[0]
    t0 = int(t[0][first])
    print(tag)
    for i in range(8):
        print("  tile %d: " % (8 + i) + "  ".join("%s=%6d" % (n, int(t[i][s]) - t0) for s, n in slots.items()))


torch.manual_seed(0)
B = 64
x = torch.randn(B, 4, 208, 208, device=DEV)
w = torch.randn(1, 4, 3, 3, device=DEV)
sc, sh = torch.ones(1, device=DEV), torch.zeros(1, device=DEV)
P8 = {0: 'prod_free', 1: 'mma_in', 2: 'mma_tfree', 3: 'mma_done', 4: 'epi_top', 5: 'epi_full', 6: 'epi_drained', 7: 'epi_arr'}
trace("P8 conv2-shrunk (4->1, pool)", lambda: T._run_window(x, 0, w, sc, sh, 1, True), P8)
w17 = torch.randn(17, 1, 3, 3, device=DEV)
x1 = torch.randn(B, 1, 104, 104, device=DEV)
trace("P8 conv3-shrunk (1->17)", lambda: T._run_window(x1, 0, w17, torch.ones(17, device=DEV), torch.zeros(17, device=DEV), 1, False), P8)
IMG = {8: 'top', 9: 'pfree', 10: 'raw', 11: 'conv', 12: 'bar', 13: 'tfree', 14: 'issued', 4: 'epi_top', 5: 'epi_full', 6: 'epi_drained', 7: 'epi_arr'}
xu = torch.randint(0, 256, (B, 3, 416, 416), dtype=torch.uint8, device=DEV)
w4 = torch.randn(4, 3, 3, 3, device=DEV)
trace("IMG u8 conv1-shrunk (3->4)", lambda: T._run_window(xu, 2, w4, torch.ones(4, device=DEV), torch.zeros(4, device=DEV), 1, True), IMG)
w32 = torch.randn(32, 3, 3, 3, device=DEV)
trace("IMG u8 conv1-dense (3->32)", lambda: T._run_window(xu, 2, w32, torch.ones(32, device=DEV), torch.zeros(32, device=DEV), 1, True), IMG)
