"""Time mc_conv_wgrad alone on given shapes (B C O H W k ...)."""
import ctypes, os, sys, statistics
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200 import _lib
dev = 'cuda:0'
lib = _lib.load()
args = [int(a) for a in sys.argv[1:]]
for i in range(0, len(args), 6):
    B, C, O, H, W, k = args[i:i + 6]
    rows = B * (H + 1) * (W + 1)
    lda, ldz = (C + 7) // 8 * 8, (O + 7) // 8 * 8
    a = torch.randn(rows, lda, device=dev).to(torch.bfloat16)
    dz = torch.randn(rows, ldz, device=dev).to(torch.bfloat16)
    dw = torch.empty(O, C, k, k, device=dev)
    nbytes = lib.mc_workspace_bytes_conv_wgrad(B, H, W, C, O, k)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    s = _lib.stream_ptr()
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _lib.check(lib.mc_conv_wgrad(a.data_ptr(), lda, C, dz.data_ptr(), ldz, O, B, H, W, k, None, dw.data_ptr(), 0,
                                         ws.data_ptr(), nbytes, s), "wgrad")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 3 * 1e3)
    us = statistics.median(ts[2:])
    fl = 2.0 * B * H * W * C * O * k * k
    print("wgrad B%d %d->%d %dx%d k%d: %.1f us (incl. reduce), %.0f TFLOP/s, operands %.0f MB, ws %.0f MB" %
          (B, C, O, H, W, k, us, fl / us / 1e6, (a.numel() + dz.numel()) * 2 / 1e6, nbytes / 1e6))
