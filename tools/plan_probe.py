import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200 import _lib
lib = _lib.load()
for (B, C, H, W, O) in [(32, 1006, 13, 13, 1018), (40, 256, 13, 13, 768), (8, 256, 26, 26, 1024), (32, 512, 13, 13, 1024),
                        (24, 512, 13, 13, 1024), (48, 512, 13, 13, 1024), (40, 512, 13, 13, 768), (16, 512, 26, 26, 1024)]:
    conv = mc.MaskedConv2d(C, O, 3, 1, 1, bias=True).cuda()
    x = torch.randn(B, C, H, W, device='cuda')
    conv(x)
    info = (ctypes.c_int * 8)()
    lib.mc_conv_last_plan(info)
    rows = B * (H + 1) * (W + 1)
    print((B, C, H, W, O), "m_tiles", (rows + 127) // 128, "plan", list(info))
