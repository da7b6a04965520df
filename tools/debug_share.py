"""Which input shift does each 3x3 tap read?  One-hot filters, identity over channels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200.engine import single_conv_forward
from modelcompression_b200.pruning.weightPruning.layers import MaskedConv2d
dev = 'cuda:0'
torch.manual_seed(0)
C = 64
x = torch.randn(1, C, 12, 20, device=dev).bfloat16().float()
for r in range(3):
    for s in range(3):
        conv = MaskedConv2d(C, C, 3, padding=1, bias=False).to(dev)
        with torch.no_grad():
            conv.weight.zero_()
            for c in range(C):
                conv.weight[c, c, r, s] = 1.0
        y = single_conv_forward(conv, x)
        ref = torch.nn.functional.conv2d(x, conv.weight, padding=1)
        best = None
        for dy in (-2, -1, 0, 1, 2):
            for dx in (-3, -2, -1, 0, 1, 2, 3):
                sh = torch.zeros_like(x)
                ys0, ys1 = max(0, -dy), min(12, 12 - dy)
                xs0, xs1 = max(0, -dx), min(20, 20 - dx)
                sh[:, :, ys0:ys1, xs0:xs1] = x[:, :, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx]
                e = (y[:, :, 2:-2, 3:-3] - sh[:, :, 2:-2, 3:-3]).abs().max().item()
                if best is None or e < best[0]:
                    best = (e, dy, dx)
        print("tap (r=%d,s=%d): expected shift (%d,%d); output matches shift (dy=%d,dx=%d) with err %.3g; err vs ref %.3g"
              % (r, s, r - 1, s - 1, best[1], best[2], best[0], (y - ref).abs().max().item()))
