"""Diagnostic (not a test): per-layer error of the tcgen05 conv kernel vs fp32 math; prints instead of asserting."""
import sys
import os
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200.engine import compile_darknet
from oracle import forward_oracle

dev = 'cuda:0'
bf = lambda t: t.to(torch.bfloat16).float()
for (B, C, H, W, O, k) in [(1, 64, 8, 8, 16, 1), (2, 64, 13, 13, 128, 1), (2, 128, 26, 26, 256, 3), (1, 32, 52, 52, 64, 3),
                           (2, 1280, 13, 13, 1024, 3), (3, 24, 16, 20, 40, 3)]:
    torch.manual_seed(0)
    conv = mc.MaskedConv2d(C, O, k, 1, (k - 1) // 2, bias=False).to(dev)
    x = torch.randn(B, C, H, W, device=dev)
    try:
        y = conv(x)
        torch.cuda.synchronize()
        ref = F.conv2d(bf(x), bf(conv.weight.data), None, 1, (k - 1) // 2)
        err = (y - ref).abs()
        print("conv B%d C%d %dx%d O%d k%d: max err %.4g (ref max %.4g) rel %.3g; frac bad %.4f" % (
            B, C, H, W, O, k, err.max().item(), ref.abs().max().item(), (err.max() / ref.abs().max()).item(),
            (err > 1e-2 * ref.abs().max()).float().mean().item()), flush=True)
        if (err.max() / ref.abs().max()).item() > 1e-2:
            bad = (err > 1e-2 * ref.abs().max()).nonzero()
            print("   first bad idx:", bad[:5].tolist(), "y", y[tuple(bad[0].tolist())].item(), "ref", ref[tuple(bad[0].tolist())].item())
            print("   y[0,0,:2,:6]", y[0, 0, :2, :6].tolist())
            print("   r[0,0,:2,:6]", ref[0, 0, :2, :6].tolist())
    except Exception as e:  # noqa
        print("conv B%d C%d %dx%d O%d k%d FAILED: %r" % (B, C, H, W, O, k, e), flush=True)
        break

cfg = mc.write_yolov2_voc_cfg()
torch.manual_seed(0)
model = mc.Darknet(cfg)
forward_oracle.kaiming_normal_init_(model, 7)
model = model.to(dev).eval()
torch.manual_seed(1)
x = torch.rand(2, 3, 416, 416, device=dev)
with torch.no_grad():
    y = model(x)
    torch.cuda.synchronize()
    y_ref, outs = forward_oracle.darknet_forward_fp32(model.blocks, model.state_dict(), x, keep_outputs=True)
plan = compile_darknet(model)
for ind in sorted(plan.block_out):
    if ind in outs:
        got = plan.block_activation(ind)
        want = outs[ind]
        if got.shape != want.shape:
            print("block %d shape %s vs %s (fused)" % (ind, tuple(got.shape), tuple(want.shape)))
            continue
        print("block %2d %-14s max-rel %.3g  l2-rel %.3g" % (ind, model.blocks[ind + 1]['type'],
              ((got - want).abs().max() / want.abs().max()).item(), ((got - want).norm() / want.norm()).item()), flush=True)
print("head max-rel %.3g l2-rel %.3g" % (((y - y_ref).abs().max() / y_ref.abs().max()).item(),
                                         ((y - y_ref).norm() / y_ref.norm()).item()))
