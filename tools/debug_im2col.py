import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import modelcompression_b200 as mc
from modelcompression_b200.engine import compile_darknet
from oracle import forward_oracle
dev = 'cuda:0'
torch.manual_seed(0)
model = mc.Darknet(mc.write_yolov2_voc_cfg())
forward_oracle.kaiming_normal_init_(model, 7)
model = model.to(dev).eval()
plan = compile_darknet(model); plan.use_graph = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = torch.rand(B, 3, 416, 416, device=dev)
# run only the first op
plan.ops = plan.ops[:1]
print(plan.ops[0]['name'], flush=True)
try:
    plan._run_eager(x)
    torch.cuda.synchronize()
    print("first op ok; timeout code: 0x%x" % mc._lib.load().mc_debug_im2col_timeout(), flush=True)
    out = plan._last_bufs[plan.ops[0]['dst_buf']]
    print("out absmax", out.float().abs().max().item(), "nan", bool(out.float().isnan().any()))
except Exception as e:
    print("FAILED", repr(e)[:300], flush=True)
