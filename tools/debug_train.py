"""Per-block comparison of the training-mode forward with the fp32 oracle (debug aid)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import modelcompression_b200 as mc
from modelcompression_b200 import _lib
from modelcompression_b200.engine_train import TrainPlan, _forward
from conftest import make_darknet
from oracle import train_oracle

dev = 'cuda:0'
model = make_darknet(mc.write_yolov2_voc_cfg(), seed=0, kn=True, randbn=True, device=dev)
model.set_masks(mc.weight_prune(model, 90.))
model.train()
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(1)
x = torch.rand(2, 3, 416, 416).to(dev)
state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
params = {k: v.clone().float() for k, v in state0.items() if k.endswith('.weight') or k.endswith('.bias')}
buffers = {k: v.clone() for k, v in state0.items() if k not in params}
outs = {}
with torch.no_grad():
    y_o = train_oracle.train_forward_fp32(model.blocks, params, buffers, x, outputs_out=outs, emulate_bf16=('fp32' not in sys.argv))
torch.backends.cudnn.allow_tf32 = False
plan = TrainPlan(model)
with torch.cuda.device(0), torch.no_grad():
    y, sv = _forward(plan, x, training_stats=False)
    lib = _lib.load()
    for L in plan.layers:
        if L.is_head:
            continue
        a = L.act
        B = 2
        if L.reorg:
            continue
        t = sv.bufs[a.name]
        got = torch.empty(B, L.O, a.H, a.W, device=dev)
        _lib.check(lib.mc_unpack_pnhwc(t.data_ptr(), got.data_ptr(), B, a.H, a.W, L.O, a.ld, a.ch_off, _lib.stream_ptr()), "unpack")
        ref = outs[L.ind]
        print("block %2d a: rel-L2 %.4f  max|ref| %.3f" % (L.ind, ((got - ref).norm() / ref.norm()).item(), ref.abs().max().item()))
    print("logits rel-L2 %.4f" % ((y - y_o).norm() / y_o.norm()).item())
