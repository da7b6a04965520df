"""GPU: small API-surface entry points of the reference that sit next to the hot path — stand-alone Reorg.forward
(src/nets.py:648-667), bbox_ious (src/nets2_utils.py:100-131), set_mask with a non-float32 / strided mask
(src/pruning/weightPruning/layers.py:41-47), and two training-mode forwards before one backward (the reference's
autograd graph keeps every forward's tensors alive)."""
import numpy as np
import pytest
import torch

import modelcompression_b200 as mc

pytestmark = pytest.mark.gpu


def _reorg_reference(x, stride):
    """The reference's view/transpose chain (src/nets.py:651-667), on the CPU."""
    B, C, H, W = x.shape
    hs = ws = stride
    x = x.view(B, C, H // hs, hs, W // ws, ws).transpose(3, 4).contiguous()
    x = x.view(B, C, H // hs * W // ws, hs * ws).transpose(2, 3).contiguous()
    x = x.view(B, C, hs * ws, H // hs, W // ws).transpose(1, 2).contiguous()
    return x.view(B, hs * ws * C, H // hs, W // ws)


@pytest.mark.parametrize("shape,stride", [((2, 64, 26, 26), 2), ((1, 3, 8, 12), 2), ((3, 5, 9, 6), 3)])
def test_reorg_standalone_equals_reference(shape, stride):
    torch.manual_seed(0)
    x = torch.randn(*shape)
    got = mc.Reorg(stride)(x.cuda()).cpu()
    assert torch.equal(got, _reorg_reference(x, stride))


def _bbox_ious_reference(boxes1, boxes2, x1y1x2y2=True):
    """src/nets2_utils.py:100-131 verbatim semantics on CPU float32 tensors."""
    if x1y1x2y2:
        mx = torch.min(boxes1[0], boxes2[0]); Mx = torch.max(boxes1[2], boxes2[2])
        my = torch.min(boxes1[1], boxes2[1]); My = torch.max(boxes1[3], boxes2[3])
        w1 = boxes1[2] - boxes1[0]; h1 = boxes1[3] - boxes1[1]
        w2 = boxes2[2] - boxes2[0]; h2 = boxes2[3] - boxes2[1]
    else:
        mx = torch.min(boxes1[0] - boxes1[2] / 2.0, boxes2[0] - boxes2[2] / 2.0)
        Mx = torch.max(boxes1[0] + boxes1[2] / 2.0, boxes2[0] + boxes2[2] / 2.0)
        my = torch.min(boxes1[1] - boxes1[3] / 2.0, boxes2[1] - boxes2[3] / 2.0)
        My = torch.max(boxes1[1] + boxes1[3] / 2.0, boxes2[1] + boxes2[3] / 2.0)
        w1 = boxes1[2]; h1 = boxes1[3]; w2 = boxes2[2]; h2 = boxes2[3]
    uw = Mx - mx
    uh = My - my
    cw = w1 + w2 - uw
    ch = h1 + h2 - uh
    mask = ((cw <= 0) + (ch <= 0) > 0)
    carea = cw * ch
    carea[mask] = 0
    uarea = w1 * h1 + w2 * h2 - carea
    return carea / uarea


@pytest.mark.parametrize("corners", [True, False])
def test_bbox_ious_bit_exact(corners):
    torch.manual_seed(3)
    n = 5000
    c = torch.rand(2, n)
    wh = torch.rand(2, n) * 0.4 + 0.01
    c2 = c + (torch.rand(2, n) - 0.5) * 0.6
    wh2 = torch.rand(2, n) * 0.4 + 0.01
    if corners:
        b1 = torch.cat([c - wh / 2, c + wh / 2])
        b2 = torch.cat([c2 - wh2 / 2, c2 + wh2 / 2])
    else:
        b1 = torch.cat([c, wh])
        b2 = torch.cat([c2, wh2])
    ref = _bbox_ious_reference(b1.clone(), b2.clone(), corners)
    got = mc.bbox_ious(b1.cuda(), b2.cuda(), corners).cpu()
    assert (ref == 0).any() and (ref > 0).any()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("kind", ["bool", "uint8", "float64", "strided"])
def test_set_mask_normalises_dtype_and_layout(kind):
    torch.manual_seed(1)
    conv = mc.MaskedConv2d(16, 24, 3, 1, 1, bias=False).cuda()
    w0 = conv.weight.data.clone()
    keep = torch.rand(24, 16, 3, 3) > 0.5
    if kind == "bool":
        m = keep
    elif kind == "uint8":
        m = keep.to(torch.uint8)
    elif kind == "float64":
        m = keep.double()
    else:
        m = keep.float().permute(1, 0, 2, 3).contiguous().permute(1, 0, 2, 3)  # right shape, wrong strides
        assert not m.is_contiguous()
    conv.set_mask(m)
    assert conv.mask.dtype == torch.float32 and conv.mask.is_contiguous() and conv.mask.is_cuda
    assert torch.equal(conv.mask.cpu(), keep.float())
    assert torch.equal(conv.weight.data.cpu(), w0.cpu() * keep.float())
    with pytest.raises(ValueError):
        conv.set_mask(torch.ones(3, 3))


def test_two_training_forwards_before_backward(cfg_path):
    """loss(model(x1)) + loss(model(x2)) -> one backward: each forward keeps its own saved activations (the advisor's
    round-1 finding: the second forward used to overwrite the first one's buffers silently)."""
    dev = torch.device('cuda')
    torch.manual_seed(0)
    model = mc.Darknet(cfg_path).to(dev)
    model.train()
    for m in model.modules():  # freeze running statistics: three forwards must see the same module state
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 0.0
    torch.manual_seed(1)
    x1 = torch.rand(2, 3, 416, 416, device=dev)
    x2 = torch.rand(2, 3, 416, 416, device=dev)
    g = torch.randn(2, 125, 13, 13, device=dev)

    def grads_of(fn):
        model.zero_grad(set_to_none=True)
        fn()
        return [p.grad.detach().clone() for p in model.parameters()]

    seen = {}

    def both():
        y1 = model(x1)
        with torch.no_grad():
            model(x2)  # a train-mode forward under no_grad between forward and backward must not disturb y1's graph
        y2 = model(x2)
        b1, b2 = y1.grad_fn.sv.bufs, y2.grad_fn.sv.bufs
        seen['distinct'] = all(b1[k].data_ptr() != b2[k].data_ptr() for k in b1)
        seen['z1'] = {k: b1[k].clone() for k in b1 if k.startswith('z') or k.startswith('a')}
        ((y1 * g).sum() + (y2 * g).sum()).backward()
        # y1's saved activations are still y1's after the later forwards and the backward of both graphs
        seen['intact'] = all(torch.equal(seen['z1'][k], b1[k]) for k in seen['z1'])

    gab = grads_of(both)
    assert seen['distinct'] and seen['intact']
    assert all(bool(torch.isfinite(t).all()) for t in gab)
    # the head bias gradient is sum(dy): exact, and the sum of both graphs' contributions
    assert torch.equal(gab[-1], 2 * g.sum(dim=(0, 2, 3)))
    # the buffer sets are recycled: further plain steps allocate nothing new
    plan = model.__dict__['_b200_train_plan']
    n_sets = sum(len(v) for v in plan._bufs.values())
    for _ in range(2):
        grads_of(lambda: (model(x1) * g).sum().backward())
    assert sum(len(v) for v in plan._bufs.values()) == n_sets == 2
    # a second backward through the same forward is refused (its buffers may already serve another forward)
    y = model(x1)
    (y * g).sum().backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="backward called twice"):
        (y * g).sum().backward()
