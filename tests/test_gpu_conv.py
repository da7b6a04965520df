"""GPU parity: the tcgen05 conv stack vs fp32 PyTorch math on the same weights (per layer, per block, whole network),
dense / weight-pruned / physically shrunk, plus the golden heads produced by the UNMODIFIED reference on CPU.

Tolerance (stated once): activations and weights are bf16 with fp32 accumulation.
  * single layer vs fp32 conv on bf16-ROUNDED operands: only the bf16 output rounding differs ->
    |err| <= 5e-3*|ref| + 2e-3*max|ref|  (measured: max-rel 3e-3 on B200).
  * whole network vs the pure fp32 reference: every layer re-rounds activations and weights to bf16 (relative rms
    ~2.3e-3 per layer), which accumulates as sqrt(#layers): measured on B200 the relative L2 error grows from 2e-3
    (block 1) to 1.28e-2 at the logits after 23 convs.  Gates: per block max|err|/max|ref| <= 3e-2, logits relative
    L2 <= 2e-2 and max-rel <= 3e-2.  (north_star's "e.g. max rel err <= 1e-2" is met per layer, not end to end.)
Default-init logits equal conv23.bias to ~3e-6, so the variance-preserving KN init is the meaningful case
(SURVEY.md §7 hard part 7)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import modelcompression_b200 as mc
from conftest import load_golden, make_darknet
from modelcompression_b200.engine import compile_darknet
from oracle import forward_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _bf16(t):
    return t.to(torch.bfloat16).float()


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("B,C,H,W,O,k,bias", [
    (2, 64, 13, 13, 128, 1, False),     # one k-block, 1x1
    (2, 128, 26, 26, 256, 3, False),    # 3x3, two k-blocks per tap
    (1, 32, 52, 52, 64, 3, True),       # Cin < 64: TMA zero-fills the k-block
    (2, 1280, 13, 13, 1024, 3, False),  # conv22 shape: 180 k-blocks, 4 N tiles
    (3, 24, 16, 20, 40, 3, True),       # ragged: Cin, O not multiples of 16/64, non-square image
    (1, 1024, 13, 13, 125, 1, True),    # head shape
    (2, 8, 8, 8, 16, 3, False),         # tiny
    (1, 72, 30, 14, 200, 1, False),
])
def test_single_conv_layer(B, C, H, W, O, k, bias):
    torch.manual_seed(B * 1000 + C + O)
    conv = mc.MaskedConv2d(C, O, k, 1, (k - 1) // 2, bias=bias).to(DEV)
    x = torch.randn(B, C, H, W, device=DEV)
    y = conv(x)
    ref = F.conv2d(_bf16(x), _bf16(conv.weight.data), conv.bias.data if bias else None, 1, (k - 1) // 2)
    assert y.shape == ref.shape
    # output is rounded to bf16 once: 2^-8 relative per element
    err = (y - ref).abs()
    assert (err <= 5e-3 * ref.abs() + 2e-3 * ref.abs().max()).all(), "max rel %.3g" % _rel(y, ref)
    # masked variant: forward uses weight * mask (layers.py:59)
    mask = (torch.rand_like(conv.weight) > 0.6).float()
    w0 = conv.weight.data.clone()
    conv.set_mask(mask)
    assert torch.equal(conv.weight.data, w0 * mask) and conv.mask_flag
    y2 = conv(x)
    ref2 = F.conv2d(_bf16(x), _bf16(w0 * mask), conv.bias.data if bias else None, 1, (k - 1) // 2)
    err2 = (y2 - ref2).abs()
    assert (err2 <= 5e-3 * ref2.abs() + 2e-3 * ref2.abs().max()).all()


def _last_plan():
    import ctypes
    from modelcompression_b200 import _lib
    info = (ctypes.c_int * 8)()
    _lib.check(_lib.load().mc_conv_last_plan(info), "mc_conv_last_plan")
    return dict(pair=info[0], block_n=info[1], ctas=info[2], resident=info[3], share=info[4], stages=info[5],
                grid=info[6], block_k=info[7])


@pytest.mark.parametrize("B,C,H,W,O,k,want", [
    (64, 512, 13, 13, 1024, 3, dict(pair=1)),            # wide 3x3 at the bench batch: CTA-pair kernel (cta_group::2)
    (64, 1006, 13, 13, 1018, 3, dict(pair=1)),           # shrunk-net shape: ragged Cin/N, odd channel-block tail
    (33, 600, 13, 13, 512, 3, dict(pair=1)),             # odd number of 128-row tiles (51): last pair half-empty
    (8, 32, 208, 208, 64, 3, dict(share=1, pair=0, block_k=32)),  # long launch, 32-wide k-blocks: shared box, 64 B swizzle
    (8, 32, 104, 104, 64, 3, dict(share=0, block_k=32)),  # same layer, short launch: one box per tap
    (8, 40, 52, 52, 72, 3, dict(share=1, block_k=64)),    # Cin 40 in a 64-wide k-block: 3 of 4 K steps issued
    (8, 16, 52, 52, 72, 3, dict(block_k=32)),
    (16, 24, 104, 104, 8, 1, dict(pair=0, resident=0)),  # narrow 1x1: several CTAs per SM
    (32, 80, 52, 52, 16, 1, dict(pair=0)),
])
def test_conv_launch_paths(B, C, H, W, O, k, want):
    """Every launch configuration of mc_conv_fwd (CTA pair, shared activation box with skipped K steps, multi-CTA narrow layers) against fp32
    conv on bf16-rounded operands, with the chosen configuration read back from the library."""
    torch.manual_seed(C * 7 + O)
    conv = mc.MaskedConv2d(C, O, k, 1, (k - 1) // 2, bias=True).to(DEV)
    x = torch.randn(B, C, H, W, device=DEV)
    y = conv(x)
    plan = _last_plan()
    for key, val in want.items():
        assert plan[key] == val, (plan, want)
    if k == 1 and O <= 16:
        assert plan['ctas'] >= 2, plan
    ref = F.conv2d(_bf16(x), _bf16(conv.weight.data), conv.bias.data, 1, (k - 1) // 2)
    err = (y - ref).abs()
    assert (err <= 5e-3 * ref.abs() + 2e-3 * ref.abs().max()).all(), "max rel %.3g (%s)" % (_rel(y, ref), plan)


def _check_blocks(model, x, tag, tol_block=3e-2, tol_head_l2=2e-2):
    with torch.no_grad():
        y = model(x)
        y_ref, outs = forward_oracle.darknet_forward_fp32(model.blocks, model.state_dict(), x, keep_outputs=True)
    plan = compile_darknet(model)
    report = []
    for ind in sorted(plan.block_out):
        if ind not in outs:
            continue
        got = plan.block_activation(ind)
        want = outs[ind]
        if got.shape != want.shape:
            # a conv fused with the following maxpool / reorg block only materialises the fused result, which is
            # checked under the following block's index
            assert model.blocks[ind + 2]['type'] in ('maxpool', 'reorg'), (tag, ind, got.shape, want.shape)
            assert plan.block_out[ind + 1] is plan.block_out[ind]
            continue
        report.append((ind, _rel(got, want), _rel_l2(got, want)))
    worst = max(report, key=lambda r: r[1])
    print("[%s] worst block %d: max-rel %.3g, l2-rel %.3g; head max-rel %.3g l2-rel %.3g" %
          (tag, worst[0], worst[1], worst[2], _rel(y, y_ref), _rel_l2(y, y_ref)))
    for ind, r, l2 in report:
        assert r <= tol_block, "%s: block %d max-rel err %.3g" % (tag, ind, r)
    assert y.shape == y_ref.shape == (x.shape[0], 125, 13, 13)
    assert _rel_l2(y, y_ref) <= tol_head_l2, "%s: head l2-rel err %.3g" % (tag, _rel_l2(y, y_ref))
    assert _rel(y, y_ref) <= 3e-2, "%s: head max-rel err %.3g" % (tag, _rel(y, y_ref))
    return y, y_ref


def test_dense_network_per_block_kn(cfg_path):
    model = make_darknet(cfg_path, seed=0, kn=True, device=DEV)
    torch.manual_seed(1)
    x = torch.rand(2, 3, 416, 416).to(DEV)
    y, _ = _check_blocks(model, x, 'dense-kn')
    g = load_golden('forward.npz')
    # image 0 of this batch is the golden image (same generator stream prefix): compare with the reference's CPU head
    torch.manual_seed(1)
    x1 = torch.rand(1, 3, 416, 416).to(DEV)
    with torch.no_grad():
        y1 = model(x1)
    want = torch.from_numpy(g['kn_head']).to(DEV)
    assert _rel_l2(y1, want) <= 2e-2 and _rel(y1, want) <= 3e-2


def test_dense_network_randbn_and_default_init(cfg_path):
    g = load_golden('forward.npz')
    torch.manual_seed(1)
    x1 = torch.rand(1, 3, 416, 416).to(DEV)
    model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=DEV)
    y, _ = _check_blocks(model, x1, 'dense-kn-randbn')
    want = torch.from_numpy(g['kn_randbn_head']).to(DEV)
    assert _rel_l2(y, want) <= 2e-2
    # default init: logits == conv23.bias +- 3e-6; absolute agreement is all that can be asked
    model = make_darknet(cfg_path, seed=0, device=DEV)
    with torch.no_grad():
        y = model(x1)
    want = torch.from_numpy(g['default_head']).to(DEV)
    assert (y - want).abs().max().item() <= 1e-4


def test_weight_pruned_network(cfg_path):
    g = load_golden('forward.npz')
    model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=DEV)
    model.set_masks(mc.weight_prune(model, 70.))
    torch.manual_seed(1)
    x1 = torch.rand(1, 3, 416, 416).to(DEV)
    y, _ = _check_blocks(model, x1, 'w70')
    want = torch.from_numpy(g['kn_randbn_w70_head']).to(DEV)
    assert _rel_l2(y, want) <= 2e-2


def test_filter_pruned_network_physically_shrunk(cfg_path):
    g = load_golden('forward.npz')
    model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=DEV)
    masks, keep = mc.quick_filter_prune(model, 40., return_keep=True)
    model.set_masks(masks)
    torch.manual_seed(1)
    x1 = torch.rand(1, 3, 416, 416).to(DEV)
    # shrunk: filters removed, constants of removed channels folded through the ones channel (rand-BN => non-zero)
    model.b200_shrink = True
    y_s, y_ref = _check_blocks(model, x1, 'f40-shrunk')
    plan = compile_darknet(model)
    convs = [op for op in plan.ops if op['kind'] in ('conv', 'direct', 'im2col')]
    kept = [int(k.numel()) for k in keep]
    # every non-head layer lost filters physically (+1 for the ones channel where constants are non-zero)
    assert all(op['N'] <= n + 1 for op, n in zip(convs[:-1], kept[:-1]))
    assert convs[-1]['N'] == 125  # the head keeps all outputs: pruned ones are bias-only (SURVEY.md §7 hard part 4)
    assert plan.flops_per_image < 0.6 * 29.36e9
    # un-shrunk (masked, dense shapes) gives the same logits
    model.b200_shrink = False
    with torch.no_grad():
        y_d = model(x1)
    assert compile_darknet(model).flops_per_image == pytest.approx(29.36e9, rel=1e-3)
    assert _rel_l2(y_s, y_d) <= 2e-2
    want = torch.from_numpy(g['kn_randbn_f40_head']).to(DEV)
    assert _rel_l2(y_d, want) <= 2e-2 and _rel_l2(y_s, want) <= 2e-2


def test_plan_invalidation_and_batch_sizes(cfg_path):
    model = make_darknet(cfg_path, seed=0, kn=True, device=DEV)
    torch.manual_seed(2)
    x = torch.rand(3, 3, 416, 416).to(DEV)
    with torch.no_grad():
        y3 = model(x)
        y1 = model(x[1:2].contiguous())
    assert torch.allclose(y3[1:2], y1, rtol=0, atol=0)  # batch-size independent, deterministic
    plan_a = compile_darknet(model)
    with torch.no_grad():
        model.models[30][0].bias.add_(1.0)  # in-place update bumps the version counter -> re-pack
        y1b = model(x[1:2].contiguous())
    assert compile_darknet(model) is not plan_a
    assert torch.allclose(y1b, y1 + 1.0, atol=1e-5)
    model.train()  # training mode is the batch-statistics path of engine_train.py (tests/test_gpu_train.py)
    yt = model(x)
    assert yt.requires_grad and yt.shape == y3.shape
    model.eval()


def test_uint8_image_input_matches_float_path(cfg_path):
    """uint8 NCHW images (do_detect's input, src/nets2_utils.py:346-352) are scaled by 1/255 inside the first-layer
    kernel: same bf16 operands as feeding img.float().div(255.0), so the head is bit-identical."""
    for shrink in (True, False):
        model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=DEV)
        if shrink:
            model.set_masks(mc.quick_filter_prune(model, 40.))
        torch.manual_seed(4)
        xu = torch.randint(0, 256, (3, 3, 416, 416), dtype=torch.uint8, device=DEV)
        with torch.no_grad():
            for _ in range(3):  # eager, then graph capture, then replay
                yu = model(xu)
            yf = model(xu.float().div(255.0))
        assert yu.dtype == torch.float32 and torch.equal(yu, yf)


@pytest.mark.parametrize("C,O,k", [(32, 64, 3), (69, 145, 3), (17, 40, 1), (96, 32, 3)])
def test_conv_k_block_32(C, O, k):
    """mc_conv_fwd with block_k = 32 (64-byte swizzle, Kc = round_up(Cin, 32)) equals the 64-wide k-block path."""
    import ctypes
    from modelcompression_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(C)
    B, H, W = 2, 13, 21
    x = torch.randn(B, C, H, W, device=DEV)
    w = torch.randn(O, C, k, k, device=DEV) * 0.1
    ld_in, Npad, ld_out = (C + 7) // 8 * 8, (O + 15) // 16 * 16, (O + 7) // 8 * 8
    outs = []
    with torch.cuda.device(0):
        s = _lib.stream_ptr()
        xin = torch.empty(B * (H + 1) * (W + 1), ld_in, dtype=torch.bfloat16, device=DEV)
        _lib.check(lib.mc_pack_pnhwc(x.data_ptr(), xin.data_ptr(), B, H, W, C, ld_in, s), "pack")
        scale, shift = torch.ones(Npad, device=DEV), torch.zeros(Npad, device=DEV)
        for kb in (64, 32):
            Kc = (C + kb - 1) // kb * kb
            wpack = torch.empty(Npad, k * k * Kc, dtype=torch.bfloat16, device=DEV)
            _lib.check(lib.mc_pack_conv_weights(w.data_ptr(), None, O, C, k, None, O, None, C, wpack.data_ptr(), Npad, Kc, s), "packw")
            yb = torch.zeros(B * (H + 1) * (W + 1), ld_out, dtype=torch.bfloat16, device=DEV)
            d = _lib.mc_conv_desc()
            d.d_in, d.d_wpack, d.d_scale, d.d_shift, d.d_out = xin.data_ptr(), wpack.data_ptr(), scale.data_ptr(), shift.data_ptr(), yb.data_ptr()
            d.B, d.H, d.W, d.Cin, d.Cin_ld, d.N, d.Npad = B, H, W, C, ld_in, O, Npad
            d.ksize, d.leaky, d.epi_mode, d.ldc, d.ch_off, d.block_n, d.stages, d.block_k = k, 1, _lib.MC_EPI_PNHWC, ld_out, 0, 0, 0, kb
            _lib.check(lib.mc_conv_fwd(ctypes.byref(d), s), "conv")
            y = torch.empty(B, O, H, W, device=DEV)
            _lib.check(lib.mc_unpack_pnhwc(yb.data_ptr(), y.data_ptr(), B, H, W, O, ld_out, 0, s), "unpack")
            outs.append(y)
    ref = F.leaky_relu(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), padding=(k - 1) // 2), 0.1)
    assert (outs[0] - ref).abs().max() <= 5e-3 * ref.abs().max() + 2e-3
    assert (outs[1] - ref).abs().max() <= 5e-3 * ref.abs().max() + 2e-3
