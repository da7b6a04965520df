"""GPU parity: region decode (tolerance contract on transcendentals, exact candidate sets away from the threshold)
and NMS (bit-exact kept-index lists, stable tie rule) vs the reference's golden vectors and the oracle."""
import numpy as np
import pytest
import torch

import modelcompression_b200 as mc
from conftest import load_golden
from modelcompression_b200.cfg import VOC_ANCHORS
from modelcompression_b200.nets2_utils import decode_device, nms_device
from oracle import detect_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
RTOL = 2e-5  # decode tolerance: expf/sigmoid/softmax differ from torch by a few ulp


def _cases():
    g = load_golden('detect.npz')
    return g, [str(n) for n in g['names']]


def test_decode_matches_reference_vectors():
    g, names = _cases()
    for name in names:
        logits = torch.from_numpy(g[name + '_logits']).to(DEV)
        T, oo, val, nt = g[name + '_cfg']
        boxes, counts, cls = decode_device(logits, T, 20, VOC_ANCHORS, 5, int(oo), want_cls=True)
        for b in range(logits.shape[0]):
            want = g['%s_%d_box' % (name, b)]
            n = int(counts[b])
            got = boxes[b, :n].cpu().numpy()
            assert n == want.shape[0], (name, b)
            assert np.array_equal(got[:, 7].astype(np.int32), g['%s_%d_pos' % (name, b)])
            assert np.array_equal(got[:, 6], want[:, 6])
            np.testing.assert_allclose(got[:, :6], want[:, :6], rtol=RTOL, atol=1e-7)
            probs = cls[b, :n].cpu().numpy()
            np.testing.assert_allclose(probs.sum(axis=1), 1.0, rtol=1e-5)
            np.testing.assert_allclose(probs.max(axis=1), want[:, 5], rtol=RTOL)


def test_nms_bit_exact_on_reference_boxes():
    # stage-isolated: feed the reference's own decoded boxes, compare kept-index lists exactly
    g, names = _cases()
    for name in names:
        nt = float(g[name + '_cfg'][3])
        nimg = g[name + '_logits'].shape[0]
        P = 845
        boxes = torch.zeros(nimg, P, 8)
        counts = torch.zeros(nimg, dtype=torch.int32)
        for b in range(nimg):
            bx = g['%s_%d_box' % (name, b)]
            boxes[b, :bx.shape[0], :7] = torch.from_numpy(bx)
            counts[b] = bx.shape[0]
        boxes, counts = boxes.to(DEV), counts.to(DEV)
        keep, kc = nms_device(boxes, counts, nt)
        for b in range(nimg):
            want = g['%s_%d_keep' % (name, b)].tolist()
            assert keep[b, :int(kc[b])].cpu().tolist() == want, (name, b)
            n = int(counts[b])
            assert np.array_equal(boxes[b, :n, 4].cpu().numpy(), g['%s_%d_conf_after' % (name, b)])


def test_list_api_matches_reference_structure():
    g, _ = _cases()
    name = 'n2_val'
    logits = torch.from_numpy(g[name + '_logits']).to(DEV)
    T, oo, val, nt = g[name + '_cfg']
    allb = mc.get_region_boxes(logits, T, 20, VOC_ANCHORS, 5, int(oo), bool(val))
    assert len(allb) == logits.shape[0]
    for b, boxes in enumerate(allb):
        want = g['%s_%d_box' % (name, b)]
        assert len(boxes) == want.shape[0]
        assert all(torch.is_tensor(v) and v.dim() == 0 for v in boxes[0][:7])
        assert boxes[0][6].dtype == torch.int64 and boxes[0][0].dtype == torch.float32
        extras = g['%s_%d_extras' % (name, b)]
        got_extras = [(r, int(bx[7 + j + 1]), float(bx[7 + j])) for r, bx in enumerate(boxes)
                      for j in range(0, len(bx) - 7, 2)]
        assert [(e[0], e[1]) for e in got_extras] == [(int(e[0]), int(e[1])) for e in extras]
        np.testing.assert_allclose([e[2] for e in got_extras], extras[:, 2], rtol=RTOL)
        # nms on the list: same kept objects (by identity), suppressed boxes get box[4] = 0
        ref_boxes = [[torch.tensor(v) for v in row[:6]] + [torch.tensor(int(row[6]))] for row in want]
        kept = mc.nms(ref_boxes, float(nt))
        ids = {id(bx): i for i, bx in enumerate(ref_boxes)}
        assert [ids[id(bx)] for bx in kept] == g['%s_%d_keep' % (name, b)].tolist()
        after = np.array([float(bx[4]) for bx in ref_boxes], dtype=np.float32)
        assert np.array_equal(after, g['%s_%d_conf_after' % (name, b)])
    assert mc.nms([], 0.45) == []


def test_decode_nms_vs_oracle_random_batches():
    torch.manual_seed(5)
    for (B, T, oo, nt, scale) in [(8, 0.5, 1, 0.45, 2.0), (3, 0.005, 0, 0.4, 1.0), (2, 0.9999, 1, 0.45, 2.0)]:
        logits = torch.randn(B, 125, 13, 13) * scale
        dec = detect_oracle.decode_np(logits, T, 20, VOC_ANCHORS, 5, oo)
        boxes, counts, _ = decode_device(logits.to(DEV), T, 20, VOC_ANCHORS, 5, oo)
        decoded = boxes.clone()
        keep, kc = nms_device(boxes, counts, nt)
        for b in range(B):
            n = int(counts[b])
            got = decoded[b, :n].cpu().numpy()
            # the candidate set may differ only for scores within a few ulp of the threshold
            if n != dec[b]['box'].shape[0]:
                pytest.fail("candidate count %d vs %d" % (n, dec[b]['box'].shape[0]))
            np.testing.assert_allclose(got[:, :6], dec[b]['box'][:, :6], rtol=RTOL, atol=1e-7)
            ko, conf_o = detect_oracle.nms_np(got[:, :5], nt)  # oracle NMS on the GPU-decoded boxes: exact
            assert keep[b, :int(kc[b])].cpu().tolist() == ko
            assert np.array_equal(boxes[b, :n, 4].cpu().numpy(), conf_o)


def test_nms_edge_cases_device():
    boxes = torch.zeros(3, 16, 8, device=DEV)
    counts = torch.tensor([0, 1, 4], dtype=torch.int32, device=DEV)
    boxes[1, 0, :5] = torch.tensor([0.5, 0.5, 0.1, 0.1, 0.9])
    boxes[2, :4, :5] = torch.tensor([0.5, 0.5, 0.1, 0.1, 0.9])  # identical boxes, tied keys
    keep, kc = nms_device(boxes, counts, 0.45)
    assert kc.cpu().tolist() == [0, 1, 1]
    assert keep[1, 0].item() == 0 and keep[2, 0].item() == 0
    assert boxes[2, :4, 4].cpu().tolist() == [pytest.approx(0.9), 0, 0, 0]
    # non power-of-two capacity, other grid sizes (19x19x5 = 1805 candidates)
    torch.manual_seed(1)
    logits = torch.randn(2, 125, 19, 19) * 2
    dec = detect_oracle.decode_np(logits, 0.3, 20, VOC_ANCHORS, 5, 1)
    b2, c2, _ = decode_device(logits.to(DEV), 0.3, 20, VOC_ANCHORS, 5, 1)
    dcopy = b2.clone()
    k2, kc2 = nms_device(b2, c2, 0.45)
    for b in range(2):
        n = int(c2[b])
        assert n == dec[b]['box'].shape[0]
        ko, _ = detect_oracle.nms_np(dcopy[b, :n, :5].cpu().numpy(), 0.45)
        assert k2[b, :int(kc2[b])].cpu().tolist() == ko


# ------------------------------------------------------------------------------------------------ fused decode epilogue
def _fused_decode_from_logits(logits, thr, only_obj, want_cls, want_head=True):
    """Run the head-convolution kernel with the region decode fused into its epilogue (MC_EPI_DECODE) on GIVEN logits:
    a 1x1 'selection' convolution whose fp32 accumulators equal `logits` bit for bit.  Every logit is split into three
    bf16 pieces (hi + mid + lo == logit exactly, non-overlapping mantissas) that sit in three input channels with weight
    1, so the tensor core's fp32 accumulation reproduces the value exactly.  Returns (boxes, cls, head)."""
    import ctypes
    from modelcompression_b200 import _lib
    lib = _lib.load()
    B, N, H, W = logits.shape
    x = logits.float()
    hi = x.to(torch.bfloat16).float()
    r1 = x - hi
    mid = r1.to(torch.bfloat16).float()
    lo = (r1 - mid).to(torch.bfloat16).float()
    assert torch.equal((hi + mid) + lo, x), "three-way bf16 split is not exact for these logits"
    src = torch.cat([hi, mid, lo], dim=1).contiguous()  # [B, 3N, H, W]
    C = 3 * N
    ld_in = (C + 63) // 64 * 64
    Npad = (N + 15) // 16 * 16
    xin = torch.zeros(B * (H + 1) * (W + 1), ld_in, dtype=torch.bfloat16, device=DEV)
    wsel = torch.zeros(N, C, 1, 1, device=DEV)
    for part in range(3):
        wsel[torch.arange(N), part * N + torch.arange(N), 0, 0] = 1.0
    wpack = torch.empty(Npad, ld_in, dtype=torch.bfloat16, device=DEV)
    scale = torch.zeros(Npad, device=DEV)
    scale[:N] = 1
    shift = torch.zeros(Npad, device=DEV)
    P = H * W * 5
    boxes = torch.full((B, P, 8), float('nan'), device=DEV)
    cls = torch.full((B, P, 20), float('nan'), device=DEV) if want_cls else None
    head = torch.empty(B, N, H, W, device=DEV) if want_head else None
    s = _lib.stream_ptr()
    _lib.check(lib.mc_pack_pnhwc(src.data_ptr(), xin.data_ptr(), B, H, W, C, ld_in, s), "mc_pack_pnhwc")
    _lib.check(lib.mc_pack_conv_weights(wsel.data_ptr(), None, N, C, 1, None, N, None, C, wpack.data_ptr(), Npad, ld_in,
                                        s), "mc_pack_conv_weights")
    dec = _lib.mc_decode_params()
    dec.d_boxes, dec.d_cls = boxes.data_ptr(), (cls.data_ptr() if cls is not None else None)
    dec.d_head = head.data_ptr() if head is not None else None
    dec.A, dec.nc, dec.conf_thresh, dec.only_objectness = 5, 20, float(thr), int(only_obj)
    for i, a in enumerate(VOC_ANCHORS):
        dec.anchors[i] = float(a)
    d = _lib.mc_conv_desc()
    d.d_in, d.d_wpack, d.d_scale, d.d_shift, d.d_out = xin.data_ptr(), wpack.data_ptr(), scale.data_ptr(), shift.data_ptr(), None
    d.B, d.H, d.W, d.Cin, d.Cin_ld, d.N, d.Npad = B, H, W, C, ld_in, N, Npad
    d.ksize, d.leaky, d.epi_mode, d.ldc, d.ch_off = 1, 0, _lib.MC_EPI_DECODE, 0, 0
    d.decode = ctypes.pointer(dec)
    _lib.check(lib.mc_conv_fwd(ctypes.byref(d), s), "mc_conv_fwd(MC_EPI_DECODE)")
    torch.cuda.synchronize()
    return boxes, cls, head


def _dense_to_lists(boxes, cls):
    """Dense slot table -> per image (rows of the candidates in slot order, their class rows)."""
    out = []
    for b in range(boxes.shape[0]):
        sel = boxes[b, :, 7] >= 0
        out.append((boxes[b][sel], None if cls is None else cls[b][sel], torch.nonzero(sel).flatten()))
    return out


def test_fused_decode_epilogue_matches_reference_vectors():
    """The reference's golden decode vectors through the FUSED path (head conv epilogue), plus bit-equality with the
    stand-alone decode kernel on the same logits, plus NMS on the dense table == NMS on the compact one."""
    g, names = _cases()
    for name in names:
        logits = torch.from_numpy(g[name + '_logits']).to(DEV)
        T, oo, val, nt = g[name + '_cfg']
        fb, fc, head = _fused_decode_from_logits(logits, T, int(oo), True)
        assert torch.equal(head, logits), "selection conv did not reproduce the logits exactly"
        cb, cc, ccls = decode_device(logits, T, 20, VOC_ANCHORS, 5, int(oo), want_cls=True)
        dense = _dense_to_lists(fb, fc)
        for b in range(logits.shape[0]):
            rows, probs, slots = dense[b]
            want = g['%s_%d_box' % (name, b)]
            assert rows.shape[0] == want.shape[0] == int(cc[b]), (name, b)
            got = rows.cpu().numpy()
            assert np.array_equal(got[:, 7].astype(np.int32), g['%s_%d_pos' % (name, b)])
            assert np.array_equal(slots.cpu().numpy().astype(np.int32), g['%s_%d_pos' % (name, b)])
            assert np.array_equal(got[:, 6], want[:, 6])
            np.testing.assert_allclose(got[:, :6], want[:, :6], rtol=RTOL, atol=1e-7)
            # same arithmetic as the stand-alone kernel: bit-equal rows and class probabilities
            assert torch.equal(rows, cb[b, :rows.shape[0]])
            assert torch.equal(probs, ccls[b, :rows.shape[0]])
        # NMS straight on the dense table: kept SLOTS == slots of the kept candidates of the compact table
        kd, kcd, rows_d = nms_device(fb, None, float(nt), fc, float(T), want_rows=True)
        kc_, kcc, rows_c = nms_device(cb, cc, float(nt), ccls, float(T), want_rows=True)
        assert torch.equal(kcd, kcc) and torch.equal(rows_d, rows_c)
        for b in range(logits.shape[0]):
            k = int(kcd[b])
            slots = dense[b][2]
            assert torch.equal(kd[b, :k].long(), slots[kc_[b, :k].long()])
            assert kd[b, :k].cpu().tolist() == g['%s_%d_pos' % (name, b)][g['%s_%d_keep' % (name, b)]].tolist()


def test_fused_decode_in_darknet_forward(cfg_path):
    """Through the engine: detect forward (decode in conv23's epilogue) == decode kernel on Darknet.forward's head,
    eager and CUDA-graph replayed, uint8 and float input, both threshold modes."""
    from conftest import make_darknet
    from modelcompression_b200.engine import darknet_detect_forward
    model = make_darknet(cfg_path, seed=0, kn=True, randbn=True, device=DEV)
    torch.manual_seed(11)
    x = torch.randint(0, 256, (3, 3, 416, 416), dtype=torch.uint8, device=DEV)
    with torch.no_grad():
        head = model(x)
        for thr, oo, want_cls in ((0.25, 1, False), (0.005, 0, True)):
            cb, cc, ccls = decode_device(head, thr, 20, model.anchors, model.num_anchors, oo, want_cls)
            for it in range(3):  # 1st eager, 2nd captures the graph, 3rd replays it
                fb, fc = darknet_detect_forward(model, x, thr, oo, want_cls)
                for b, (rows, probs, slots) in enumerate(_dense_to_lists(fb, fc)):
                    n = int(cc[b])
                    assert rows.shape[0] == n > 0
                    assert torch.equal(rows, cb[b, :n])
                    if want_cls:
                        assert torch.equal(probs, ccls[b, :n])
        assert torch.equal(model(x), head)  # the raw-head mode is untouched
