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
