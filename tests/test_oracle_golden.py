"""CPU: the oracle restatements (oracle/*.py) against the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz, report in tests/golden/PINNING.txt)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import SmallNet, load_golden, make_darknet
from oracle import detect_oracle, forward_oracle, prune_oracle
from modelcompression_b200.cfg import VOC_ANCHORS


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bits(m):
    return np.packbits(np.asarray(m).astype(bool).ravel())


def test_weight_prune_oracle_small_net():
    g = load_golden('weight_prune_small.npz')
    net = SmallNet.build()
    ws = [p.data.numpy() for p in net.parameters() if p.dim() != 1]
    assert len(ws) == 4  # three conv weights + the 2-D fc weight; biases excluded (methods.py:16)
    for i, perc in enumerate(g['percs']):
        thr, masks = prune_oracle.weight_prune_np(ws, float(perc))
        assert np.float32(thr) == g['thr_%d' % i]
        assert np.array_equal(np.concatenate([_bits(m) for m in masks]), g['bits_%d' % i])


@pytest.fixture(scope='module')
def darknet_weights(cfg_path):
    model = make_darknet(cfg_path, seed=0)
    g = load_golden('weight_prune_darknet.npz')
    allp = np.concatenate([p.data.numpy().ravel() for p in model.parameters()])
    assert _sha(allp) == str(g['weights_sha256']), "seed-0 Darknet init differs from the golden run"
    return model


def test_weight_prune_oracle_darknet(darknet_weights):
    g = load_golden('weight_prune_darknet.npz')
    ws = [p.data.numpy() for p in darknet_weights.parameters() if p.dim() != 1]
    for i in (0, 3):  # 70 % and 90 %
        thr, masks = prune_oracle.weight_prune_np(ws, float(g['percs'][i]))
        assert np.float32(thr) == g['thr_%d' % i]
        assert [int(m.size - m.sum()) for m in masks] == g['zeros_%d' % i].tolist()
        assert [_sha(_bits(m)) for m in masks] == g['sha_%d' % i].tolist()
    # README.md:50 of the reference: 15,211,174 params left at 70 % on the pretrained weights (k+1 +- ties); here the
    # structural part of that known answer: pruned count = k+1 + ties with k = 35,444,212
    assert int(g['zeros_0'].sum()) >= 35444213


def test_filter_prune_oracle_darknet(darknet_weights):
    g = load_golden('filter_prune_darknet.npz')
    cw = [p.data.numpy() for p in darknet_weights.parameters() if p.dim() == 4]
    assert [w.shape[0] for w in cw] == g['filters_per_layer'].tolist() and sum(w.shape[0] for w in cw) == 10461
    for i, perc in enumerate(g['percs']):
        values, thr, keep, _ = prune_oracle.quick_filter_prune_np(cw, float(perc), want_masks=False)
        assert np.array_equal(values, g['values'])
        assert thr == g['thr_%d' % i]
        assert np.array_equal(_bits(np.concatenate(keep)), g['keep_%d' % i])
    # explicit summation order (what the CUDA kernel implements) == NumPy, layer by layer
    for w in cw[:6] + cw[-3:]:
        assert np.array_equal(prune_oracle.filter_values_explicit(w), prune_oracle.filter_values_np(w))
    # each layer's largest filter has value 1.0 and is never pruned
    off = 0
    for w in cw:
        assert values[off:off + w.shape[0]].max() == 1.0
        off += w.shape[0]


def test_np_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 64, 125, 128, 129, 255, 256, 512, 1000, 1024, 1280, 4097):
        a = rng.standard_normal(n).astype(np.float32)
        assert prune_oracle.np_pairwise_sum(a) == a.sum(), n


def test_detect_oracle_against_reference_vectors():
    g = load_golden('detect.npz')
    for name in g['names']:
        name = str(name)
        logits = torch.from_numpy(g[name + '_logits'])
        T, oo, val, nt = g[name + '_cfg']
        dec = detect_oracle.decode_np(logits, T, 20, VOC_ANCHORS, 5, int(oo))
        for b in range(logits.shape[0]):
            box = g['%s_%d_box' % (name, b)]
            assert np.array_equal(dec[b]['pos'], g['%s_%d_pos' % (name, b)])
            assert np.array_equal(dec[b]['box'][:, 6], box[:, 6])
            np.testing.assert_allclose(dec[b]['box'], box, rtol=3e-7, atol=0)
            keep, conf_after = detect_oracle.nms_np(box[:, :5], nt)
            assert keep == g['%s_%d_keep' % (name, b)].tolist()
            assert np.array_equal(conf_after, g['%s_%d_conf_after' % (name, b)])


def test_nms_oracle_edge_cases():
    keep, conf = detect_oracle.nms_np(np.zeros((0, 5), np.float32), 0.45)
    assert keep == [] and conf.size == 0
    one = np.array([[0.5, 0.5, 0.1, 0.1, 0.9]], np.float32)
    assert detect_oracle.nms_np(one, 0.45)[0] == [0]
    # identical boxes, equal confidence: the first in list order wins (stable tie rule), the rest are suppressed
    same = np.repeat(one, 4, axis=0)
    keep, conf = detect_oracle.nms_np(same, 0.45)
    assert keep == [0] and conf.tolist() == [np.float32(0.9), 0, 0, 0]
    # zero-confidence boxes are never kept (nets2_utils.py:252)
    z = same.copy()
    z[:, 4] = 0
    assert detect_oracle.nms_np(z, 0.45)[0] == []


def test_forward_oracle_against_reference_vectors(cfg_path):
    import modelcompression_b200 as mc
    g = load_golden('forward.npz')
    torch.manual_seed(int(g['image_seed']))
    img = torch.rand(1, 3, 416, 416)
    model = make_darknet(cfg_path, seed=0, kn=True, randbn=True)
    blocks = mc.parse_cfg(cfg_path)
    with torch.no_grad():
        y, outs = forward_oracle.darknet_forward_fp32(blocks, model.state_dict(), img, keep_outputs=True)
    assert np.array_equal(y.numpy(), g['kn_randbn_head'])
    ids = g['kn_randbn_block_ids'].tolist()
    stats = g['kn_randbn_block_stats']
    for row, i in zip(stats, ids):
        f = outs[i].flatten()
        idx = torch.linspace(0, f.numel() - 1, 32).long()
        np.testing.assert_allclose(f[idx].numpy(), row[3:], rtol=0, atol=0)
    # reorg restatement: out[b,(i*2+j)*C+c,y,x] = in[b,c,2y+i,2x+j]
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).view(2, 3, 4, 6)
    r = forward_oracle.reorg(x)
    for i in range(2):
        for j in range(2):
            assert torch.equal(r[:, (i * 2 + j) * 3:(i * 2 + j + 1) * 3], x[:, :, i::2, j::2])


def test_voc_scorer_matches_reference_golden():
    """N1: the NumPy oracle reproduces the reference's voc_eval APs (tests/golden/voc_map.npz, pinned by
    oracle/make_golden_map.py): 20 classes, ties and 'difficult' boxes included.  (The product's scorer — kernels in
    csrc/voc_eval.cu — is checked against the same vectors on the GPU: tests/test_gpu_eval.py.)"""
    from oracle import map_oracle
    g = load_golden('voc_map.npz')
    # the oracle from the flat rows (one (cls_conf, cls_id) pair per row)
    kept = [[] for _ in range(int(g['n_images']))]
    for r in g['dets']:
        kept[int(r[0])].append([np.float32(v) for v in r[1:7]] + [int(r[7])])
    gt_list = [[] for _ in range(int(g['n_images']))]
    for r in g['gts']:
        gt_list[int(r[0])].append(tuple(int(v) for v in r[1:]))
    rows = map_oracle.detection_rows(kept, [(416, 416)] * len(kept))
    aps_o, m_o = map_oracle.mean_ap(rows, gt_list, 20, 0.5, True)
    assert aps_o == g['ap07'].tolist()


def test_region_loss_matches_reference_golden():
    """N3: the region-loss oracle equals the reference's RegionLoss + build_targets (loss and gradient stored by
    oracle/make_golden_region.py); the GPU test checks the product's kernels against the same vectors and the oracle."""
    import torch
    from oracle.region_oracle import region_loss
    g = load_golden('region_loss.npz')
    anchors = [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071]
    for name in ('ones', 'cfg', 'mixed'):
        cs, ns, os_, cl = g['scales_' + name].tolist()
        out = torch.from_numpy(g['output']).clone().requires_grad_(True)
        loss = region_loss(out, torch.from_numpy(g['target']), anchors, 5, 20, cs, ns, os_, cl, 0.6)
        loss.backward()
        assert abs(float(loss.detach()) - float(g['loss_' + name])) <= 2e-6 * abs(float(g['loss_' + name]))
        gref = torch.from_numpy(g['grad_' + name])
        assert float((out.grad - gref).abs().max()) <= 2e-6 * float(gref.abs().max())


def test_weights_file_written_by_the_reference(tmp_path):
    """a darknet .weights file written by the UNMODIFIED reference's save_weights
    (oracle/make_golden_weights.py; src/nets.py:1007-1051) loads here tensor for tensor, this package's save_weights
    reproduces it byte for byte, a darknet v0.2 copy (int64 `seen`) loads to the same tensors, and the shrunk-model
    writer round-trips through its side-car."""
    import torch
    import modelcompression_b200 as mc
    g = load_golden('weights_mini.npz')
    cfg = tmp_path / 'mini.cfg'
    cfg.write_bytes(g['cfg'].tobytes())
    raw = g['file_bytes'].tobytes()
    wfile = tmp_path / 'ref.weights'
    wfile.write_bytes(raw)
    model = mc.Darknet(str(cfg))
    model.load_weights(str(wfile))
    sd = model.state_dict()
    keys = [k[3:] for k in g.files if k.startswith('sd.')]
    assert sorted(keys) == sorted(sd.keys())
    for k in keys:
        if not k.endswith('num_batches_tracked'):
            assert np.array_equal(sd[k].numpy(), g['sd.' + k]), k
    assert model.seen == int(g['seen'])
    out = tmp_path / 'mine.weights'
    model.save_weights(str(out))
    assert out.read_bytes() == raw
    # v0.2 header variant of the same payload
    v2 = tmp_path / 'v2.weights'
    v2.write_bytes(np.array([0, 2, 0], np.int32).tobytes() + np.array([int(g['seen'])], np.int64).tobytes() + raw[16:])
    m2 = mc.Darknet(str(cfg))
    m2.load_weights(str(v2))
    for k in keys:
        if not k.endswith('num_batches_tracked'):
            assert np.array_equal(m2.state_dict()[k].numpy(), g['sd.' + k]), k
    # shrunk writer: zero three filters of the second conv, write, restore
    conv = model.masked_convs()[1]
    mask = torch.ones_like(conv.weight.data)
    mask[[1, 5, 9]] = 0
    conv.register_buffer('mask', mask)
    conv.weight.data *= mask
    conv.mask_flag = True
    side = model.save_shrunk_weights(str(tmp_path / 's.weights'), str(tmp_path / 's.json'))
    assert len(side['layers'][1]['keep_out']) == 13 and len(side['layers'][2]['keep_in']) == 13
    assert 'filters=13' in side['cfg']
    m3 = mc.Darknet(str(cfg))
    masks = m3.load_shrunk_weights(str(tmp_path / 's.weights'), str(tmp_path / 's.json'))
    assert torch.equal(masks[1], mask)
    alive = torch.ones(16)
    alive[[1, 5, 9]] = 0
    sd3 = m3.state_dict()
    for ka, a in model.state_dict().items():
        if ka.endswith('num_batches_tracked') or ka.endswith('.mask'):
            continue
        if ka.endswith('conv3.weight'):
            a = a * alive.view(1, -1, 1, 1)  # weights on removed input channels are dropped (side-car `fold`)
        assert torch.equal(a, sd3[ka]), ka
    assert side['layers'][2]['fold'] is not None  # rand BN statistics: the removed filters' constants are non-zero
