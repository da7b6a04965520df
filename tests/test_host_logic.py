"""CPU: host-side logic that mirrors the reference interface (cfg parsing, module tree, percentile rank arithmetic,
weights IO, sharding helpers)."""
import os

import numpy as np
import pytest
import torch

import modelcompression_b200 as mc
from modelcompression_b200.eval import shard_range
from modelcompression_b200.pruning.weightPruning.methods import percentile_rank
from modelcompression_b200.pruning.weightPruning.utils import arg_nonzero_min


def _lerp(a, b, t):
    d = b - a
    r = a + d * t
    if t >= 0.5:
        r = b - d * (1 - t)
    return r


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_percentile_rank_matches_numpy(dtype):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 17, 1000, 10461, 65537):
        a = np.abs(rng.standard_normal(n)).astype(dtype)
        s = np.sort(a)
        for perc in (0., 0.01, 5., 20., 33.3, 40., 50., 60., 70., 75., 80., 90., 99.99, 100.):
            k, g = percentile_rank(n, perc, dtype)
            got = _lerp(s[k], s[min(k + 1, n - 1)], dtype(g))
            assert got == np.percentile(a, perc), (n, perc)


def test_percentile_rank_darknet_float32_virtual_index():
    # n = 50,634,592 float32 magnitudes -> the virtual index is evaluated in float32
    n = 50634592
    assert percentile_rank(n, 70., np.float32) == (35444212, 0.0)
    assert percentile_rank(n, 75., np.float32) == (37975944, 0.0)
    assert percentile_rank(n, 80., np.float32) == (40507676, 0.0)
    assert percentile_rank(n, 90., np.float32) == (45571132, 0.0)
    # 10,461 filters, float64: integer ranks at 20/40/60/80 %
    assert [percentile_rank(10461, p, np.float64)[0] for p in (20., 40., 60., 80.)] == [2092, 4184, 6276, 8368]
    with pytest.raises(ValueError):
        percentile_rank(10, 101., np.float32)


def test_parse_cfg_and_module_tree(cfg_path):
    blocks = mc.parse_cfg(cfg_path)
    assert blocks[0]['type'] == 'net' and blocks[-1]['type'] == 'region'
    assert sum(b['type'] == 'convolutional' for b in blocks) == 23
    assert sum(b['type'] == 'maxpool' for b in blocks) == 5
    assert blocks[-2]['batch_normalize'] == 0  # default for [convolutional] without the key
    model = mc.Darknet(cfg_path)
    assert len(model.models) == 32
    assert sum(p.numel() for p in model.parameters()) == 50655389  # README.md:34 of the reference
    assert sum(p.numel() for p in model.parameters() if p.dim() != 1) == 50634592
    assert (model.width, model.height, model.num_anchors, model.num_classes) == (416, 416, 5, 20)
    assert model.anchor_step == 2.0 and len(model.anchors) == 10
    keys = list(model.state_dict().keys())
    assert keys[0] == 'models.0.conv1.weight' and 'models.30.conv23.bias' in keys
    assert 'models.29.bn22.running_var' in keys and 'models.0.bn1.num_batches_tracked' in keys
    convs = model.masked_convs()
    assert len(convs) == 23 and all(c.name == 'MaskedConv2d' and c.mask_flag is False for c in convs)
    # 1x1 convs get padding 0 although the cfg says pad=1 (nets.py:796)
    assert model.models[5][0].padding == (0, 0) and model.models[4][0].padding == (1, 1)
    with pytest.raises(ValueError):
        model.set_masks([])


def test_parse_cfg_type_key(tmp_path):
    p = tmp_path / 'x.cfg'
    p.write_text("[net]\nwidth=32\nheight=32\nchannels=3\n# comment\n\n[cost]\ntype=sse\n")
    blocks = mc.parse_cfg(str(p))
    assert blocks[1] == {'type': 'cost', '_type': 'sse'}


def test_weights_file_roundtrip(cfg_path, tmp_path):
    torch.manual_seed(5)
    a = mc.Darknet(cfg_path)
    for m in a.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 1.5)
    a.seen = 12345
    path = str(tmp_path / 'w.weights')
    a.save_weights(path)
    # header = 4 x int32 (0,0,0,seen) + 50,655,389 + 20,672 running stats floats
    n_float = sum(p.numel() for p in a.parameters()) + sum(m.num_features * 2 for m in a.modules()
                                                         if isinstance(m, torch.nn.BatchNorm2d))
    import os
    assert os.path.getsize(path) == 16 + 4 * n_float
    b = mc.Darknet(cfg_path)
    b.load_weights(path)
    assert b.seen == 12345
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if not k.endswith('num_batches_tracked'):
            assert torch.equal(sa[k], sb[k]), k


def test_weights_v02_header_load_then_save_is_readable(cfg_path, tmp_path):
    """A darknet v0.2 file (int64 `seen`, 20-byte header — what the official yolov2-voc.weights is) must survive
    load -> save -> load.  The reference discards the file's version on load (src/nets.py:899-905) and always writes
    0.0.0 with an int32 `seen`; copying major/minor into self.header made save_weights write an unreadable file."""
    torch.manual_seed(6)
    a = mc.Darknet(cfg_path)
    p0 = str(tmp_path / 'v0.weights')
    a.save_weights(p0)
    raw = open(p0, 'rb').read()
    p2 = str(tmp_path / 'v2.weights')
    with open(p2, 'wb') as f:
        np.array([0, 2, 0], dtype=np.int32).tofile(f)
        np.array([32013312], dtype=np.int64).tofile(f)
        f.write(raw[16:])
    b = mc.Darknet(cfg_path)
    b.load_weights(p2)
    assert b.header.tolist()[:3] == [0, 0, 0]  # the reference never touches self.header on load
    assert b.seen == 32013312
    p3 = str(tmp_path / 'resaved.weights')
    b.save_weights(p3)
    assert os.path.getsize(p3) == len(raw)  # 16-byte header again
    c = mc.Darknet(cfg_path)
    c.load_weights(p3)
    assert c.seen == 32013312
    sa, sc = a.state_dict(), c.state_dict()
    for k in sa:
        if not k.endswith('num_batches_tracked'):
            assert torch.equal(sa[k], sc[k]), k


def test_arg_nonzero_min_reference_quirks():
    assert arg_nonzero_min([0.0, 3.0, 2.0, 5.0]) == (2.0, 2)
    assert arg_nonzero_min([]) is None
    assert arg_nonzero_min([4.0, 0.0]) == (np.inf, np.inf)  # only nonzero at index 0 -> "all zero" (reference quirk)


def test_bbox_iou_host():
    a = [0.5, 0.5, 0.2, 0.2]
    assert mc.bbox_iou(a, a, x1y1x2y2=False) == pytest.approx(1.0)
    assert mc.bbox_iou(a, [0.9, 0.9, 0.1, 0.1], x1y1x2y2=False) == 0.0
    assert mc.bbox_iou([0, 0, 2, 2], [1, 1, 3, 3]) == pytest.approx(1.0 / 7.0)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 4952):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
    assert shard_range(4952, 7, 8) == (4333, 4952) and shard_range(4952, 0, 8) == (0, 619)


def test_param_index_tracks_module_tree(cfg_path):
    """The pruners' cached parameter index must equal model.parameters() after any edit of the module tree."""
    import torch
    from modelcompression_b200.pruning.weightPruning import methods
    model = mc.Darknet(cfg_path)

    def same():
        return [id(p) for p in methods._all_parameters(model)] == [id(p) for p in model.parameters()]
    assert same() and same()
    model.models[0][0].weight = torch.nn.Parameter(model.models[0][0].weight.data.clone())  # replaced Parameter
    assert same()
    model.models[2] = torch.nn.Sequential(torch.nn.Conv2d(32, 64, 3))                       # replaced sub-module
    assert same()
    model.extra = torch.nn.Linear(3, 3)                                                      # added sub-module
    assert same()
    del model.extra                                                                          # removed sub-module
    assert same()
    model.models[4][0].weight = model.models[5][0].weight                                    # shared Parameter
    assert same()


def test_train_plan_graph(cfg_path):
    """Training plan of yolov2-voc.cfg (host logic only): 23 convolutions, route -9 reads the UNPOOLED block-16
    activation, conv21's Reorg output and conv20 share the concat buffer (reorg part first), head is linear."""
    from modelcompression_b200.engine_train import TrainPlan
    from modelcompression_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("libmcb200.so not built")
    model = mc.Darknet(cfg_path)
    plan = TrainPlan(model)
    Ls = {L.ind: L for L in plan.layers}
    assert len(plan.layers) == 23 and plan.layers[-1].is_head and plan.layers[-1].O == 125
    assert Ls[26].src.name == 'a16' and (Ls[26].src.H, Ls[26].src.W) == (26, 26) and Ls[26].reorg
    assert Ls[18].src.name == 'p16' and Ls[16].pool
    assert Ls[26].act.name == Ls[24].act.name and Ls[26].act.ch_off == 0 and Ls[24].act.ch_off == 256
    assert Ls[29].C == 1280 and Ls[29].src.name == Ls[24].act.name
    ps = plan.parameters()
    assert sum(p.numel() for p in ps) == 50655389 and len({id(p) for p in ps}) == len(list(model.parameters()))


def test_activation_pitch_policy():
    """engine._pitch: 8-channel granularity up to 16 channels (the im2col kernel's pitch), 32 for 17..32, otherwise the
    next multiple of 64 when that grows the row by <= 25 %, else the next multiple of 8; always >= n and 16-byte rows."""
    from modelcompression_b200.engine import _pitch
    assert [_pitch(n) for n in (1, 4, 8, 9, 16)] == [8, 8, 8, 16, 16]
    assert [_pitch(n) for n in (17, 22, 32)] == [32, 32, 32]
    assert _pitch(33) == 40 and _pitch(69) == 72 and _pitch(91) == 96 and _pitch(145) == 152   # > 25 % growth: stay
    assert _pitch(56) == 64 and _pitch(158) == 192 and _pitch(811) == 832 and _pitch(1006) == 1024 and _pitch(1018) == 1024
    for n in range(1, 1300):
        ld = _pitch(n)
        assert ld >= n and ld % 8 == 0 and (n <= 32 or ld <= max((n + 7) // 8 * 8, int(1.25 * ((n + 7) // 8 * 8))))


def test_bench_numa_binding_is_optional():
    """bench.bind_to_gpu_numa must not raise where NVML / the GPU is missing (it returns None and the run proceeds)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(os.path.dirname(os.path.dirname(__file__)), 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    got = bench.bind_to_gpu_numa(0)
    assert got is None or (isinstance(got, int) and got >= 1)
    if got is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_thin_kernel_geometry_and_weight_layout():
    """csrc/conv_thin.cu takes the degenerate layers of a filter-pruned net; which shapes it takes (a host-side rule of
    the library: no GPU involved) and the kernel-parameter weight layout the engine builds for it."""
    import ctypes
    from modelcompression_b200 import _lib
    from modelcompression_b200.engine import _thin_host_weights
    lib = _lib.load()

    def geom(k, cin, n, pool, n2):
        ct, nt, n2t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        ok = lib.mc_conv_thin_geometry(k, cin, n, pool, n2, ctypes.byref(ct), ctypes.byref(nt), ctypes.byref(n2t))
        return (ct.value, nt.value, n2t.value) if ok == 1 else None

    # the bench network's stem after 40 % filter pruning
    assert geom(3, 4, 1, 1, 0) == (4, 1, 0)      # conv2 4 -> 1 + pool
    assert geom(3, 1, 17, 0, 0) == (2, 20, 0)    # conv3 1 -> 17
    assert geom(3, 1, 17, 0, 4) == (2, 20, 4)    # conv3 with conv4 behind it
    assert geom(1, 17, 4, 0, 0) == (24, 4, 0)    # conv4 17 -> 4 (1x1)
    assert geom(3, 4, 11, 1, 0) == (4, 12, 0)    # conv5 4 -> 11 + pool
    # real contractions stay on the tensor cores
    assert geom(3, 11, 78, 0, 0) is None         # conv6
    assert geom(3, 8, 16, 0, 0) is None          # 9 * 8 * 16 multiply-adds per pixel: above the budget
    assert geom(1, 78, 16, 0, 0) is None         # 1x1 with more than 32 inputs
    assert geom(3, 5, 12, 1, 0) is None          # pooled: four positions per thread, budget 4x tighter
    assert geom(1, 17, 4, 1, 0) is None          # no pooled 1x1
    assert geom(3, 4, 30, 0, 0) is None          # more than 24 outputs

    # weight layout [(tap*ct + c)*nt + n], bf16-rounded, zero padded; scale / shift padded to nt
    torch.manual_seed(0)
    wd = torch.randn(3, 2, 3, 3)
    sc, sh = torch.rand(3) + 0.5, torch.randn(3)
    hw, hsc, hsh = _thin_host_weights(wd, sc, sh, 4, 8)
    w = np.array(list(hw), dtype=np.float32).reshape(9, 4, 8)
    want = wd.to(torch.bfloat16).float().numpy()
    for n in range(3):
        for c in range(2):
            assert np.array_equal(w[:, c, n], want[n, c].reshape(9))
    assert not w[:, 2:, :].any() and not w[:, :, 3:].any()
    assert np.allclose(np.array(list(hsc))[:3], sc.numpy()) and not np.array(list(hsc))[3:].any()
    assert np.allclose(np.array(list(hsh))[:3], sh.numpy()) and not np.array(list(hsh))[3:].any()
