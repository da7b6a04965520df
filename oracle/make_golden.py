"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded synthetic
inputs, and pin the oracle restatements (oracle/*.py) against it.  Run in the build container only:

    python oracle/make_golden.py            # ~5 min on 8 cores

/root/reference does not exist on the GPU box; tests read only the committed fixtures.
"""
import hashlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import detect_oracle, forward_oracle, prune_oracle, ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
REPORT = []


def log(msg):
    print(msg, flush=True)
    REPORT.append(msg)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def packbits(mask):
    return np.packbits(np.asarray(mask).astype(bool).ravel())


class SmallNet(torch.nn.Module):
    """4-D, 2-D and 1-D parameters: exercises the dim != 1 / dim == 4 selection rules of the pruners."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 8, 3)
        self.c2 = torch.nn.Conv2d(8, 16, 3)
        self.c3 = torch.nn.Conv2d(16, 12, 1)
        self.fc = torch.nn.Linear(12, 10)


def small_net(seed=11):
    torch.manual_seed(seed)
    return SmallNet()


def golden_small_weight_prune(ref):
    model = small_net()
    percs = [0., 10., 33.3, 50., 70., 90., 99.5, 100.]
    out = {'percs': np.array(percs)}
    ws = [p.data.numpy() for p in model.parameters() if p.dim() != 1]
    for i, perc in enumerate(percs):
        masks = ref_shim.quiet(ref['methods'].weight_prune, model, perc)
        thr_o, masks_o = prune_oracle.weight_prune_np(ws, perc)
        for a, b in zip(masks, masks_o):
            assert np.array_equal(a.numpy(), b), 'oracle weight_prune_np != reference (small net, %s)' % perc
        out['thr_%d' % i] = np.float32(thr_o)
        out['bits_%d' % i] = np.concatenate([packbits(m.numpy()) for m in masks])
    log('weight_prune small net: oracle == reference at %s' % percs)
    np.savez_compressed(os.path.join(GOLD, 'weight_prune_small.npz'), **out)


def golden_darknet_prune(ref):
    torch.manual_seed(0)
    model = ref_shim.quiet(ref['nets'].Darknet, ref['cfg'])
    # same-seed model from the product package must be identical (same module construction order)
    import modelcompression_b200 as mc
    torch.manual_seed(0)
    mine = mc.Darknet(mc.write_yolov2_voc_cfg())
    sd_r, sd_m = model.state_dict(), mine.state_dict()
    assert list(sd_r.keys()) == list(sd_m.keys()), 'state_dict keys differ'
    for k in sd_r:
        assert torch.equal(sd_r[k], sd_m[k]), k
    log('Darknet(seed 0): state_dict keys and values identical between reference and modelcompression_b200 '
        '(%d tensors, %d params)' % (len(sd_r), sum(p.numel() for p in model.parameters())))
    wsha = sha(np.concatenate([p.data.numpy().ravel() for p in model.parameters()]))

    ws = [p.data.numpy() for p in model.parameters() if p.dim() != 1]
    out = {'weights_sha256': np.array(wsha), 'percs': np.array([70., 75., 80., 90.])}
    for i, perc in enumerate([70., 75., 80., 90.]):
        t0 = time.time()
        masks = ref_shim.quiet(ref['methods'].weight_prune, model, perc)
        t_ref = time.time() - t0
        thr_o, masks_o = prune_oracle.weight_prune_np(ws, perc)
        for a, b in zip(masks, masks_o):
            assert np.array_equal(a.numpy(), b), 'oracle weight_prune_np != reference (darknet, %s)' % perc
        out['thr_%d' % i] = np.float32(thr_o)
        out['zeros_%d' % i] = np.array([int(m.numel() - m.sum().item()) for m in masks], dtype=np.int64)
        out['sha_%d' % i] = np.array([sha(packbits(m.numpy())) for m in masks])
        log('weight_prune darknet %.0f%%: oracle == reference; thr=%r pruned=%d (reference took %.1f s)' %
            (perc, float(thr_o), int(out['zeros_%d' % i].sum()), t_ref))
    np.savez_compressed(os.path.join(GOLD, 'weight_prune_darknet.npz'), **out)

    # ---- filter pruner
    cw = [p.data.numpy() for p in model.parameters() if p.dim() == 4]
    for w in cw:
        assert np.array_equal(prune_oracle.filter_values_np(w), prune_oracle.filter_values_explicit(w)), \
            'explicit summation order != NumPy for shape %s' % (w.shape,)
    log('filter values: explicit float32 summation order == NumPy on all %d conv layers' % len(cw))
    percs = [5., 20., 33.3, 40., 60., 80., 97.25]
    fout = {'percs': np.array(percs)}
    for i, perc in enumerate(percs):
        t0 = time.time()
        masks = ref_shim.quiet(ref['methods'].quick_filter_prune, model, perc)
        t_ref = time.time() - t0
        values, thr, keep, masks_o = prune_oracle.quick_filter_prune_np(cw, perc)
        for a, b in zip(masks, masks_o):
            assert np.array_equal(a.numpy(), b), 'oracle quick_filter_prune_np != reference (%s)' % perc
        fout['thr_%d' % i] = np.float64(thr)
        fout['keep_%d' % i] = packbits(np.concatenate(keep))
        log('quick_filter_prune %.2f%%: oracle == reference; thr=%r kept=%d/%d (reference took %.2f s)' %
            (perc, float(thr), int(sum(k.sum() for k in keep)), values.size, t_ref))
    fout['values'] = values
    fout['filters_per_layer'] = np.array([w.shape[0] for w in cw])
    np.savez_compressed(os.path.join(GOLD, 'filter_prune_darknet.npz'), **fout)

    # ---- set_masks / prune_rate / are_masks_consistent (70% weight pruning)
    masks = ref_shim.quiet(ref['methods'].weight_prune, model, 90.)
    model.set_masks(masks)
    rate = ref_shim.quiet(ref['putils'].prune_rate, model, False)
    cons = ref['putils'].are_masks_consistent(model, masks)
    rate_o = prune_oracle.prune_rate_np([p.data.numpy() for p in model.parameters()])
    assert rate == rate_o
    log('prune_rate after 90%% weight pruning: %.6f (oracle equal); are_masks_consistent=%s; mask buffers in '
        'state_dict: %d' % (rate, cons, sum(1 for k in model.state_dict() if k.endswith('.mask'))))
    np.savez_compressed(os.path.join(GOLD, 'prune_rate.npz'), rate90=np.float64(rate), consistent=np.array(cons))
    return model


def golden_detect(ref):
    n2 = ref['nets2_utils']
    from modelcompression_b200.cfg import VOC_ANCHORS
    g = torch.Generator().manual_seed(2)
    cases = []
    # (name, logits, conf_thresh, only_objectness, validation, nms_thresh)
    lg = torch.randn(2, 125, 13, 13, generator=g) * 2.0
    cases.append(('n2_obj', lg, 0.25, 1, False, 0.45))
    lg = torch.randn(2, 125, 13, 13, generator=g) * 2.0
    cases.append(('n2_val', lg, 0.2, 0, True, 0.45))
    lg = torch.randn(1, 125, 13, 13, generator=g) * 0.02  # init-like logits: every box passes, conf ~ 0.5
    cases.append(('init_like', lg, 0.005, 1, False, 0.4))
    lg = torch.randn(1, 125, 13, 13, generator=g) * 2.0
    for a in range(5):  # exact ties in det_conf: groups of equal objectness logits
        lg[0, a * 25 + 4] = torch.round(lg[0, a * 25 + 4])
    cases.append(('ties', lg, 0.3, 1, False, 0.45))
    out = {'names': np.array([c[0] for c in cases])}
    for name, logits, T, oo, val, nt in cases:
        t0 = time.time()
        allb = n2.get_region_boxes(logits, T, 20, VOC_ANCHORS, 5, oo, val)
        dec_o = detect_oracle.decode_np(logits, T, 20, VOC_ANCHORS, 5, oo)
        out[name + '_logits'] = logits.numpy()
        out[name + '_cfg'] = np.array([T, oo, 1 if val else 0, nt], dtype=np.float64)
        for b, boxes in enumerate(allb):
            box7 = np.array([[float(x) for x in bx[:7]] for bx in boxes], dtype=np.float32).reshape(-1, 7)
            o = dec_o[b]
            assert box7.shape[0] == o['box'].shape[0], (name, b, box7.shape, o['box'].shape)
            # restatement vs reference: ids and geometry exact; sigmoid may differ by 1 ulp (layout-dependent)
            assert np.array_equal(box7[:, 6], o['box'][:, 6])
            np.testing.assert_allclose(box7, o['box'], rtol=3e-7, atol=0)
            extras = []
            if val and not oo:
                for r, bx in enumerate(boxes):
                    ex = bx[7:]
                    for j in range(0, len(ex), 2):
                        extras.append((r, int(ex[j + 1]), float(ex[j])))
            with ref_shim.stable_sort():
                kept = n2.nms(boxes, nt)
            ids = {id(bx): i for i, bx in enumerate(boxes)}
            kept_idx = np.array([ids[id(bx)] for bx in kept], dtype=np.int32)
            conf_after = np.array([float(bx[4]) for bx in boxes], dtype=np.float32)
            keep_o, conf_o = detect_oracle.nms_np(box7[:, :5], nt)
            assert list(kept_idx) == keep_o, 'oracle nms_np kept list != reference (%s img %d)' % (name, b)
            assert np.array_equal(conf_after, conf_o), 'oracle nms_np conf mutation != reference'
            out['%s_%d_pos' % (name, b)] = o['pos']
            out['%s_%d_box' % (name, b)] = box7
            out['%s_%d_keep' % (name, b)] = kept_idx
            out['%s_%d_conf_after' % (name, b)] = conf_after
            out['%s_%d_extras' % (name, b)] = np.array(extras, dtype=np.float64).reshape(-1, 3)
            nties = int(box7.shape[0] - np.unique((np.float32(1) - box7[:, 4]).astype(np.float32)).size)
            log('detect %s img %d: %d candidates (%d tied keys) -> %d kept; oracle decode/nms == reference '
                '(%.1f s)' % (name, b, box7.shape[0], nties, kept_idx.size, time.time() - t0))
    np.savez_compressed(os.path.join(GOLD, 'detect.npz'), **out)


def block_stats(outputs):
    """Per-block fingerprints small enough to commit: mean, std, absmax and 32 strided samples."""
    rows = {}
    for ind, t in outputs.items():
        f = t.detach().flatten()
        idx = torch.linspace(0, f.numel() - 1, 32).long()
        rows[ind] = np.concatenate([[f.mean().item(), f.std().item(), f.abs().max().item()], f[idx].numpy()])
    return rows


def golden_forward(ref):
    nets = ref['nets']
    out = {}
    torch.manual_seed(1)
    img = torch.rand(1, 3, 416, 416)

    def run(tag, model):
        model.eval()
        blocks = model.blocks
        outputs = {}
        # the reference forward keeps its per-block outputs in a local dict: re-walk with hooks on self.models
        hooks = []
        for i, m in enumerate(model.models):
            hooks.append(m.register_forward_hook(lambda mod, inp, o, i=i: outputs.__setitem__(i, o)))
        with torch.no_grad():
            t0 = time.time()
            y = model(img)
            dt = time.time() - t0
        for h in hooks:
            h.remove()
        y_o, outs_o = forward_oracle.darknet_forward_fp32(blocks, model.state_dict(), img, keep_outputs=True)
        assert torch.equal(y, y_o), 'forward oracle != reference (%s)' % tag
        for i, o in outputs.items():
            if torch.is_tensor(o):
                assert torch.equal(o, outs_o[i]), (tag, i)
        out[tag + '_head'] = y.numpy()
        st = block_stats(outs_o)
        out[tag + '_block_ids'] = np.array(sorted(st))
        out[tag + '_block_stats'] = np.stack([st[i] for i in sorted(st)])
        log('forward %s: oracle == reference bit-exactly on head and all %d hooked blocks; head std %.4g '
            '(reference CPU forward b1: %.2f s)' % (tag, len(outputs), y.std().item(), dt))

    torch.manual_seed(0)
    m = ref_shim.quiet(nets.Darknet, ref['cfg'])
    run('default', m)
    forward_oracle.kaiming_normal_init_(m, 7)
    run('kn', m)
    forward_oracle.randomize_bn_(m, 1)
    run('kn_randbn', m)
    masks = ref_shim.quiet(ref['methods'].weight_prune, m, 70.)
    m.set_masks(masks)
    run('kn_randbn_w70', m)
    # filter pruning on a fresh KN + rand-BN model
    torch.manual_seed(0)
    m = ref_shim.quiet(nets.Darknet, ref['cfg'])
    forward_oracle.kaiming_normal_init_(m, 7)
    forward_oracle.randomize_bn_(m, 1)
    masks = ref_shim.quiet(ref['methods'].quick_filter_prune, m, 40.)
    m.set_masks(masks)
    run('kn_randbn_f40', m)
    out['image_seed'] = np.array(1)
    np.savez_compressed(os.path.join(GOLD, 'forward.npz'), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load_reference()
    log('reference imported from %s; torch %s numpy %s' % (ref_shim.REFERENCE_ROOT, torch.__version__, np.__version__))
    golden_small_weight_prune(ref)
    golden_detect(ref)
    golden_forward(ref)
    golden_darknet_prune(ref)
    with open(os.path.join(GOLD, 'PINNING.txt'), 'w') as f:
        f.write('\n'.join(REPORT) + '\n')


if __name__ == '__main__':
    main()
