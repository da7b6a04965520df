"""GPU: batch-sharded evaluation equals the unsharded run (same detections, same order) and the per-image reference
flow forward -> get_region_boxes -> nms."""
import numpy as np
import pytest
import torch

import modelcompression_b200 as mc
from conftest import make_darknet
from modelcompression_b200.eval import evaluate_sharded, shard_range
from oracle import detect_oracle

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_sharded_eval_matches_single_pass(cfg_path):
    model = make_darknet(cfg_path, seed=0, kn=True, device=DEV)
    n = 11
    torch.manual_seed(4)
    images = torch.rand(n, 3, 416, 416)

    def get_batch(lo, hi):
        return images[lo:hi].to(DEV)

    full = evaluate_sharded(model, get_batch, n, 4, conf_thresh=0.3, nms_thresh=0.45, only_objectness=1)
    assert full.shape[1] == 8 and full.shape[0] > 0
    assert torch.all(full[1:, 0] >= full[:-1, 0])  # image-major order
    # emulate 3 ranks in one process (the collective itself is covered by the gloo test)
    parts = [evaluate_sharded(model, get_batch, n, 4, 0.3, 0.45, 1, rank=r, world_size=3, gather=False)
             for r in range(3)]
    assert [shard_range(n, r, 3) for r in range(3)] == [(0, 4), (4, 8), (8, 11)]
    assert torch.equal(torch.cat(parts, 0), full)
    # against the oracle's decode+nms on the same head, image by image
    with torch.no_grad():
        head = model(images[:2].to(DEV))
    dec = detect_oracle.decode_np(head.cpu(), 0.3, 20, model.anchors, model.num_anchors, 1)
    for b in range(2):
        keep, _ = detect_oracle.nms_np(dec[b]['box'][:, :5], 0.45)
        mine = full[full[:, 0] == b][:, 1:8].cpu().numpy()
        assert mine.shape[0] == len(keep)
        np.testing.assert_allclose(mine, dec[b]['box'][keep], rtol=2e-5, atol=1e-7)


def test_do_detect_single_image(cfg_path):
    model = make_darknet(cfg_path, seed=0, kn=True, device=DEV)
    torch.manual_seed(9)
    img = (torch.rand(416, 416, 3) * 255).to(torch.uint8).numpy()
    boxes = mc.do_detect(model, img, 0.5, 0.4)
    x = torch.from_numpy(img.transpose(2, 0, 1)).float().div(255.0).unsqueeze(0).to(DEV)
    with torch.no_grad():
        head = model(x)
    dec = detect_oracle.decode_np(head.cpu(), 0.5, 20, model.anchors, model.num_anchors, 1)
    keep, _ = detect_oracle.nms_np(dec[0]['box'][:, :5], 0.4)
    assert len(boxes) == len(keep)
    got = np.array([[float(v) for v in b[:7]] for b in boxes], dtype=np.float32).reshape(-1, 7)
    np.testing.assert_allclose(got, dec[0]['box'][keep], rtol=2e-5, atol=1e-7)
