"""Pins the vectorised region loss (oracle/region_oracle.py) against the UNMODIFIED reference
RegionLoss + build_targets (src/nets.py:282-636), CPU only.  The reference allocates torch.cuda tensors
unconditionally; the shim maps torch.cuda.FloatTensor/LongTensor to the CPU types (Tensor.cuda is already the
identity, oracle/ref_shim.py).  Asserts loss and d loss / d output agree to float32 round-off for several scale settings,
including two boxes on the same (anchor, cell), a box with zero width and an early list terminator, and writes
tests/golden/region_loss.npz for the GPU-box test."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.region_oracle import region_loss  # noqa: E402

ANCHORS = [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071]


def make_targets(nB, seed=4):
    rng = np.random.RandomState(seed)
    t = np.zeros((nB, 250), dtype=np.float32)
    for b in range(nB):
        n = rng.randint(1, 6)
        for k in range(n):
            t[b, k * 5:(k + 1) * 5] = [rng.randint(0, 20), rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9),
                                       rng.uniform(0.05, 0.5), rng.uniform(0.05, 0.5)]
    # two boxes in the same cell with the same best anchor (the later one must win)
    t[0, 0:5] = [3, 0.52, 0.48, 0.30, 0.35]
    t[0, 5:10] = [7, 0.53, 0.47, 0.31, 0.34]
    # a degenerate box (zero width): no anchor has positive IoU -> the reference indexes anchor -1
    t[1, 5:10] = [5, 0.30, 0.70, 0.0, 0.2]
    # early terminator: boxes after an x == 0 entry are ignored
    t[2, 5:10] = [2, 0.0, 0.5, 0.2, 0.2]
    t[2, 10:15] = [9, 0.6, 0.6, 0.2, 0.2]
    return torch.from_numpy(t)


def main():
    ref = ref_shim.load_reference()
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = torch.LongTensor
    nets = ref['nets']
    nB = 4
    torch.manual_seed(6)
    out0 = torch.randn(nB, 125, 13, 13) * 0.5
    target = make_targets(nB)
    gold = {'output': out0.numpy(), 'target': target.numpy()}
    for name, (cs, ns, os_, cl) in {'ones': (1, 1, 1, 1), 'cfg': (1, 1, 5, 1), 'mixed': (2.0, 0.5, 5, 1.5)}.items():
        crit = nets.RegionLoss(20, ANCHORS, 5)
        crit.coord_scale, crit.noobject_scale, crit.object_scale, crit.class_scale = cs, ns, os_, cl
        o_ref = out0.clone().requires_grad_(True)
        loss_ref = ref_shim.quiet(crit, o_ref, target)
        loss_ref.backward()
        o_new = out0.clone().requires_grad_(True)
        loss_new = region_loss(o_new, target, ANCHORS, 5, 20, cs, ns, os_, cl, 0.6)
        loss_new.backward()
        rel = abs(float(loss_new) - float(loss_ref)) / abs(float(loss_ref))
        gerr = float((o_new.grad - o_ref.grad).abs().max()) / float(o_ref.grad.abs().max())
        assert rel < 2e-6 and gerr < 2e-6, (name, float(loss_ref), float(loss_new), gerr)
        gold['loss_' + name] = np.float64(float(loss_ref))
        gold['grad_' + name] = o_ref.grad.numpy()
        gold['scales_' + name] = np.array([cs, ns, os_, cl], dtype=np.float64)
        print("%s: reference loss %.6f, vectorised %.6f (rel %.1e), grad max rel err %.1e" %
              (name, float(loss_ref), float(loss_new), rel, gerr))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'region_loss.npz'), **gold)
    with open(os.path.join(ROOT, 'tests', 'golden', 'PINNING.txt'), 'a') as f:
        f.write("region loss (B=4, overwrite / zero-width / terminator cases, 3 scale settings): vectorised == reference "
                "RegionLoss to 2e-6 on loss and gradient (oracle/make_golden_region.py)\n")


if __name__ == '__main__':
    main()
