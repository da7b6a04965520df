"""Pins oracle/train_oracle.py against the UNMODIFIED reference (run in the build container, CPU only):
the reference Darknet (src/nets.py) in train() mode with 90 % weight_prune masks set, one forward + backward of
loss = (y*g).sum() on seeded inputs; asserts the oracle's logits, every gradient and the running statistics equal the
reference's, and writes tests/golden/train_step.npz (checksums of the reference results + the tensors small enough to
ship) for the GPU-box tests, where /root/reference does not exist."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import forward_oracle, ref_shim, train_oracle  # noqa: E402


def main():
    ref = ref_shim.load_reference()
    torch.manual_seed(0)
    model = ref_shim.quiet(ref['nets'].Darknet, ref['cfg'])
    forward_oracle.kaiming_normal_init_(model, 7)
    forward_oracle.randomize_bn_(model, 1)
    masks = ref_shim.quiet(ref['methods'].weight_prune, model, 90.)
    model.set_masks(masks)
    model.train()
    B = 2
    torch.manual_seed(1)
    x = torch.rand(B, 3, 416, 416)
    torch.manual_seed(3)
    g = torch.randn(B, 125, 13, 13)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    y = model(x)
    (y * g).sum().backward()
    y_o, grads_o, buf_o = train_oracle.train_step_fp32(model.blocks, state0, x, g)
    assert torch.equal(y_o, y.detach()), "oracle logits differ from the reference"
    out = {'B': np.int64(B), 'y_sha': hashlib.sha256(y.detach().numpy().tobytes()).hexdigest()}
    names = []
    for name, p in model.named_parameters():
        assert torch.equal(grads_o[name], p.grad), "oracle gradient differs from the reference: " + name
        names.append(name)
        out['gnorm.' + name] = np.float64(p.grad.double().norm().item())
        if p.grad.numel() <= 4096:
            out['grad.' + name] = p.grad.numpy()
    sd = model.state_dict()
    for k, v in buf_o.items():
        if 'running' in k:
            assert torch.equal(v, sd[k]), "oracle running statistic differs from the reference: " + k
    # masked weights get exactly zero gradient
    for m, p in zip(masks, [p for p in model.parameters() if p.dim() == 4]):
        assert float((p.grad * (1 - m)).abs().max()) == 0.0
    out['names'] = np.array(names)
    out['y'] = y.detach().numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'train_step.npz'), **out)
    with open(os.path.join(ROOT, 'tests', 'golden', 'PINNING.txt'), 'a') as f:
        f.write("train step (KN init, rand-BN, 90%% weight masks, B=%d): oracle == reference on logits, all %d gradients and "
                "the running statistics (oracle/make_golden_train.py)\n" % (B, len(names)))
    print("pinned: logits, %d gradients, running statistics equal the reference" % len(names))


if __name__ == '__main__':
    main()
