"""Golden darknet ``.weights`` file written by the UNMODIFIED reference (src/nets.py:1007-1051 save_weights, save_conv_bn
:236-248, save_conv :202-208) for a small cfg with every block type of yolov2-voc.cfg (TEST INFRASTRUCTURE).

Writes tests/golden/weights_mini.npz: the cfg text, the reference's file bytes, and the reference model's state_dict.
tests/test_oracle_golden.py::test_weights_file_written_by_the_reference loads the bytes with this package's
load_weights, compares every tensor, and checks that this package's save_weights reproduces the file byte for byte.
Run here (needs /root/reference): python oracle/make_golden_weights.py"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

MINI_CFG = """[net]
batch=1
height=64
width=64
channels=3

[convolutional]
batch_normalize=1
filters=8
size=3
stride=1
pad=1
activation=leaky

[maxpool]
size=2
stride=2

[convolutional]
batch_normalize=1
filters=16
size=3
stride=1
pad=1
activation=leaky

[convolutional]
batch_normalize=1
filters=8
size=1
stride=1
pad=1
activation=leaky

[maxpool]
size=2
stride=2

[convolutional]
batch_normalize=1
filters=24
size=3
stride=1
pad=1
activation=leaky

[route]
layers=-3

[convolutional]
batch_normalize=1
filters=4
size=1
stride=1
pad=1
activation=leaky

[reorg]
stride=2

[route]
layers=-1,-4

[convolutional]
batch_normalize=1
filters=32
size=3
stride=1
pad=1
activation=leaky

[convolutional]
filters=16
size=1
stride=1
pad=1
activation=linear

[region]
anchors = 1.0,1.5, 2.5,2.0
bias_match=1
classes=3
coords=4
num=2
softmax=1
jitter=.3
rescore=1
object_scale=5
noobject_scale=1
class_scale=1
coord_scale=1
absolute=1
thresh = .6
random=0
"""


def main():
    ref = ref_shim.load_reference()
    d = tempfile.mkdtemp()
    cfg = os.path.join(d, 'mini.cfg')
    with open(cfg, 'w') as f:
        f.write(MINI_CFG)
    torch.manual_seed(21)
    model = ref_shim.quiet(ref['nets'].Darknet, cfg)
    g = torch.Generator().manual_seed(22)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    model.seen = 4242
    path = os.path.join(d, 'mini.weights')
    ref_shim.quiet(model.save_weights, path)
    raw = np.fromfile(path, dtype=np.uint8)
    out = {'cfg': np.frombuffer(MINI_CFG.encode(), dtype=np.uint8), 'file_bytes': raw, 'seen': np.array(4242)}
    for k, v in model.state_dict().items():
        out['sd.' + k] = v.numpy()
    # the reference's loader must read its own file back (sanity of the fixture)
    model2 = ref_shim.quiet(ref['nets'].Darknet, cfg)
    ref_shim.quiet(model2.load_weights, path)
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        if not k.endswith('num_batches_tracked'):
            assert torch.equal(a, b), k
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'weights_mini.npz'), **out)
    print('wrote weights_mini.npz: %d file bytes, %d tensors' % (raw.size, len(out) - 3))


if __name__ == '__main__':
    main()
