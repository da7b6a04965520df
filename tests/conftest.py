import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope='session')
def cfg_path(tmp_path_factory):
    import modelcompression_b200 as mc
    return mc.write_yolov2_voc_cfg(str(tmp_path_factory.mktemp('cfg') / 'yolov2-voc.cfg'))


def make_darknet(cfg_path, seed=0, kn=False, randbn=False, device=None):
    import torch
    import modelcompression_b200 as mc
    from oracle import forward_oracle
    torch.manual_seed(seed)
    model = mc.Darknet(cfg_path)
    if kn:
        forward_oracle.kaiming_normal_init_(model, 7)
    if randbn:
        forward_oracle.randomize_bn_(model, 1)
    if device is not None:
        model = model.to(device)
    return model.eval()


class SmallNet(object):
    """Mirror of oracle/make_golden.py::SmallNet (4-D, 2-D and 1-D parameters)."""

    @staticmethod
    def build(seed=11):
        import torch

        class _Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.c1 = torch.nn.Conv2d(3, 8, 3)
                self.c2 = torch.nn.Conv2d(8, 16, 3)
                self.c3 = torch.nn.Conv2d(16, 12, 1)
                self.fc = torch.nn.Linear(12, 10)

        torch.manual_seed(seed)
        return _Net()
