"""Import the UNMODIFIED reference from /root/reference inside this (GPU-less) container.

Used only by oracle/make_golden.py to generate tests/golden/ fixtures.  /root/reference does not exist on the GPU
box, so nothing that runs there imports this module.  Shims:
  1. empty ``matplotlib`` / ``matplotlib.pyplot`` modules (the reference imports them for plotting only);
  2. ``torch.Tensor.cuda`` -> identity so get_region_boxes (src/nets2_utils.py:163-172) runs on CPU;
  3. ``torch.sort`` forced stable inside ``nms`` so tie order is defined (ascending candidate index).
"""
import contextlib
import io
import sys
import types

REFERENCE_ROOT = '/root/reference'
_loaded = {}


def load_reference():
    if _loaded:
        return _loaded
    import torch
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    torch.Tensor.cuda = lambda self, *a, **k: self
    with contextlib.redirect_stdout(io.StringIO()):
        import src.nets as nets
        import src.nets2_utils as nets2_utils
        import src.pruning.weightPruning.methods as methods
        import src.pruning.weightPruning.utils as putils
        import src.pruning.weightPruning.layers as layers
    _loaded.update(nets=nets, nets2_utils=nets2_utils, methods=methods, putils=putils, layers=layers,
                   cfg=REFERENCE_ROOT + '/src/yolov2-voc.cfg')
    return _loaded


@contextlib.contextmanager
def stable_sort():
    import torch
    orig = torch.sort

    def _sort(x, *a, **k):
        k['stable'] = True
        return orig(x, *a, **k)

    torch.sort = _sort
    try:
        yield
    finally:
        torch.sort = orig


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
