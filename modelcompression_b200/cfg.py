"""Darknet ``.cfg`` text for YOLOv2-VOC (Darknet-19 + passthrough), generated from a compact layer spec.

The reference ships this network as ``src/yolov2-voc.cfg`` and builds it with ``Darknet(cfgfile)``
(src/nets.py:694, parse_cfg :39-73).  Users of the drop-in pass their own cfg file; tests and bench generate this
one so that nothing is copied from, or read out of, the reference tree at run time.
"""
import os
import tempfile

# (filters, kernel) for conv blocks, 'M' = 2x2/2 maxpool
_BACKBONE = [(32, 3), 'M', (64, 3), 'M', (128, 3), (64, 1), (128, 3), 'M', (256, 3), (128, 1), (256, 3), 'M',
             (512, 3), (256, 1), (512, 3), (256, 1), (512, 3), 'M',
             (1024, 3), (512, 1), (1024, 3), (512, 1), (1024, 3), (1024, 3), (1024, 3)]

VOC_ANCHORS = [1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071]


def _conv(filters, size, bn=1, activation='leaky'):
    lines = ['[convolutional]']
    if bn:
        lines.append('batch_normalize=1')
    lines += ['filters=%d' % filters, 'size=%d' % size, 'stride=1', 'pad=1', 'activation=%s' % activation, '']
    return lines


def yolov2_voc_cfg_text(width=416, height=416, classes=20, anchors=None):
    anchors = list(VOC_ANCHORS if anchors is None else anchors)
    num = len(anchors) // 2
    out = ['[net]', 'batch=1', 'subdivisions=1', 'height=%d' % height, 'width=%d' % width, 'channels=3',
           'momentum=0.9', 'decay=0.0005', 'learning_rate=0.001', 'max_batches=80200', 'policy=steps',
           'steps=40000,60000', 'scales=.1,.1', '']
    for item in _BACKBONE:
        if item == 'M':
            out += ['[maxpool]', 'size=2', 'stride=2', '']
        else:
            out += _conv(*item)
    out += ['[route]', 'layers=-9', '']
    out += _conv(64, 1)
    out += ['[reorg]', 'stride=2', '']
    out += ['[route]', 'layers=-1,-4', '']
    out += _conv(1024, 3)
    out += _conv(num * (5 + classes), 1, bn=0, activation='linear')
    out += ['[region]', 'anchors = ' + ', '.join(repr(a) for a in anchors), 'bias_match=1', 'classes=%d' % classes,
            'coords=4', 'num=%d' % num, 'softmax=1', 'jitter=.3', 'rescore=1', 'object_scale=5', 'noobject_scale=1',
            'class_scale=1', 'coord_scale=1', 'absolute=1', 'thresh = .6', 'random=1', '']
    return '\n'.join(out)


def write_yolov2_voc_cfg(path=None, **kw):
    """Write the cfg and return its path (a temp file when ``path`` is None)."""
    text = yolov2_voc_cfg_text(**kw)
    if path is None:
        fd, path = tempfile.mkstemp(prefix='yolov2-voc-', suffix='.cfg')
        os.close(fd)
    with open(path, 'w') as f:
        f.write(text)
    return path
