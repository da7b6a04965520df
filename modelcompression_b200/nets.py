"""Drop-in for the YOLOv2 part of src/nets.py of the reference: parse_cfg (:39-73), Reorg (:648-667),
MaxPoolStride1 (:640-646), GlobalAvgPool2d (:669-680), EmptyModule (:683-688), Darknet (:692-1061), getYOLOv2
(:1069-1074), and the darknet ``.weights`` reader/writer (:194-279, :897-948, :1007-1051).

Module tree, attribute names and ``state_dict`` keys are the reference's (``models.{i}.conv{n}.weight``,
``models.{i}.bn{n}.*``, ``models.{i}.conv{n}.mask``), so checkpoints interchange.  What differs is the forward:
``Darknet.forward`` hands the whole graph to the B200 engine (engine.py: tcgen05 implicit-GEMM convs with folded
BatchNorm + leaky-ReLU epilogues, fused reorg/concat addressing) instead of walking ``nn.Module``s.
"""
import numpy as np
import torch
import torch.nn as nn

from .pruning.weightPruning.layers import MaskedConv2d
from .pruning.weightPruning.methods import quick_filter_prune, weight_prune  # noqa: F401  (re-exported like nets.py:18)
from .pruning.weightPruning.utils import are_masks_consistent, prune_rate  # noqa: F401


def parse_cfg(cfgfile, verbose=0):
    """nets.py:39-73 — darknet cfg -> list of dict blocks.  ``[convolutional]`` defaults batch_normalize=0; a key
    called ``type`` inside a block is stored as ``_type`` (the section name owns ``type``)."""
    blocks = []
    block = None
    with open(cfgfile, 'r') as fp:
        for raw in fp:
            line = raw.rstrip()
            if line == '' or line[0] == '#':
                continue
            if line[0] == '[':
                if block is not None:
                    if verbose:
                        print(' - block : ', block)
                    blocks.append(block)
                block = {'type': line.lstrip('[').rstrip(']')}
                if block['type'] == 'convolutional':
                    block['batch_normalize'] = 0
            else:
                key, value = line.split('=')
                key = key.strip()
                if key == 'type':
                    key = '_type'
                block[key] = value.strip()
    if block is not None:
        blocks.append(block)
    return blocks


class MaxPoolStride1(nn.Module):
    """nets.py:640-646 (not used by yolov2-voc.cfg; out of the B200 path)."""

    def forward(self, x):
        raise NotImplementedError("MaxPoolStride1 is not used by yolov2-voc.cfg (SURVEY.md §2 #4)")


class Reorg(nn.Module):
    """nets.py:648-667: out[b,(i*s+j)*C+c,y,x] = in[b,c,s*y+i,s*x+j].  Inside ``Darknet`` the engine fuses this into
    the producing conv's store addressing (MC_EPI_REORG2); a direct call of the module runs mc_reorg_nchw."""

    def __init__(self, stride=2):
        super(Reorg, self).__init__()
        self.stride = stride

    def forward(self, x):
        """Stand-alone call (float32 NCHW CUDA tensor in and out, the reference's layout): one libmcb200 kernel."""
        from . import _lib
        lib = _lib.load()
        _lib.require_cuda(x, "Reorg.forward")
        assert x.dim() == 4
        B, C, H, W = x.shape
        s = int(self.stride)
        assert H % s == 0 and W % s == 0
        xin = x.detach().float().contiguous()
        out = torch.empty(B, s * s * C, H // s, W // s, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.mc_reorg_nchw(xin.data_ptr(), out.data_ptr(), B, C, H, W, s, _lib.stream_ptr()), "mc_reorg_nchw")
        return out


class GlobalAvgPool2d(nn.Module):
    def forward(self, x):
        raise NotImplementedError("GlobalAvgPool2d is not used by yolov2-voc.cfg (SURVEY.md §2 #4)")


class EmptyModule(nn.Module):
    """Placeholder for route/shortcut blocks (nets.py:683-688)."""

    def forward(self, x):
        return x


class RegionLoss(nn.Module):
    """nets.py:442-636.  Hyper-parameters as in the reference (filled by create_network :873-889).  ``forward`` is the
    device-resident, vectorised restatement in region_loss.py: no copy to the CPU, no Python loop
    over images / boxes / anchors; loss and gradient equal the reference's to float32 round-off
    (oracle/make_golden_region.py)."""

    def __init__(self, num_classes=20, anchor_list=None, anchors_cell=5):
        super(RegionLoss, self).__init__()
        anchor_list = list(anchor_list) if anchor_list is not None else []
        self.num_classes = num_classes
        self.anchors = anchor_list
        self.num_anchors = anchors_cell
        self.anchor_step = int(len(anchor_list) / anchors_cell) if anchors_cell else 0
        self.coord_scale = 1
        self.noobject_scale = 1
        self.object_scale = 1
        self.class_scale = 1
        self.thresh = 0.6
        self.seen = 0

    def forward(self, output, target, verbose=0):
        from .region_loss import region_loss
        return region_loss(output, target, self.anchors, self.num_anchors, self.num_classes, self.coord_scale,
                           self.noobject_scale, self.object_scale, self.class_scale, self.thresh)


class Darknet(nn.Module):
    """nets.py:692-1061."""

    def __init__(self, cfgfile, verbose=0):
        super(Darknet, self).__init__()
        self.blocks = parse_cfg(cfgfile)
        self.models = self.create_network(self.blocks)
        self.loss = self.models[len(self.models) - 1]

        self.width = int(self.blocks[0]['width'])
        self.height = int(self.blocks[0]['height'])

        if self.blocks[(len(self.blocks) - 1)]['type'] == 'region':
            self.anchors = self.loss.anchors
            self.num_anchors = self.loss.num_anchors
            self.anchor_step = self.loss.anchor_step
            self.num_classes = self.loss.num_classes
            if verbose:
                print('  -- [Darknet] anchors', self.anchors, 'num_anchors', self.num_anchors, 'num_classes',
                      self.num_classes)

        self.header = torch.IntTensor([0, 0, 0, 0])
        self.seen = 0
        # B200 engine state (not part of state_dict)
        self._b200_plan = None
        self._b200_plan_key = None
        self._b200_watch = None
        self.b200_shrink = True      # physically drop filters whose (masked) weights are all zero
        self.b200_keep_blocks = False  # keep every block's activation buffer alive for per-block parity checks

    # ------------------------------------------------------------------ graph construction (nets.py:779-895)
    def create_network(self, blocks):
        models = nn.ModuleList()
        prev_filters = 3
        out_filters = []
        conv_id = 0
        for block in blocks:
            btype = block['type']
            if btype == 'net':
                prev_filters = int(block['channels'])
                continue
            elif btype == 'convolutional':
                conv_id += 1
                batch_normalize = int(block['batch_normalize'])
                filters = int(block['filters'])
                kernel_size = int(block['size'])
                stride = int(block['stride'])
                pad = int((kernel_size - 1) / 2) if int(block['pad']) else 0
                activation = block['activation']
                model = nn.Sequential()
                if batch_normalize:
                    model.add_module('conv{0}'.format(conv_id),
                                     MaskedConv2d(prev_filters, filters, kernel_size, stride, pad, bias=False))
                    model.add_module('bn{0}'.format(conv_id), nn.BatchNorm2d(filters))
                else:
                    model.add_module('conv{0}'.format(conv_id),
                                     MaskedConv2d(prev_filters, filters, kernel_size, stride, pad))
                if activation == 'leaky':
                    model.add_module('leaky{0}'.format(conv_id), nn.LeakyReLU(0.1, inplace=True))
                elif activation == 'relu':
                    model.add_module('relu{0}'.format(conv_id), nn.ReLU(inplace=True))
                prev_filters = filters
                out_filters.append(prev_filters)
                models.append(model)
            elif btype == 'maxpool':
                pool_size = int(block['size'])
                stride = int(block['stride'])
                model = nn.MaxPool2d(pool_size, stride) if stride > 1 else MaxPoolStride1()
                out_filters.append(prev_filters)
                models.append(model)
            elif btype == 'avgpool':
                out_filters.append(prev_filters)
                models.append(GlobalAvgPool2d())
            elif btype == 'reorg':
                stride = int(block['stride'])
                prev_filters = stride * stride * prev_filters
                out_filters.append(prev_filters)
                models.append(Reorg(stride))
            elif btype == 'route':
                layers = block['layers'].split(',')
                ind = len(models)
                layers = [int(i) if int(i) > 0 else int(i) + ind for i in layers]
                if len(layers) == 1:
                    prev_filters = out_filters[layers[0]]
                elif len(layers) == 2:
                    assert (layers[0] == ind - 1)
                    prev_filters = out_filters[layers[0]] + out_filters[layers[1]]
                out_filters.append(prev_filters)
                models.append(EmptyModule())
            elif btype == 'shortcut':
                ind = len(models)
                prev_filters = out_filters[ind - 1]
                out_filters.append(prev_filters)
                models.append(EmptyModule())
            elif btype == 'region':
                loss = RegionLoss()
                loss.anchors = [float(i) for i in block['anchors'].split(',')]
                loss.num_classes = int(block['classes'])
                loss.num_anchors = int(block['num'])
                loss.anchor_step = len(loss.anchors) / loss.num_anchors
                loss.object_scale = float(block['object_scale'])
                loss.noobject_scale = float(block['noobject_scale'])
                loss.class_scale = float(block['class_scale'])
                loss.coord_scale = float(block['coord_scale'])
                out_filters.append(prev_filters)
                models.append(loss)
            else:
                raise NotImplementedError("cfg block type '%s' is not supported by the B200 path" % btype)
        return models

    # ------------------------------------------------------------------ forward (nets.py:720-774)
    def forward(self, x):
        """[B,3,H,W] float32 CUDA -> raw region head [B, A*(5+classes), H/32, W/32] float32 (the region block is
        skipped, nets.py:761-762)."""
        from .engine import darknet_forward
        return darknet_forward(self, x)

    def print_network(self):
        for i, block in enumerate(self.blocks[1:]):
            print('%3d %-14s %s' % (i, block['type'], {k: v for k, v in block.items() if k != 'type'}))

    # ------------------------------------------------------------------ masks (nets.py:1053-1061)
    def masked_convs(self):
        """MaskedConv2d layers in set_masks order (module order == parameter order)."""
        out = []
        for m in self.modules():
            if isinstance(m, nn.Sequential) and len(m) > 0 and getattr(m[0], 'name', None) == 'MaskedConv2d':
                out.append(m[0])
        return out

    def set_masks(self, masks):
        """Assign masks[count] to the count-th MaskedConv2d (module order).  The reference silently ignores every
        error here (bare ``except: pass``, nets.py:1060-1061); a wrong-length list raises instead."""
        convs = self.masked_convs()
        if len(masks) != len(convs):
            raise ValueError("set_masks: got %d masks for %d MaskedConv2d layers" % (len(masks), len(convs)))
        for conv, mask in zip(convs, masks):
            conv.set_mask(mask)
        self._b200_plan = None
        self._b200_watch = None  # new mask buffers: rebuild the change-detector's tensor list

    # ------------------------------------------------------------------ darknet .weights IO
    def _conv_blocks(self, cutoff=None):
        ind = -2
        for bi, block in enumerate(self.blocks):
            ind += 1
            if cutoff is not None and bi > cutoff:
                break
            if block['type'] == 'convolutional':
                yield block, self.models[ind]

    def load_weights(self, weightfile):
        """nets.py:897-948: header int32 x3 (+ seen as int64 when major*10+minor >= 2, else int32), then per conv
        [bn.bias, bn.weight, running_mean, running_var, conv.weight] or [conv.bias, conv.weight], float32."""
        with open(weightfile, 'rb') as f:
            major, minor, revision = np.fromfile(f, dtype=np.int32, count=3)
            if major * 10 + minor >= 2 and major < 1000 and minor < 1000:
                seen = np.fromfile(f, dtype=np.int64, count=1)
            else:
                seen = np.fromfile(f, dtype=np.int32, count=1)
            # The reference reads and DISCARDS the file header (nets.py:899-905): self.header stays [0,0,0,0], so a later
            # save_weights always writes version 0.0.0 with an int32 `seen` — a 16-byte header both loaders read back.
            # Copying a v0.2 file's major/minor here would make save_weights emit "0.2" with a 16-byte header, which every
            # reader then mis-aligns by 4 bytes.  Only the image counter is kept (clamped to what int32 `seen` can hold).
            self.seen = int(seen[0]) & 0x7fffffff

            def read_into(t):
                buf = np.fromfile(f, dtype=np.float32, count=t.numel())
                if buf.size != t.numel():
                    raise IOError("%s: truncated weights file" % weightfile)
                t.data.copy_(torch.from_numpy(buf).view_as(t))

            for block, model in self._conv_blocks():
                conv = model[0]
                if int(block['batch_normalize']):
                    bn = model[1]
                    for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var, conv.weight):
                        read_into(t)
                else:
                    read_into(conv.bias)
                    read_into(conv.weight)
        self._b200_plan = None

    def save_weights(self, outfile, cutoff=0):
        """nets.py:1007-1051: header = IntTensor[4] with header[3] = seen, then the per-conv records."""
        if cutoff <= 0:
            cutoff = len(self.blocks) - 1
        with open(outfile, 'wb') as fp:
            self.header[3] = self.seen
            self.header.numpy().tofile(fp)
            for block, model in self._conv_blocks(cutoff):
                conv = model[0]
                if int(block['batch_normalize']):
                    bn = model[1]
                    tensors = (bn.bias.data, bn.weight.data, bn.running_mean, bn.running_var, conv.weight.data)
                else:
                    tensors = (conv.bias.data, conv.weight.data)
                for t in tensors:
                    t.detach().cpu().numpy().astype(np.float32).tofile(fp)


def getYOLOv2(cfgfile, weightfile):
    """nets.py:1069-1074."""
    model = Darknet(cfgfile)
    model.load_weights(weightfile)
    if torch.cuda.is_available():
        model.cuda()
    return model
