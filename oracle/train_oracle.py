"""fp32 PyTorch restatement of the masked retrain step (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows src/train.py:221-235 on the Darknet of src/nets.py:720-822 with model.train(): the same library calls the
reference's modules make — F.conv2d(x, weight*mask) (src/pruning/weightPruning/layers.py:53-64), nn.BatchNorm2d with
batch statistics (eps 1e-5, momentum 0.1, src/nets.py:802), nn.LeakyReLU(0.1) (:809), nn.MaxPool2d(2,2) (:821),
Reorg (:648-667), torch.cat (:745) — recorded by autograd, then backward.  The loss is the synthetic one of SURVEY.md
§8d (config 3): loss = (y * g).sum() for a fixed upstream gradient g (RegionLoss is row N3).
Pinned against the unmodified reference by oracle/make_golden_train.py (tests/golden/train_step.npz).
"""
import torch
import torch.nn.functional as F

from .forward_oracle import reorg


class _Round(torch.autograd.Function):
    """bf16 rounding of the forward value and/or of the gradient flowing back through this point."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(torch.bfloat16).to(x.dtype) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).to(g.dtype) if ctx.bwd else g), None, None


def _r(x, on, fwd=True, bwd=True):
    return _Round.apply(x, fwd, bwd) if on else x


def train_forward_fp32(blocks, params, buffers, x, update_running=True, outputs_out=None, emulate_bf16=False):
    """emulate_bf16=False is the reference's arithmetic (pinned).  emulate_bf16=True inserts bf16 roundings at exactly
    the points where the B200 kernels store bf16 (image and masked weights as GEMM operands; conv output z; activation
    a; and, going back, dY, dz and da), all other math staying fp32 — the same arithmetic contract as the kernels, so
    what remains is summation order.  Needed because this random-init network in train mode is chaotic: the fp32
    oracle's own logits move by 12 % when only the input image is rounded to bf16."""
    e = emulate_bf16
    x = _r(x, e, True, False)
    """params: dict name -> tensor (requires_grad leaves) with the reference's state_dict names; buffers: dict with
    running_mean / running_var / mask entries (updated in place like nn.BatchNorm2d when update_running)."""
    outputs = {}
    ind = -2
    conv_id = 0
    for block in blocks:
        ind += 1
        t = block['type']
        if t == 'net':
            continue
        if t == 'convolutional':
            conv_id += 1
            pre = 'models.%d.' % ind
            w = params[pre + 'conv%d.weight' % conv_id]
            mk = buffers.get(pre + 'conv%d.mask' % conv_id)
            if mk is not None:
                w = w * mk
            w = _r(w, e, True, False)
            k = int(block['size'])
            pad = (k - 1) // 2 if int(block['pad']) else 0
            bias = params.get(pre + 'conv%d.bias' % conv_id)
            if not e:
                x = F.conv2d(x, w, bias, 1, pad)                             # the reference's call, bit for bit
            elif int(block['batch_normalize']):
                x = _r(F.conv2d(x, w, bias, 1, pad), e)                      # z is stored in bf16, dz likewise
            else:
                x = _r(F.conv2d(x, w, None, 1, pad), e, False, True)         # head: fp32 logits, dY rounded for the GEMMs
                if bias is not None:
                    x = x + bias.view(1, -1, 1, 1)
            if int(block['batch_normalize']):
                rm = buffers[pre + 'bn%d.running_mean' % conv_id] if update_running else None
                rv = buffers[pre + 'bn%d.running_var' % conv_id] if update_running else None
                x = F.batch_norm(x, rm, rv, params[pre + 'bn%d.weight' % conv_id], params[pre + 'bn%d.bias' % conv_id],
                                 True, 0.1, 1e-5)
            if block['activation'] == 'leaky':
                x = F.leaky_relu(x, 0.1)
            if int(block['batch_normalize']):
                x = _r(x, e)                                                 # a is stored in bf16, da likewise
        elif t == 'maxpool':
            x = F.max_pool2d(x, int(block['size']), int(block['stride']))
        elif t == 'reorg':
            x = reorg(x, int(block['stride']))
        elif t == 'route':
            layers = [int(i) if int(i) > 0 else int(i) + ind for i in block['layers'].split(',')]
            x = outputs[layers[0]] if len(layers) == 1 else torch.cat((outputs[layers[0]], outputs[layers[1]]), 1)
        elif t == 'region':
            continue
        else:
            raise NotImplementedError(t)
        if outputs_out is not None and x.requires_grad:
            x.retain_grad()  # tests read the gradient of every block output
        outputs[ind] = x
    if outputs_out is not None:
        outputs_out.update(outputs)
    return x


def train_step_fp32(blocks, state, x, g, emulate_bf16=False, dtype=torch.float32):
    """One forward + backward of loss = (y*g).sum().  state: state_dict (reference key names; masks included when
    set).  Returns (y, grads dict name -> tensor, updated running-stat dict).  dtype=torch.float64 runs the same
    graph in double precision (with emulate_bf16 the roundings stay where they are): the distance between the float32
    and float64 runs measures how far two CORRECT implementations that only differ in accumulation error drift apart
    on this chaotic network — the yardstick for the end-to-end parity gate."""
    params, buffers = {}, {}
    for k, v in state.items():
        if k.endswith('.weight') or k.endswith('.bias'):
            params[k] = v.detach().clone().to(dtype).requires_grad_(True)
        elif v.is_floating_point():
            buffers[k] = v.detach().clone().to(dtype)
        else:
            buffers[k] = v.detach().clone()
    x, g = x.to(dtype), g.to(dtype)
    y = train_forward_fp32(blocks, params, buffers, x, emulate_bf16=emulate_bf16)
    (y * g).sum().backward()
    grads = {k: p.grad for k, p in params.items()}
    return y.detach(), grads, buffers
