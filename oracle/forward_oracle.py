"""fp32 PyTorch restatement of Darknet.forward (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows src/nets.py:720-774 block by block with the same library calls the reference's modules make:
MaskedConv2d.forward = F.conv2d(x, weight*mask) (src/pruning/weightPruning/layers.py:53-64), nn.BatchNorm2d in eval
mode (src/nets.py:802), nn.LeakyReLU(0.1) (:809), nn.MaxPool2d(2,2) (:821), Reorg (:648-667), torch.cat (:745).
Works from a state_dict with the reference's key names, on CPU (the timed CPU baseline) or on CUDA (the fp32
reference of the GPU parity tests).
"""
import torch
import torch.nn.functional as F


def reorg(x, stride=2):
    B, C, H, W = x.shape
    hs = ws = stride
    x = x.view(B, C, H // hs, hs, W // ws, ws).transpose(3, 4).contiguous()
    x = x.view(B, C, (H // hs) * (W // ws), hs * ws).transpose(2, 3).contiguous()
    x = x.view(B, C, hs * ws, H // hs, W // ws).transpose(1, 2).contiguous()
    return x.view(B, hs * ws * C, H // hs, W // ws)


def darknet_forward_fp32(blocks, state, x, keep_outputs=False, masks_applied=True):
    """blocks: parse_cfg() output; state: state_dict (reference key names).  Returns (head, outputs dict)."""
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
            w = state[pre + 'conv%d.weight' % conv_id]
            mk = state.get(pre + 'conv%d.mask' % conv_id)
            if mk is not None:
                w = w * mk.to(w.device)
            k = int(block['size'])
            pad = (k - 1) // 2 if int(block['pad']) else 0
            bias = state.get(pre + 'conv%d.bias' % conv_id)
            x = F.conv2d(x, w, bias, 1, pad)
            if int(block['batch_normalize']):
                x = F.batch_norm(x, state[pre + 'bn%d.running_mean' % conv_id], state[pre + 'bn%d.running_var' % conv_id],
                                 state[pre + 'bn%d.weight' % conv_id], state[pre + 'bn%d.bias' % conv_id], False, 0.1, 1e-5)
            if block['activation'] == 'leaky':
                x = F.leaky_relu(x, 0.1)
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
        outputs[ind] = x
    return x, (outputs if keep_outputs else None)


def kaiming_normal_init_(model, seed=7):
    """Variance-preserving weights (SURVEY.md §7 hard part 7): conv W ~ N(0, 2/(1.01*fan_in)); default-init signal
    decays ~0.4x per block, which would make logit-level parity vacuous."""
    g = torch.Generator().manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() == 4:
            fan_in = p.shape[1] * p.shape[2] * p.shape[3]
            std = (2.0 / (1.01 * fan_in)) ** 0.5
            p.data.copy_(torch.randn(p.shape, generator=g) * std)
    return model


def randomize_bn_(model, seed=1):
    """rand-BN variant: exercises BN folding and constant-channel handling of the shrunk network."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            n = m.num_features
            m.weight.data.copy_(torch.rand(n, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_mean.copy_(torch.randn(n, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=g) + 0.5)
    return model
