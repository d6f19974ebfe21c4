"""Pruned-weight transfer on the device: the step right after top-k.

The reference fills the pruned model from the original one with element-wise Python loops,
one tensor assignment per (kept output channel, kept input channel) pair
(/root/reference/utils/load_models.py: load_vgg_model :17-61, load_resnet_model :64-142,
load_resnet_imagenet_model :441-582).  Here the same bookkeeping - which score file selects a
convolution's output channels, which earlier selection its input channels follow, which
BatchNorm vectors ride along - is computed once from the two models' shapes as a list of ops
(`transfer_plan`), and every op is one `dctp_gather_weight` launch
(out[i][j][:] = w[sel_out[i]][sel_in[j]][:]).  Scores -> kept ids -> pruned weights never leave
the GPU.  The loaders' quirks are kept (they decide what the fine-tuning stage starts from):
a selection is not reset after an unpruned VGG convolution; ResNet-50's downsample convolution
takes its input ids from the block's conv2; when the remembered selection is shorter than the
destination's input dimension only that leading part is filled and the rest keeps the pruned
model's own initialisation.

Covered nets: vgg_16_bn, resnet_56, resnet_110, resnet_50 (the other loaders are next).
There is no CPU path: `gather_weight` raises on CPU tensors.
"""
from collections import namedtuple

import torch
import torch.nn as nn

from . import _lib

# kind: 'copy' (tensor taken over as is) or 'gather'; out / inp: score-file stems whose kept ids select the
# output / input channels (None = all channels, in order)
Op = namedtuple('Op', 'kind name out inp')

SUPPORTED = ('vgg_16_bn', 'resnet_56', 'resnet_110', 'resnet_50')
_BN_PARTS = ('.weight', '.bias', '.running_mean', '.running_var')


def _convs_and_linears(model):
    convs, linears = [], []
    for name, module in model.named_modules():
        name = name.replace('module.', '')
        if isinstance(module, nn.Conv2d):
            convs.append(name)
        elif isinstance(module, nn.Linear):
            linears.append(name)
    return convs, linears


def transfer_plan(net_name, pruned_model, ori_shapes):
    """The ops that turn the original state dict into the pruned model's, in the reference's order.
    `ori_shapes`: {state-dict key: shape} of the unpruned net."""
    new_shapes = {k: tuple(v.shape) for k, v in pruned_model.state_dict().items()}
    convs, linears = _convs_and_linears(pruned_model)
    ops = []

    def width(shapes, conv):
        return shapes[conv + '.weight'][0]

    if net_name == 'vgg_16_bn':                                  # load_models.py:17-61
        last = None
        for cnt, conv in enumerate(convs, start=1):
            w = conv + '.weight'
            if width(ori_shapes, conv) != width(new_shapes, conv):
                stem = 'imp_conv%d' % cnt
                ops.append(Op('gather', w, stem, last))
                last = stem
            elif last is not None:
                ops.append(Op('gather', w, None, last))          # (:53-57: the selection is NOT reset here)
            else:
                ops.append(Op('copy', w, None, None))
                last = None
        return ops

    if net_name in ('resnet_56', 'resnet_110'):                  # load_models.py:64-142
        last, cnt, listed = None, 1, set()
        for stage in range(3):
            for b in range(9 if net_name == 'resnet_56' else 18):
                for l in (1, 2):
                    cnt += 1
                    conv = 'layer%d.%d.conv%d' % (stage + 1, b, l)
                    w = conv + '.weight'
                    listed.add(w)
                    if width(ori_shapes, conv) != width(new_shapes, conv):
                        stem = 'imp_conv%d' % cnt
                        ops.append(Op('gather', w, stem, last))
                        last = stem
                    elif last is not None:
                        ops.append(Op('gather', w, None, last))
                        last = None
                    else:
                        ops.append(Op('copy', w, None, None))
                        last = None
        for conv in convs:                                       # :129-136
            if 'shortcut' not in conv and conv + '.weight' not in listed:
                ops.append(Op('copy', conv + '.weight', None, None))
        for lin in linears:                                      # :138-140
            ops.append(Op('copy', lin + '.weight', None, None))
            ops.append(Op('copy', lin + '.bias', None, None))
        return ops

    if net_name == 'resnet_50':                                  # load_models.py:441-582
        listed = set()

        def conv_and_bn(conv, bn, stem, last, record_last):
            w = conv + '.weight'
            listed.add(w)
            if width(ori_shapes, conv) != width(new_shapes, conv):
                ops.append(Op('gather', w, stem, last))
                for part in _BN_PARTS:
                    ops.append(Op('gather', bn + part, stem, None))
                new_last = stem if record_last else last
            elif last is not None:
                ops.append(Op('gather', w, None, last))
                for part in _BN_PARTS:
                    ops.append(Op('copy', bn + part, None, None))
                new_last = None if record_last else last
            else:
                ops.append(Op('copy', w, None, None))
                for part in _BN_PARTS:
                    ops.append(Op('copy', bn + part, None, None))
                new_last = None if record_last else last
            ops.append(Op('copy', bn + '.num_batches_tracked', None, None))
            return new_last

        last = conv_and_bn('conv1', 'bn1', 'imp_conv1', None, True)
        cnt = 2
        for stage, repeat in enumerate((3, 4, 6, 3)):
            for b in range(repeat):
                base = 'layer%d.%d.' % (stage + 1, b)
                for l in range(4 if b == 0 else 3):
                    if b == 0 and l == 2:
                        conv, bn, record = base + 'downsample.0', base + 'downsample.1', False
                    elif b == 0 and l == 3:
                        conv, bn, record = base + 'conv3', base + 'bn3', True
                    else:
                        conv, bn, record = base + 'conv%d' % (l + 1), base + 'bn%d' % (l + 1), True
                    last = conv_and_bn(conv, bn, 'imp_conv%d' % cnt, last, record)
                    cnt += 1
        for conv in convs:                                       # :571-575
            if conv + '.weight' not in listed:
                ops.append(Op('copy', conv + '.weight', None, None))
        for lin in linears:                                      # :577-579
            ops.append(Op('copy', lin + '.weight', None, None))
            ops.append(Op('copy', lin + '.bias', None, None))
        return ops

    raise ValueError('weight transfer is not implemented for %r yet (supported: %s)' % (net_name, ', '.join(SUPPORTED)))


def gather_weight(w, sel_out=None, sel_in=None):
    """out[i][j][...] = w[sel_out[i]][sel_in[j]][...] on the device (None = all channels).  `w`: CUDA fp32 tensor with
    1 (vector), 2 (linear) or 4 (convolution) dimensions; selections: int64 CUDA tensors of ascending channel ids."""
    if not w.is_cuda:
        raise RuntimeError('gather_weight runs on CUDA only; there is no CPU fallback')
    if w.dtype != torch.float32:
        raise TypeError('gather_weight moves float32 tensors (got %s)' % w.dtype)
    lib = _lib.load()
    w = w.contiguous()
    c_out = w.shape[0]
    c_in = w.shape[1] if w.dim() >= 2 else 1
    inner = 1
    for d in w.shape[2:]:
        inner *= d
    if sel_in is not None and w.dim() < 2:
        raise ValueError('a vector has no input channels to select')

    def ids(sel):
        if sel is None:
            return None
        sel = sel.to(device=w.device, dtype=torch.int64).contiguous()
        return sel

    so, si = ids(sel_out), ids(sel_in)
    k_out = c_out if so is None else so.numel()
    k_in = c_in if si is None else si.numel()
    shape = (k_out,) + ((k_in,) + tuple(w.shape[2:]) if w.dim() >= 2 else ())
    out = torch.empty(shape, dtype=torch.float32, device=w.device)
    if out.numel() == 0:                                         # a layer pruned to nothing (k = int(C * (1 - r)) can be 0)
        return out
    _lib.check(lib.dctp_gather_weight(_lib.ptr(w), c_out, c_in, inner, _lib.ptr(so) if so is not None else None, k_out,
                                      _lib.ptr(si) if si is not None else None, k_in, _lib.ptr(out), _lib.current_stream()))
    return out


def apply_plan(plan, ori_state, new_state, kept, gather=gather_weight):
    """Run the ops.  `kept`: {score-file stem: int64 ids}.  Tensors not named by any op keep the pruned model's values."""
    for op in plan:
        src = ori_state[op.name]
        if op.kind == 'copy':
            new_state[op.name] = src.clone()
            continue
        g = gather(src, kept[op.out] if op.out is not None else None, kept[op.inp] if op.inp is not None else None)
        dst = new_state[op.name]
        if tuple(g.shape) == tuple(dst.shape):
            new_state[op.name] = g
        else:                                                    # the remembered selection is shorter than the input dimension
            dst = dst.clone()
            dst[:g.shape[0], :g.shape[1]] = g
            new_state[op.name] = dst
    return new_state


def transfer_weights(net_name, pruned_model, ori_state, kept, check=True):
    """Fill `pruned_model` (on a CUDA device) from the unpruned `ori_state` with the kept channels of every pruned
    layer; `kept` as returned by `topk.kept_channels` ([(Selection, ids)]) or a {stem: ids} mapping."""
    device = next(pruned_model.parameters()).device
    if device.type != 'cuda':
        raise RuntimeError('transfer_weights needs the pruned model on a CUDA device (got %s); there is no CPU fallback' % device)
    if not isinstance(kept, dict):
        kept = {sel.stem: ids for sel, ids in kept}
    kept = {k: torch.as_tensor(v, dtype=torch.int64).to(device) for k, v in kept.items()}
    ori = {k: v.to(device) for k, v in ori_state.items()}
    plan = transfer_plan(net_name, pruned_model, {k: tuple(v.shape) for k, v in ori.items()})
    state = apply_plan(plan, ori, dict(pruned_model.state_dict()), kept)
    if check:
        _lib.check(_lib.load().dctp_check(_lib.current_stream()))
    pruned_model.load_state_dict(state)
    return plan
