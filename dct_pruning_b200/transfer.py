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

Covered nets: all seven of the reference (vgg_16_bn, resnet_56, resnet_110, resnet_50, densenet_40, googlenet, u2netp).
There is no CPU path: `gather_weight` raises on CPU tensors.
"""
from collections import namedtuple

import torch
import torch.nn as nn

from . import _lib

# kind: 'copy' (tensor taken over as is) or 'gather'.  out: score-file stem whose kept ids select the output channels
# (None = all, in order).  inp: which input channels are filled, None = all, else a tuple of pieces that are
# concatenated - Piece(stem, n, offset) stands for `kept[stem] + offset`, or `range(n) + offset` when stem is None
# (DenseNet and GoogLeNet feed a convolution with the concatenation of several earlier layers' kept channels).
Op = namedtuple('Op', 'kind name out inp')
Piece = namedtuple('Piece', 'stem n offset')


def _one(stem):
    return None if stem is None else (Piece(stem, 0, 0),)

SUPPORTED = ('vgg_16_bn', 'resnet_56', 'resnet_110', 'resnet_50', 'densenet_40', 'googlenet', 'u2netp')
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


def transfer_plan(net_name, pruned_model, ori_shapes, origin_rates=None):
    """The ops that turn the original state dict into the pruned model's, in the reference's order.
    `ori_shapes`: {state-dict key: shape} of the net the weights come from; `origin_rates`: its compress rates when it is
    itself a pruned net (only GoogLeNet's loader takes them: its `cpr` argument, used by prune_dynamic.py:154)."""
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
                ops.append(Op('gather', w, stem, _one(last)))
                last = stem
            elif last is not None:
                ops.append(Op('gather', w, None, _one(last)))          # (:53-57: the selection is NOT reset here)
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
                        ops.append(Op('gather', w, stem, _one(last)))
                        last = stem
                    elif last is not None:
                        ops.append(Op('gather', w, None, _one(last)))
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
                ops.append(Op('gather', w, stem, _one(last)))
                for part in _BN_PARTS:
                    ops.append(Op('gather', bn + part, stem, None))
                new_last = stem if record_last else last
            elif last is not None:
                ops.append(Op('gather', w, None, _one(last)))
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

    if net_name == 'densenet_40':                                # load_models.py:383-438
        last = []                                                # pieces of the running input-channel list
        for cnt, conv in enumerate(convs, start=1):
            w = conv + '.weight'
            C = width(ori_shapes, conv)
            if C != width(new_shapes, conv):
                stem = 'imp_conv%d' % cnt
                ops.append(Op('gather', w, stem, tuple(last)))   # (the first convolution meets an empty list: nothing is written)
                select = Piece(stem, 0, 0)
            else:
                ops.append(Op('gather', w, None, tuple(last)))
                select = Piece(None, C, 0)
            if cnt in (1, 14, 27):                               # :434-438: a dense stage starts over, the others append
                last = [select]
            else:
                last = last + [select._replace(offset=cnt * 12 - (cnt - 1) // 13 * 12)]
        return ops

    if net_name == 'googlenet':                                  # load_models.py:146-380 (cpr=None, as load_model calls it)
        filters = [[64, 128, 32, 32], [128, 192, 96, 64], [192, 208, 48, 64], [160, 224, 64, 64], [128, 256, 64, 64],
                   [112, 288, 64, 64], [256, 320, 128, 128], [256, 320, 128, 128], [384, 384, 128, 128]]
        if origin_rates is not None:                             # :160-163: the source net's own branch widths
            for i, f in enumerate(filters):
                f[1] = int(f[1] * (1 - origin_rates[i + 1]))
                f[2] = int(f[2] * (1 - origin_rates[i + 1]))
        listed_convs, listed_bns = set(), set()
        cur_last, cnt = [], 0

        def in_pieces(w, remembered):                            # input channels of a convolution: remembered list or all
            c = ori_shapes[w][1]
            return tuple(remembered) if c != new_shapes[w][1] else (Piece(None, c, 0),)

        def out_choice(w, stem):                                 # (stem or None, piece describing the kept outputs)
            c = ori_shapes[w][0]
            return (stem, Piece(stem, 0, 0)) if c != new_shapes[w][0] else (None, Piece(None, c, 0))

        for name, module in pruned_model.named_modules():
            name = name.replace('module.', '')
            if type(module).__name__ == 'Inception':
                cnt += 1
                f = filters[cnt - 2]
                listed_bns.update(name + i for i in ('.branch3x3.4', '.branch5x5.4', '.branch5x5.7'))
                last, cur_last = list(cur_last), []
                for idx in ('.branch1x1.0', '.branch3x3.0', '.branch5x5.0', '.branch_pool.1'):        # :201-237 inputs only
                    w = name + idx + '.weight'
                    listed_convs.add(name + idx)
                    ops.append(Op('gather', w, None, in_pieces(w, last)))
                    if '1x1' in idx:
                        cur_last.append(Piece(None, new_shapes[w][0], 0))
                    elif 'pool' in idx:
                        cur_last.append(Piece(None, new_shapes[w][0], f[0] + f[1] + f[2]))
                five = None
                for idx, branch in (('.branch3x3.3', '_n3x3'), ('.branch5x5.3', '_n5x5')):            # :239-277 outputs only
                    w = name + idx + '.weight'
                    listed_convs.add(name + idx)
                    stem, piece = out_choice(w, 'imp_conv%d%s' % (cnt, branch))
                    ops.append(Op('gather', w, stem, None))
                    if branch == '_n3x3':
                        cur_last.append(piece._replace(offset=f[0]))
                    else:
                        five = piece                             # the inputs of branch5x5.6 follow this selection
                w = name + '.branch5x5.6.weight'                                                        # :279-322 both
                listed_convs.add(name + '.branch5x5.6')
                stem, piece = out_choice(w, 'imp_conv%d_n5x5' % cnt)
                cur_last.append(piece._replace(offset=f[0] + f[1]))
                ops.append(Op('gather', w, stem, in_pieces(w, [five])))
            elif name == 'pre_layers':                                                                 # :324-359
                cnt += 1
                listed_bns.add('pre_layers.1')
                listed_convs.add('pre_layers.0')
                w = 'pre_layers.0.weight'
                if ori_shapes[w][0] != new_shapes[w][0]:
                    ops.append(Op('gather', w, 'imp_conv%d' % cnt, None))
                    cur_last = [Piece('imp_conv%d' % cnt, 0, 0)]
        for name, module in pruned_model.named_modules():                                              # :361-380
            name = name.replace('module.', '')
            if isinstance(module, nn.Conv2d) and name not in listed_convs:
                ops.append(Op('copy', name + '.weight', None, None))
                ops.append(Op('copy', name + '.bias', None, None))
            elif isinstance(module, nn.BatchNorm2d) and name not in listed_bns:
                for part in _BN_PARTS:
                    ops.append(Op('copy', name + part, None, None))
            elif isinstance(module, nn.Linear):
                ops.append(Op('copy', name + '.weight', None, None))
                ops.append(Op('copy', name + '.bias', None, None))
        return ops

    if net_name == 'u2netp':                                     # load_models.py:583-772
        last = None                                              # pieces of the previous convolution's kept outputs
        stage_id, sides = 1, 0
        in_block, enc_stages, dec_stages = [], [], []            # selections saved per block / encoder stage / decoder stage

        def saved(sel):
            return None if sel is None else list(sel)

        def shifted(pieces, by):
            return [p._replace(offset=p.offset + by) for p in pieces]

        for conv in convs:
            if conv == 'outconv':
                break
            w = conv + '.weight'
            head = conv.split('.')[0]
            decode = head[-1] == 'd'
            block = None if head[:4] == 'side' else conv.split('.')[1]
            C, k, c_in = ori_shapes[w][0], new_shapes[w][0], ori_shapes[w][1]
            if decode and head[-2] != str(stage_id):             # :618-625 a new decoder / encoder stage begins
                stage_id -= 1
                dec_stages.append(saved(last))
                in_block = []
            elif not decode and head[-1] != str(stage_id):
                stage_id += 1
                enc_stages.append(saved(last))
                in_block = []
            if block is None:                                    # :745-769 side convolutions
                sides += 1
                if C != k:
                    ops.append(Op('gather', w, 'net.side%d' % sides, None if last is None else tuple(last)))
                elif last is not None:
                    ops.append(Op('gather', w, None, tuple(last)))
                else:
                    ops.append(Op('copy', w, None, None))
                last = dec_stages[5 - sides]
                continue
            stem = 'net.stage%d%s%s.relu_s1' % (stage_id, 'd.' if decode else '.', block)
            tag = block[-2:]                                     # 'in', 'v1' ... 'v7', '6d' ... '1d'
            first_of_decoder = decode and tag == 'in'
            if first_of_decoder or tag[1] == 'd':                # two concatenated inputs: the running one, then a saved one
                if first_of_decoder:
                    if C == k and last is not None:
                        raise NotImplementedError('%s: the reference indexes save_select_index[int("i")] here and stops' % conv)
                    other = enc_stages[stage_id - 1]
                else:
                    other = in_block[int(tag[0])]
                if C != k:                                       # :630-645, :706-720
                    ops.append(Op('gather', w, stem, tuple(last) + tuple(shifted(other, c_in // 2))))
                    last = [Piece(stem, 0, 0)]
                elif last is not None:                           # :722-736 (argsort with k == C keeps every channel)
                    ops.append(Op('gather', w, None, tuple(last) + tuple(shifted(other, c_in // 2))))
                    last = [Piece(stem, 0, 0)]
                else:
                    ops.append(Op('copy', w, None, None))
                    last = None
                if first_of_decoder:
                    in_block.append(saved(last))
            else:                                                # :667-704 one input
                if C != k:
                    ops.append(Op('gather', w, stem, None if last is None else tuple(last)))
                    last = [Piece(stem, 0, 0)]
                elif last is not None:
                    ops.append(Op('gather', w, None, tuple(last)))
                    last = [Piece(stem, 0, 0)]
                else:
                    ops.append(Op('copy', w, None, None))
                    last = None
                in_block.append(saved(last))
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


def _resolve(pieces, kept):
    """The input-channel ids a tuple of pieces stands for (None = all channels)."""
    if pieces is None:
        return None
    parts = []
    for p in pieces:
        if p.stem is None:
            ids = torch.arange(p.n, dtype=torch.int64)
        else:
            ids = torch.as_tensor(kept[p.stem], dtype=torch.int64)
        parts.append(ids + p.offset if p.offset else ids)
    if not parts:
        return torch.zeros(0, dtype=torch.int64)
    dev = next((t.device for t in parts if t.is_cuda), parts[0].device)
    return torch.cat([t.to(dev) for t in parts])


def apply_plan(plan, ori_state, new_state, kept, gather=gather_weight):
    """Run the ops.  `kept`: {score-file stem: int64 ids}.  Tensors not named by any op keep the pruned model's values."""
    for op in plan:
        src = ori_state[op.name]
        if op.kind == 'copy':
            new_state[op.name] = src.clone()
            continue
        g = gather(src, kept[op.out] if op.out is not None else None, _resolve(op.inp, kept))
        dst = new_state[op.name]
        if tuple(g.shape) == tuple(dst.shape):
            new_state[op.name] = g
        else:                                                    # the remembered selection is shorter than the input dimension
            dst = dst.clone()
            dst[:g.shape[0], :g.shape[1]] = g
            new_state[op.name] = dst
    return new_state


def transfer_weights(net_name, pruned_model, ori_state, kept, check=True, origin_rates=None):
    """Fill `pruned_model` (on a CUDA device) from the unpruned `ori_state` with the kept channels of every pruned
    layer; `kept` as returned by `topk.kept_channels` ([(Selection, ids)]) or a {stem: ids} mapping."""
    device = next(pruned_model.parameters()).device
    if device.type != 'cuda':
        raise RuntimeError('transfer_weights needs the pruned model on a CUDA device (got %s); there is no CPU fallback' % device)
    if not isinstance(kept, dict):
        kept = {sel.stem: ids for sel, ids in kept}
    kept = {k: torch.as_tensor(v, dtype=torch.int64).to(device) for k, v in kept.items()}
    ori = {k: v.to(device) for k, v in ori_state.items()}
    plan = transfer_plan(net_name, pruned_model, {k: tuple(v.shape) for k, v in ori.items()}, origin_rates)
    state = apply_plan(plan, ori, dict(pruned_model.state_dict()), kept)
    if check:
        _lib.check(_lib.load().dctp_check(_lib.current_stream()))
    pruned_model.load_state_dict(state)
    return plan
