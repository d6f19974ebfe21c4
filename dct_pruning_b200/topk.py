"""Kept-channel selection on the device.

Replaces the seven copies of ``select_index = np.argsort(imp)[C-k:]; select_index.sort()`` in
/root/reference/utils/load_models.py (:39-41, :102-104, :265-267, :313-315, :352-354, :407-409,
:469-471, :521-523, :629-631 ... :746-748) with one segmented kernel launch per net.

Tie rule: among channels equal to the cut value the highest channel ids are kept, i.e. exactly
``np.argsort(imp, kind='stable')[C-k:]``.  The reference's default argsort is unstable, so for
exact ties its own answer is implementation-defined; everything strictly above the cut is
always identical (tests/test_topk_*.py).
"""
import numpy as np
import torch

from . import _lib
from .compress import get_compress_rate, selection_plan


def topk_segmented(scores, offsets, ks):
    """scores: float32 CUDA tensor (all segments back to back); offsets: n_seg+1 ints; ks: n_seg ints.
    Returns a list of int64 CUDA tensors (ascending kept ids, relative to each segment)."""
    if not scores.is_cuda:
        raise RuntimeError('top-k selection runs on CUDA only; there is no CPU fallback')
    lib = _lib.load()
    n_seg = len(ks)
    if n_seg == 0:
        return []
    offsets = [int(o) for o in offsets]
    ks = [max(0, min(int(k), offsets[i + 1] - offsets[i])) for i, k in enumerate(ks)]
    out_off = np.concatenate([[0], np.cumsum(ks)]).astype(np.int32)
    dev = scores.device
    meta = torch.tensor(np.concatenate([np.asarray(offsets, np.int32), np.asarray(ks, np.int32), out_off]),
                        dtype=torch.int32).to(dev, non_blocking=True)
    seg_off, seg_k, o_off = meta[:n_seg + 1], meta[n_seg + 1:2 * n_seg + 1], meta[2 * n_seg + 1:]
    out = torch.empty(max(int(out_off[-1]), 1), dtype=torch.int64, device=dev)
    scores = scores.contiguous()
    _lib.check(lib.dctp_topk_segmented(_lib.ptr(scores), _lib.ptr(seg_off), _lib.ptr(seg_k), n_seg,
                                       _lib.ptr(out), _lib.ptr(o_off), _lib.current_stream()))
    return [out[int(out_off[i]):int(out_off[i + 1])] for i in range(n_seg)]


def select_index(imp, k, device='cuda'):
    """Single-vector form, same contract as the reference's two lines: ascending kept channel ids."""
    t = torch.as_tensor(np.asarray(imp, dtype=np.float32)).to(device)
    return topk_segmented(t, [0, t.numel()], [k])[0].cpu().numpy()


def kept_channels_device(net_name, compress_rate, device_scores, segments, origin_rates=None):
    """Same selections as `kept_channels`, from the flat score vector still on the device (ScoreSession.finalize_device) and its
    [(file stem, offset, length)] segments (ScoreSession.file_segments): the scores go from the finalise kernel to the top-k
    kernel without a host round trip.  Returns [(Selection, int64 numpy array)]."""
    rates = get_compress_rate(compress_rate) if isinstance(compress_rate, str) else list(compress_rate)
    plan = selection_plan(net_name, rates, origin_rates)
    if not plan:
        return []
    where = {stem: (off, n) for stem, off, n in segments}
    pieces, offsets, ks = [], [0], []
    for sel in plan:
        off, n = where[sel.stem]
        if n != sel.C:
            raise ValueError('%s: score vector has %d entries, layer has %d channels' % (sel.stem, n, sel.C))
        pieces.append(device_scores[off:off + n])
        offsets.append(offsets[-1] + n)
        ks.append(sel.k)
    kept = topk_segmented(torch.cat(pieces), offsets, ks)
    host = torch.cat(kept).cpu().numpy() if kept else np.zeros(0, np.int64)
    out, at = [], 0
    for sel in plan:
        out.append((sel, host[at:at + sel.k].copy()))
        at += sel.k
    return out


def kept_channels(net_name, compress_rate, scores, device='cuda', origin_rates=None):
    """For every selection the reference's loader for `net_name` performs under `compress_rate`
    (string or list of floats), the kept channel ids.  `scores` maps file stem -> vector (numpy or
    tensor) or is a directory holding the .npy files.  `origin_rates`: the rates of the scored net when it is itself a
    pruned one (iterative pruning).  Returns [(Selection, int64 numpy array)]."""
    rates = get_compress_rate(compress_rate) if isinstance(compress_rate, str) else list(compress_rate)
    plan = selection_plan(net_name, rates, origin_rates)
    if isinstance(scores, str):
        import os
        scores = {s.stem: np.load(os.path.join(scores, s.stem + '.npy')) for s in plan}
    vecs, offsets, ks = [], [0], []
    for sel in plan:
        v = torch.as_tensor(np.asarray(scores[sel.stem], dtype=np.float32))
        if v.numel() != sel.C:
            raise ValueError('%s: score vector has %d entries, layer has %d channels' % (sel.stem, v.numel(), sel.C))
        vecs.append(v)
        offsets.append(offsets[-1] + sel.C)
        ks.append(sel.k)
    if not plan:
        return []
    flat = torch.cat(vecs).to(device)
    kept = topk_segmented(flat, offsets, ks)
    host = torch.cat(kept).cpu().numpy() if kept else np.zeros(0, np.int64)
    out, at = [], 0
    for sel in plan:
        out.append((sel, host[at:at + sel.k].copy()))
        at += sel.k
    return out
