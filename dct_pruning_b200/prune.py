"""Score -> select -> rebuild -> fill, all on the device: the non-resume branch of the reference's `load_model`
(/root/reference/utils/load_models.py:803-838: build the unpruned net, `imp_score` it, then `load_<net>_model` copies
the kept filters into the pruned net), and the body of `prune_dynamic.py:150-154`'s loop, which repeats it on nets
that are already pruned.  When the scores are generated here they go from the finalise kernel to the top-k kernel on the
device (ScoreSession.finalize_device -> topk.kept_channels_device); the host copy that is returned / written as .npy files is
taken afterwards and nothing on the selection path waits for it.  Scores passed in (a directory of files, a mapping) are
uploaded once.
"""
import types

import torch

from .compress import get_compress_rate
from .generate import score_session, write_score_files
from .sites import score_dir
from .topk import kept_channels, kept_channels_device
from .transfer import transfer_weights
from .zoo import get_network


def pruned_model(net_name, compress_rate, origin_model, limit=5, batch_size=128, loader=None, scores=None,
                 input_side=None, seed=None, out_root=None, origin_rates=None):
    """Returns (pruned net on origin_model's device, {file stem: score vector}, [(Selection, kept ids)]).

    compress_rate: the reference's string form (`'[0.]+[0.18]*29'`) or a list of floats.  `scores` skips the scoring
    pass (e.g. files of an earlier run: a directory or a {stem: vector} mapping).  `seed` seeds the pruned net's own
    initialisation - the tensors the loaders do not fill (biases, most BatchNorms, the classifier of VGG) keep it, as in
    the reference.  `out_root` additionally writes the reference's importance_score/<net>_limit<N>/*.npy files.
    `origin_rates`: the rates `origin_model` was built with when it is itself a pruned net - one round of
    prune_dynamic.py's loop (score the fine-tuned pruned net, prune it further)."""
    device = next(origin_model.parameters()).device
    if device.type != 'cuda':
        raise RuntimeError('pruned_model needs the unpruned net on a CUDA device (got %s); there is no CPU fallback' % device)
    rates = get_compress_rate(compress_rate) if isinstance(compress_rate, str) else list(compress_rate)
    if scores is None:
        import torch.distributed as dist
        args = types.SimpleNamespace(net=net_name, limit=limit, batch_size=batch_size, input_side=input_side)
        session = score_session(origin_model, args, loader=loader)
        device_scores = session.finalize_device()
        kept = kept_channels_device(net_name, rates, device_scores, session.file_segments(), origin_rates=origin_rates)
        scores = session.split_files(device_scores.cpu().numpy())
        rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
        if out_root is not None and rank0:
            write_score_files(scores, score_dir(net_name, limit, out_root))
    else:
        kept = kept_channels(net_name, rates, scores, device=device, origin_rates=origin_rates)
    if seed is not None:
        torch.manual_seed(seed)
    net = get_network(net_name, rates).to(device).eval()
    transfer_weights(net_name, net, origin_model.state_dict(), kept, origin_rates=origin_rates)
    return net, scores, kept
