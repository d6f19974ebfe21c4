"""TEST INFRASTRUCTURE ONLY - CPU oracle, never imported by the product path.

Restatement ("port") of the reference's importance-generation hot path, op for
op and in the same order, on host cores.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker or the timed CPU baseline.

What follows what (all in /root/reference/utils/common.py):
  torch2dct                 :230-239   cv2.dct of one slice, odd-row front pad
  cnt_score                 :249-255   per-slice sum(d*d).item()
  ScoreState.update         :275-277   fp32 running mean over images
  hook_output               :262-277   get_feature_hook
  hook_densenet             :280-293   get_feature_hook_densenet (last 12 channels)
  hook_u2net_input          :296-309   get_feature_hook_u2net_input (input[0])
  inference                 :312-332   eval + no_grad forward over `limit` batches
  imp_score_port            :367-980   one hook session per site, re-running the net
  select_index_*            utils/load_models.py:39-41 (and the 20 sibling sites)

PARITY: tier-1 = this file with oracle/torch_dct_port.py standing in for the
absent `torch_dct` (see that file: "parity unpinned" for the DCT arithmetic);
tier-2 = scipy float64 `dctn(norm='ortho')`; tier-3 = Parseval sum of squares.
tests/golden/make_golden.py ran the *real* reference (imported unmodified under
stubs for its absent optional imports) on seeded inputs and committed its
outputs; tests/test_oracle.py checks this port against those vectors.
"""
import numpy as np
import torch

from . import torch_dct_port as dct


# --------------------------------------------------------------------- per slice
def torch2dct(feature_map):
    import cv2
    t = np.float32(feature_map.cpu().numpy())
    if t.shape[0] % 2 != 0:
        t = np.pad(t, (1, 0), 'constant')      # one zero row on top AND one zero column left
    return torch.from_numpy(cv2.dct(t))


def cnt_score(dct_list):
    for idx, d in enumerate(dct_list):
        dct_list[idx] = torch.sum(d.mul(d)).item()
    return torch.tensor(dct_list)


class ScoreState:
    """The reference's module globals `feature_result` / `total` as an object."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.feature_result = torch.tensor(0.)
        self.total = torch.tensor(0.)

    def update(self, c, a):
        self.feature_result = self.feature_result * self.total + c
        self.total = self.total + a
        self.feature_result = self.feature_result / self.total


def hook_output(state):
    def hook(module, inputs, output):
        a, b = output.shape[0], output.shape[1]
        c = [dct.dct_2d(output[i, j, :, :], norm='ortho') for i in range(a) for j in range(b)]
        c = cnt_score(c).view(a, -1).sum(0)
        state.update(c, a)
    return hook


def hook_densenet(state):
    def hook(module, inputs, output):
        a, b = output.shape[0], output.shape[1]
        c = [torch2dct(output[i, j, :, :]) for i in range(a) for j in range(b - 12, b)]
        c = cnt_score(c).view(a, -1).float().sum(0)
        state.update(c, a)
    return hook


def hook_u2net_input(state):
    def hook(module, inputs, output):
        x = inputs[0]
        a, b = x.shape[0], x.shape[1]
        c = [torch2dct(x[i, j, :, :]) for i in range(a) for j in range(b)]
        c = cnt_score(c).view(a, -1).sum(0)
        state.update(c, a)
    return hook


HOOKS = {'O': hook_output, 'D': hook_densenet, 'I': hook_u2net_input}


# ------------------------------------------------------------------- per session
def inference(net, batches, limit):
    net.eval()
    for batch_idx, data in enumerate(batches):
        if batch_idx >= limit:
            break
        with torch.no_grad():
            net(data)


def resolve(net, path):
    """'features.6' / 'layer2.3.relu1' -> module (digits index containers)."""
    mod = net
    for part in path.split('.'):
        mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
    return mod


def imp_score_port(net, sessions, make_batches, limit):
    """sessions: [(module_path, variant, [(file_stem, lo, hi), ...])] with lo/hi = None
    for the whole vector.  Returns {file_stem: float32 array}.  Like the reference, each
    session registers one hook and re-runs the net over a fresh batch iterator."""
    out = {}
    state = ScoreState()
    for path, variant, files in sessions:
        handle = resolve(net, path).register_forward_hook(HOOKS[variant](state))
        inference(net, make_batches(), limit)
        handle.remove()
        vec = state.feature_result.numpy()
        for stem, lo, hi in files:
            out[stem] = np.array(vec if lo is None else vec[lo:hi], dtype=np.float32, copy=True)
        state.reset()
    return out


# ------------------------------------------------------- vectorised cross-checks
def energy_scipy64(x):
    """Tier-2: per-(image, channel) energy of the float64 orthonormal 2-D DCT-II."""
    from scipy.fft import dctn
    z = dctn(np.asarray(x, dtype=np.float64), type=2, norm='ortho', axes=(-2, -1))
    return (z * z).sum(axis=(-2, -1))


def energy_parseval64(x):
    """Tier-3: the orthonormal transform preserves energy, so this is the same number."""
    x = np.asarray(x, dtype=np.float64)
    return (x * x).sum(axis=(-2, -1))


def score_scipy64(x, c_begin=0, c_count=None):
    """Mean over images of the per-channel energy, float64 (what a-1..a-7 compute)."""
    x = np.asarray(x)
    c_count = x.shape[1] - c_begin if c_count is None else c_count
    return energy_scipy64(x[:, c_begin:c_begin + c_count]).mean(axis=0)


# ------------------------------------------------------------------------- top-k
def select_index_reference(imp, k):
    """utils/load_models.py:40-41: default (unstable) argsort, keep the k largest, ascending ids."""
    imp = np.asarray(imp)
    sel = np.argsort(imp)[len(imp) - k:]
    sel.sort()
    return sel


def select_index_stable(imp, k):
    """The documented tie rule: among scores equal to the cut value keep the
    highest channel ids == np.argsort(kind='stable')[C-k:]."""
    imp = np.asarray(imp)
    sel = np.argsort(imp, kind='stable')[len(imp) - k:]
    sel.sort()
    return sel


def topk_equivalent(imp, got, k):
    """True when `got` is a legal answer to the reference's selection: every score above the
    cut is kept, everything else kept equals the cut, and the count is k."""
    imp = np.asarray(imp)
    got = np.asarray(got)
    if len(got) != k or len(np.unique(got)) != k:
        return False
    if k == 0:
        return True
    cut = np.sort(imp)[len(imp) - k]
    must = np.nonzero(imp > cut)[0]
    return bool(np.isin(must, got).all() and (imp[got] >= cut).all()
                and (np.diff(got) > 0).all())
