"""TEST INFRASTRUCTURE ONLY - CPU oracle, never imported by the product path.

Restatement of the copy loops the reference runs after top-k to fill a pruned model
(/root/reference/utils/load_models.py:43-51, :106-114, :482-500, :526-542, :633-639 ...):

    for index_i, i in enumerate(select_index):
        for index_j, j in enumerate(last_select_index):
            state_dict[name][index_i][index_j] = oristate_dict[name][i][j]

`copy_loops` is that double loop literally (small cases); `gather_numpy` is the same map as one
fancy-indexing expression.  Pinned against the reference itself: tests/golden/transfer_<net>.json
holds digests of every tensor the unmodified reference loaders wrote (make_golden.py transfer).
"""
import numpy as np


def copy_loops(w, select_index=None, last_select_index=None):
    """w: ndarray [C_out, C_in, ...]; returns the [k_out, k_in, ...] block the loops write."""
    sel_o = list(range(w.shape[0])) if select_index is None else [int(i) for i in select_index]
    if w.ndim == 1:
        out = np.empty((len(sel_o),), w.dtype)
        for index_i, i in enumerate(sel_o):
            out[index_i] = w[i]
        return out
    sel_i = list(range(w.shape[1])) if last_select_index is None else [int(j) for j in last_select_index]
    out = np.empty((len(sel_o), len(sel_i)) + w.shape[2:], w.dtype)
    for index_i, i in enumerate(sel_o):
        for index_j, j in enumerate(sel_i):
            out[index_i][index_j] = w[i][j]
    return out


def gather_numpy(w, select_index=None, last_select_index=None):
    out = w if select_index is None else w[np.asarray(select_index, dtype=np.int64)]
    if last_select_index is not None:
        out = out[:, np.asarray(last_select_index, dtype=np.int64)]
    return np.ascontiguousarray(out)
