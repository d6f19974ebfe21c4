"""dct_pruning_b200: B200-native importance-score path of DCT filter pruning.

Hook capture -> per-channel 2-D DCT-II energy -> accumulation over --limit batches -> top-k
kept channels, as hand-written CUDA (sm_100a) behind the C ABI in include/dctp.h.  Importing
the package does not load the library; the first scoring call does, and fails loudly if the
in-tree build is missing.  There is no CPU fallback.
"""
from .compress import get_compress_rate, selection_plan          # noqa: F401
from .sites import hook_sites, score_dir                         # noqa: F401

__all__ = ['get_compress_rate', 'selection_plan', 'hook_sites', 'score_dir']
