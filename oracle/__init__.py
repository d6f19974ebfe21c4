"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's importance-generation path, used as the parity
checker by tests/, by __graft_entry__.smoke() and as bench.py's timed CPU baseline.
Nothing under dct_pruning_b200/ imports it: the product path is CUDA-only and fails
loudly when its extension is missing.
"""
