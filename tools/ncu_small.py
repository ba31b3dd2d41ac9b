"""Tiny driver for ncu launch lists of the secondary kernels (unique / touched-rows optimizer / sampler)."""
import numpy as np
import torch
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops
dev = "cuda"
rng = np.random.default_rng(0)
B, F, V = 65536, 26, 38462
R = F * V
rows = torch.from_numpy((rng.integers(1, V, size=(B, F)) + np.arange(F) * V).astype(np.int32)).to(dev)
for _ in range(3):
    ops.unique_ids(rows, R, sync=False)                                          # first + inverse
    ops.unique_ids(rows, R, want_first=False, want_inverse=False, sync=False)    # touched-rows mode
ids = torch.from_numpy(rng.integers(0, 10_000_000, size=(8192, 11)).astype(np.int64)).to(dev)
for _ in range(3):
    ops.unique_ids(ids, 10_000_001, sync=False)
pos = torch.randint(0, 10_000_000, (1_000_000,), device=dev)
for _ in range(3):
    ops.sample_negatives(1_000_000, 10, 10_000_000, 1, pos=pos)
torch.cuda.synchronize()
