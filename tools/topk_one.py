"""One rbx_topk_ip search (1 024 users x 10 M items x 64, k = 100) for ncu launch lists: python tools/topk_one.py [U]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops  # noqa: E402

U = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
items = torch.randn(10_000_000, 64, device="cuda")
q = torch.randn(U, 64, device="cuda")
ops.topk_ip(q, items, 100)
torch.cuda.synchronize()
