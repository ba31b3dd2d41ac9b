"""ncu --set full targets: inner_product interaction fwd/bwd at configs[1] shape, the top-k filter pass, unique rank pass."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
E = torch.randn(65536, 39, 16, generator=g).to(dev)
do = torch.randn(65536, 741, generator=g).to(dev)
for _ in range(2):
    ops.interact_fwd(E, 2)
    ops.interact_bwd(E, do, 2)
q = torch.randn(1024, 64, generator=g).to(dev)
items = torch.randn(1_000_000, 64, generator=g).to(dev)
ops.topk_ip(q, items, 100)
torch.cuda.synchronize()
