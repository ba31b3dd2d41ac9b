"""One forward and one backward launch of interact[inner_product] at the BASELINE shape (for ncu captures):
ncu --set full --clock-control none -k regex:k_ip_ -o gpurun_out/<tag> python tools/interact_one.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops  # noqa: E402

B, F, D = 65536, 39, 16
E = torch.randn(B, F, D, device="cuda")
dout = torch.randn(B, F * (F - 1) // 2, device="cuda")
out = ops.interact_fwd(E, 2)
dE = ops.interact_bwd(E, dout, 2)
torch.cuda.synchronize()
print(float(out[0, 0]), float(dE[0, 0, 0]))
