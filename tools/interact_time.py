"""interact[inner_product] forward / backward at the BASELINE shape (B = 65 536, F = 39, D = 16) on both engines:
RBX_IP_ENGINE=1 (default) the warp-level 3xTF32 mma kernels, =0 the SIMT register-tile kernels; time against the HBM roofline
(L2 flushed between reps) and the error of each against a float64 product on the first 4 096 samples."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops  # noqa: E402

B, F, D = 65536, int(os.environ.get("IP_F", "39")), 16
torch.manual_seed(0)
E = torch.randn(B, F, D, device="cuda")
P = F * (F - 1) // 2
dout = torch.randn(B, P, device="cuda")
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
n = 4096
iu = torch.triu_indices(F, F, 1, device="cuda")
E64 = E[:n].double()
ref_f = torch.bmm(E64, E64.transpose(1, 2))[:, iu[0], iu[1]]
G = torch.zeros(n, F, F, dtype=torch.float64, device="cuda")
G[:, iu[0], iu[1]] = dout[:n].double()
ref_b = torch.bmm(G + G.transpose(1, 2), E64)


def t(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


bf, bb = B * (F * D * 4 + P * 4), B * (2 * F * D * 4 + P * 4)
for eng in ("1", "0"):
    os.environ["RBX_IP_ENGINE"] = eng
    ef = float((ops.interact_fwd(E[:n].contiguous(), 2).double() - ref_f).abs().max() / ref_f.abs().max())
    eb = float((ops.interact_bwd(E[:n].contiguous(), dout[:n].contiguous(), 2).double() - ref_b).abs().max() / ref_b.abs().max())
    fw = t(lambda: ops.interact_fwd(E, 2))
    bw = t(lambda: ops.interact_bwd(E, dout, 2))
    print("inner_product F=%d engine=%s (%s): fwd %.1f us %.0f GB/s (%.2f)  bwd %.1f us %.0f GB/s (%.2f)  max err / max |ref| vs float64: fwd %.2e bwd %.2e"
          % (F, eng, "mma 3xTF32" if eng == "1" else "SIMT", fw, bf / fw / 1e3, bf / fw / 1e3 / 6425.6, bw, bb / bw / 1e3,
             bb / bw / 1e3 / 6425.6, ef, eb))
