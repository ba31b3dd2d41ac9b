import sys, torch
sys.path.insert(0, "/root/repo")
from recbox_b200 import ops
B, F, D = 65536, 39, 16
E = torch.randn(B, F, D, device="cuda")
P = F * (F - 1) // 2
dout = torch.randn(B, P, device="cuda")
flush = torch.zeros(64 * 1024 * 1024, device="cuda")
def t(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[2:]); return ts[len(ts) // 2]
fw = t(lambda: ops.interact_fwd(E, 2))
bw = t(lambda: ops.interact_bwd(E, dout, 2))
bf, bb = B * (F * D * 4 + P * 4), B * (2 * F * D * 4 + P * 4)
print("inner_product fwd %.1f us %.0f GB/s (%.2f)  bwd %.1f us %.0f GB/s (%.2f)" % (fw, bf / fw / 1e3, bf / fw / 1e3 / 6425.6, bw, bb / bw / 1e3, bb / bw / 1e3 / 6425.6))
