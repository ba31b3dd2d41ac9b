#!/usr/bin/env python
"""Break the cfg-2 hot path into its parts on one B200 (CUDA-event timing, 20 reps each, L2 flushed
by a 256 MB memset between reps): which part of the fused forward / backward costs what."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recbox_b200 import ops  # noqa: E402

B, F, Fn, D, V = 65536, 26, 13, 16, 38462
dev = torch.device("cuda")
R = F * V
g = torch.Generator().manual_seed(0)
table = (torch.randn(R, D, generator=g) * 0.01).to(dev)
table_lr = (torch.randn(R, generator=g) * 0.01).to(dev)
dense_w = (torch.randn(Fn, D, generator=g) * 0.1).to(dev)
dense_w_lr = (torch.randn(Fn, generator=g) * 0.1).to(dev)
bias = torch.zeros(1, device=dev)
rng = np.random.default_rng(0)
rows = torch.from_numpy((rng.integers(1, V, size=(B, F)) + np.arange(F) * V).astype(np.int32)).to(dev)
dx = torch.rand(B, Fn, generator=g).to(dev)
cat_pos, num_pos = list(range(Fn, Fn + F)), list(range(Fn))
pad = [f * V for f in range(F)]
dE = (torch.randn(B, F + Fn, D, generator=g) * 1e-3).to(dev)
d1 = (torch.randn(B, generator=g) * 1e-3).to(dev)
gbuf = torch.zeros(R * D + R + Fn * D + Fn + 4, device=dev)
g_table = gbuf[:R * D].view(R, D)
g_lr = gbuf[R * D:R * D + R]
g_w = gbuf[R * D + R:R * D + R + Fn * D].view(Fn, D)
g_w1 = gbuf[R * D + R + Fn * D:R * D + R + Fn * D + Fn]
g_b = gbuf[-1:]
flush = torch.empty(64 * 1024 * 1024, device=dev)
big_a = torch.empty(163 * 1024 * 1024 // 4, device=dev)
big_b = torch.empty_like(big_a)
E, S, fm, lr = ops.embed_fm_fwd(table, table_lr, rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias)


def timeit(fn, reps=20, do_flush=True):
    ts = []
    for _ in range(reps + 3):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


cases = {
    "fwd full (E,S,fm,lr)": lambda: ops.embed_fm_fwd(table, table_lr, rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias),
    "fwd E only": lambda: ops.embed_fm_fwd(table, None, rows, cat_pos, dx, dense_w, None, num_pos, None, want_S=False, want_fm=False, want_lr=False),
    "fwd gather only (fm, no E)": lambda: ops.embed_fm_fwd(table, None, rows, cat_pos, dx, dense_w, None, num_pos, None, want_E=False, want_S=False, want_lr=False),
    "fwd lr only": lambda: ops.embed_fm_fwd(None, table_lr, rows, cat_pos, dx, None, dense_w_lr, num_pos, bias, want_E=False, want_S=False, want_fm=False),
    "fwd cat only (no numeric)": lambda: ops.embed_fm_fwd(table, table_lr, rows, list(range(F)), None, None, None, [], bias),
    "copy 163MB (torch)": lambda: big_b.copy_(big_a),
    "memset 163MB (torch)": lambda: big_a.zero_(),
    "zero grads 68MB": lambda: gbuf.zero_(),
    "bwd full": lambda: ops.embed_fm_bwd(table, rows, cat_pos, pad, dx, dense_w, num_pos, E, S, dE, d1, d1, g_table, g_lr, g_w, g_w1, g_b, D, R),
    "bwd dE only (no fm/lr)": lambda: ops.embed_fm_bwd(table, rows, cat_pos, pad, dx, dense_w, num_pos, None, None, dE, None, None, g_table, None, g_w, None, None, D, R),
    "bwd cat part only (Fn=0 view)": lambda: ops.embed_fm_bwd(table, rows, cat_pos, pad, None, None, [], E, S, dE, d1, d1, g_table, g_lr, None, None, None, D, R) if False else None,
    "gather_rows 1.7M x 64B": lambda: ops.gather_rows(table, rows.view(-1)),
    "scatter_add_rows 1.7M x 64B": lambda: ops.scatter_add_rows(dE.view(-1, D)[:B * F], rows.view(-1), None, g_table),
}
out = {}
for name, fn in cases.items():
    if fn() is None and "Fn=0" in name:
        continue
    out[name] = timeit(fn)
    out[name + " [warm L2]"] = timeit(fn, do_flush=False)
    print("%-36s %8.1f us   warm-L2 %8.1f us" % (name, out[name], out[name + " [warm L2]"]))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("RBX_TAG", "micro")
with open(os.path.join(ROOT, "gpurun_out", tag + ".json"), "w") as f:
    json.dump(out, f, indent=1)
