#!/usr/bin/env python
"""Every kernel of the C ABI against the HBM roofline on one B200, at the sizes of BASELINE.json's configs
(SURVEY.md section 8d): achieved = ALGORITHMIC bytes / CUDA-event time (median of 15 reps, a 256 MB read flushes the
126 MB L2 between reps), peak = MEASURED_PEAKS.json hbm_gbs.

  python tools/kernel_rooflines.py [tag]      -> gpurun_out/<tag>_kernels.json + a markdown table on stdout

The fused forward / backward pair of configs[1] is bench.py's job; here it is swept over D = 16..128 and the
other entry points (a1 split, a9 pooled lookup, a10 row-dot, a11 SASRec gathers, a6 interaction modes, a12/f1
optimizers, a14 unique, f3 top-k retrieval) get their own line.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from recbox_b200 import ops  # noqa: E402

dev = torch.device("cuda")
tag = sys.argv[1] if len(sys.argv) > 1 else "kern"
peak = 6650.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.zeros(64 * 1024 * 1024, device=dev)
results = []


def timeit(fn, reps=15):
    """L2 flush = READING 256 MB (torch.sum): it evicts the previous rep's lines and leaves the L2 full of CLEAN lines.
    (Flushing with a memset leaves ~126 MB of dirty lines whose write-back is then billed to the kernel under test:
    that cost small kernels 10-20 us in the r1u / r1v tables.)"""
    ts = []
    for _ in range(reps + 3):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


def case(row, name, shape, alg_bytes, fn):
    try:
        us = timeit(fn)
    except Exception as e:            # a failing case must not hide the others
        print("%-44s FAILED: %s" % (name, e))
        results.append({"row": row, "kernel": name, "shape": shape, "error": str(e)})
        return
    gbs = alg_bytes / us / 1e3
    results.append({"row": row, "kernel": name, "shape": shape, "alg_bytes": int(alg_bytes), "us": us, "gbs": gbs,
                    "frac": gbs / peak})
    print("| %s | `%s` | %s | %.1f MB | %.1f | %.0f | %.2f |" % (row, name, shape, alg_bytes / 1e6, us, gbs, gbs / peak))


print("| row | kernel | shape | alg. bytes | us | GB/s | frac of %.0f |" % peak)
print("|---|---|---|---|---|---|---|")
g = torch.Generator().manual_seed(0)
rng = np.random.default_rng(0)

# ---- a1: batch split -------------------------------------------------------------------------
B, F, Fn, V = 65536, 26, 13, 38462
M = torch.cat([torch.rand(B, Fn, dtype=torch.float64), torch.from_numpy(rng.integers(1, V, size=(B, F))).double(),
               torch.zeros(B, 1, dtype=torch.float64)], 1).to(dev)
ck, cs, fo = [2] * Fn + [1] * F + [3], list(range(Fn)) + list(range(F)) + [0], [f * V for f in range(F)]
case("a1", "rbx_split_batch_f64", "B=65536, 40 cols", B * (40 * 8 + (F + Fn + 1) * 4), lambda: ops.split_batch(M, ck, cs, fo, F, Fn))
del M

# ---- a2-a8: fused pair over D (configs[1] shape, uniform ids) -------------------------------
for D in (16, 32, 64, 128):
    R = F * V
    table = (torch.randn(R, D, generator=g) * 0.01).to(dev)
    table_lr = (torch.randn(R, generator=g) * 0.01).to(dev)
    dw, dw1, bias = (torch.randn(Fn, D, generator=g) * 0.1).to(dev), (torch.randn(Fn, generator=g) * 0.1).to(dev), torch.zeros(1, device=dev)
    rows = torch.from_numpy((rng.integers(1, V, size=(B, F)) + np.arange(F) * V).astype(np.int32)).to(dev)
    dx = torch.rand(B, Fn, generator=g).to(dev)
    cp, npos = list(range(Fn, Fn + F)), list(range(Fn))
    dE = (torch.randn(B, F + Fn, D, generator=g) * 1e-3).to(dev)
    d1 = (torch.randn(B, generator=g) * 1e-3).to(dev)
    gt, gl = torch.zeros(R, D, device=dev), torch.zeros(R, device=dev)
    gw, gw1, gb = torch.zeros(Fn, D, device=dev), torch.zeros(Fn, device=dev), torch.zeros(1, device=dev)
    E, S, fm, lr = ops.embed_fm_fwd(table, table_lr, rows, cp, dx, dw, dw1, npos, bias)
    bf = B * (F * (4 + 4 * D) + (F + Fn) * 4 * D + 4 * F + 4 * Fn + 4)
    bb = B * (4 * F + (F + Fn) * 4 * D + F * 4 * D + F * 8 * D + 8 * F + 4)
    case("a2-a8", "rbx_embed_fm_fwd", "B=65536 F=26+13 D=%d" % D, bf,
         lambda: ops.embed_fm_fwd(table, table_lr, rows, cp, dx, dw, dw1, npos, bias))
    case("a5/a6 bwd", "rbx_embed_fm_bwd", "B=65536 F=26+13 D=%d" % D, bb,
         lambda: ops.embed_fm_bwd(table, rows, cp, fo, dx, dw, npos, E, S, dE, d1, d1, gt, gl, gw, gw1, gb, D, R))
    if D == 16:
        # a6: the four modes on a materialised E [B,39,16]
        Ft = F + Fn
        P = Ft * (Ft - 1) // 2
        for mode, nm, out_b in ((0, "product_sum", 4), (1, "bi_interaction", 4 * D), (2, "inner_product", 4 * P)):
            do = torch.randn(ops.interact_out_shape(B, Ft, D, mode), generator=g).to(dev)
            case("a6", "rbx_interact_fwd[%s]" % nm, "B=65536 F=39 D=16", B * (4 * Ft * D + out_b), lambda: ops.interact_fwd(E, mode))
            case("a6", "rbx_interact_bwd[%s]" % nm, "B=65536 F=39 D=16", B * (8 * Ft * D + out_b), lambda: ops.interact_bwd(E, do, mode))
        for order in (2, 5):
            dP = torch.randn(B, order, D, generator=g).to(dev)
            case("f4", "rbx_power_sums_fwd[order %d]" % order, "B=65536 F=39 D=16", B * (4 * Ft * D + 4 * order * D), lambda: ops.power_sums_fwd(E, order))
            case("f4", "rbx_power_sums_bwd[order %d]" % order, "B=65536 F=39 D=16", B * (8 * Ft * D + 4 * order * D), lambda: ops.power_sums_bwd(E, dP))
        Bs = 8192
        Es = E[:Bs].contiguous()
        do = torch.randn(Bs, P, D, generator=g).to(dev)
        case("a6", "rbx_interact_fwd[elementwise_product]", "B=8192 F=39 D=16", Bs * (4 * Ft * D + 4 * P * D), lambda: ops.interact_fwd(Es, 3))
        case("a6", "rbx_interact_bwd[elementwise_product]", "B=8192 F=39 D=16", Bs * (8 * Ft * D + 4 * P * D), lambda: ops.interact_bwd(Es, do, 3))
        del do, Es
        # a12 exact-dense mode and f1 touched-rows mode on the same table
        m, v = torch.zeros_like(table), torch.zeros_like(table)
        acc = torch.zeros(1, dtype=torch.float64, device=dev)
        coef = torch.ones(1, device=dev)
        gt.normal_()
        case("a12", "rbx_sqnorm", "1 000 012 x 16", R * D * 4, lambda: ops.sqnorm_(gt, acc))
        case("a12", "rbx_adam_dense", "1 000 012 x 16", R * D * 4 * 7, lambda: ops.adam_dense_(table, gt, m, v, 1, clip=coef))
        uniq, first, inv = ops.unique_ids(rows, R)
        U = uniq.numel()
        n_rows = torch.tensor([U], dtype=torch.int64, device=dev)
        case("f1", "rbx_sqnorm_rows", "%d touched rows x 16" % U, U * (4 + 4 * D), lambda: ops.sqnorm_rows_(gt, uniq, n_rows, acc))
        case("f1", "rbx_optim_rows[adam_rows]", "%d touched rows x 16" % U, U * (4 + 4 * D * 8),
             lambda: ops.optim_rows_(table, gt, m, v, uniq, n_rows, 1, kind="adam_rows", clip=coef))
        case("a14", "rbx_unique_ids_i32", "n=1.7M ids, vocab 1 000 012", rows.numel() * (4 + 4) + U * 12 + R // 4,
             lambda: ops.unique_ids(rows, R, sync=False))
        del m, v
    del table, table_lr, dE, E, S, gt, gl, rows

# ---- a9 / a10: configs[2] two-tower DSSM, 10M items, D=64 -----------------------------------
D, N_items = 64, 10_000_001
items = torch.empty(N_items, D, device=dev).normal_(0, 0.01)
g_items = torch.zeros_like(items)
for Bm, L in ((8192, 20), (65536, 20)):
    hist = torch.from_numpy(rng.integers(0, N_items - 1, size=(Bm, L)).astype(np.int32)).to(dev)
    go = torch.randn(Bm, D, generator=g).to(dev)
    out, cnt = ops.pooled_gather_fwd(items, hist, 1)
    case("a9", "rbx_pooled_gather_fwd[avg]", "B=%d L=20 D=64, 10M-row table" % Bm, Bm * (L * (4 + 4 * D) + 4 * D),
         lambda: ops.pooled_gather_fwd(items, hist, 1))
    case("a9", "rbx_pooled_gather_bwd[avg]", "B=%d L=20 D=64, 10M-row table" % Bm, Bm * (L * (4 + 8 * D) + 4 * D),
         lambda: ops.pooled_gather_bwd(go, hist, cnt, N_items - 1, g_items, 1))
    K = 11
    u = torch.randn(Bm, D, generator=g).to(dev)
    vv = torch.randn(Bm, K, D, generator=g).to(dev)
    dy = torch.randn(Bm, K, generator=g).to(dev)
    case("a10", "rbx_rowdot_fwd", "B=%d K=11 D=64" % Bm, Bm * (4 * D * (1 + K) + 4 * K), lambda: ops.rowdot_fwd(u, vv))
    case("a10", "rbx_rowdot_bwd", "B=%d K=11 D=64" % Bm, Bm * (8 * D * (1 + K) + 4 * K), lambda: ops.rowdot_bwd(u, vv, dy))
emb = torch.randn(8192, 200, D, generator=g).to(dev)
case("a9", "rbx_pool_fwd[avg]", "[8192,200,64] materialised", emb.numel() * 4 + 8192 * D * 4, lambda: ops.pool_fwd(emb, None, 1))
del emb
ids = torch.from_numpy(rng.integers(0, N_items - 1, size=(8192, 11)).astype(np.int64)).to(dev)
case("a14", "rbx_unique_ids_i64", "n=8192x11 ids, vocab 10M (collate_fn_unique)", ids.numel() * 16 + 2 * (N_items // 8),
     lambda: ops.unique_ids(ids, N_items, sync=False))
# f3: brute-force inner-product retrieval, 10M items (bound: fp32 FMA pipe, 2*U*N*D flop; nominal peak
# 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s -- computed, not measured)
FP32_PEAK = 148 * 128 * 2 * 1.965e9
for U_, k_ in ((1024, 100), (1024, 500), (128, 100)):
    q = torch.randn(U_, D, generator=g).to(dev)
    try:
        us = timeit(lambda: ops.topk_ip(q, items, k_), reps=3)
        fl = 2.0 * U_ * N_items * D
        results.append({"row": "f3", "kernel": "rbx_topk_ip", "shape": "U=%d x 10M items D=64 k=%d" % (U_, k_), "flop": fl,
                        "us": us, "tflops": fl / us / 1e6, "frac_fp32_peak": fl / (us * 1e-6) / FP32_PEAK,
                        "corpus_gbs": N_items * D * 4 / us / 1e3})
        print("| f3 | `rbx_topk_ip` | U=%d x 10M items D=64 k=%d | %.2f TFLOP | %.0f | %.1f TFLOP/s fp32 | %.2f of 74.4 (FMA pipe) |" % (
            U_, k_, fl / 1e12, us, fl / us / 1e6, fl / (us * 1e-6) / FP32_PEAK))
    except Exception as e:
        print("rbx_topk_ip FAILED:", e)
pos = torch.randint(0, N_items, (1_000_000,), device=dev)
case("f2", "rbx_sample_negatives", "1M queries x 10 negs over 10M items", 1_000_000 * 11 * 8 + 1_000_000 * 8,
     lambda: ops.sample_negatives(1_000_000, 10, N_items, 1, pos=pos))
del items, g_items

# ---- a11: configs[4] SASRec, 1M items, L=200, D=64, B=1024: three shared-table lookups -------
N_items = 1_000_001
items = torch.empty(N_items, D, device=dev).normal_(0, 0.01)
g_items = torch.zeros_like(items)
seq = torch.from_numpy(rng.integers(1, N_items, size=(3, 1024, 200)).astype(np.int32)).to(dev)
gr = torch.randn(3 * 1024 * 200, D, generator=g).to(dev)
n = seq.numel()
case("a11", "rbx_gather_rows", "3 x [1024,200] ids, 1M-row table D=64", n * (4 + 8 * D), lambda: ops.gather_rows(items, seq))
case("a11", "rbx_scatter_add_rows", "3 x [1024,200] ids, 1M-row table D=64", n * (4 + 12 * D),
     lambda: ops.scatter_add_rows(gr, seq.view(-1), 0, g_items))

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", tag + "_kernels.json"), "w") as f:
    json.dump({"peak_gbs": peak, "timing": "CUDA events, median of 15, L2 flushed by a 256 MB read (torch.sum) between reps",
               "kernels": results}, f, indent=1)
