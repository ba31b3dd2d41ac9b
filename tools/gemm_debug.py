"""Bring-up / tuning harness of the tcgen05 GEMM (csrc/gemm.cu): error against float64 and time per configuration.
    python tools/gemm_debug.py [--time]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recbox_b200 import ops  # noqa: E402


def run(M, N, K, a_mn, b_mn, prec, env, time_it=False):
    for k in ("RBX_GEMM_KB", "RBX_GEMM_BN", "RBX_GEMM_STAGES", "RBX_GEMM_SPLITS", "RBX_GEMM_WRITE_HI"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) * 0.1
    a = (A.t().contiguous() if a_mn else A).cuda()
    b = (B.t().contiguous() if b_mn else B).cuda()
    want = A.double() @ B.double().t() if M * N * K < 2e10 else None
    try:
        got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, precision=prec)
        torch.cuda.synchronize()
    except Exception as e:
        print("M=%d N=%d K=%d a_mn=%d b_mn=%d prec=%d %s: EXC %r" % (M, N, K, a_mn, b_mn, prec, env, e))
        return
    msg = ""
    if want is not None:
        err = (got.double().cpu() - want).abs()
        rel = float(err.max() / want.abs().max())
        msg = "max_rel=%.2e" % rel
        if rel > (1e-5 if prec == 3 else 4e-3):
            bad = err > 1e-3 * want.abs().max()
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            msg += " BAD frac=%.3f rows[%d..%d]n=%d cols[%d..%d]n=%d nan=%d" % (
                float(bad.float().mean()), int(rows.min()) if len(rows) else -1, int(rows.max()) if len(rows) else -1, len(rows),
                int(cols.min()) if len(cols) else -1, int(cols.max()) if len(cols) else -1, len(cols), int(torch.isnan(got).sum()))
    if time_it:
        for _ in range(3):
            ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, precision=prec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        msg += " %.1f us %.1f TFLOP/s(fp32-equivalent)" % (ms * 1e3, 2.0 * M * N * K / ms / 1e9)
    print("M=%d N=%d K=%d a_mn=%d b_mn=%d prec=%d %s: %s" % (M, N, K, a_mn, b_mn, prec, env, msg), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--one", default="", help="M,N,K,a_mn,b_mn,prec: one configuration, three launches (for ncu)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.one:
        M, N, K, a_mn, b_mn, prec = (int(x) for x in args.one.split(","))
        run(M, N, K, a_mn, b_mn, prec, {}, False)
        return
    if not args.time:
        for prec in (1, 3):
            for KB in (32, 16):
                run(128, 64, 64, 0, 0, prec, {"RBX_GEMM_KB": KB})
                run(256, 400, 624, 0, 0, prec, {"RBX_GEMM_KB": KB})
                run(256, 416, 624, 0, 1, prec, {"RBX_GEMM_KB": KB})
                run(256, 416, 624, 1, 0, prec, {"RBX_GEMM_KB": KB})
                run(400, 624, 4096, 1, 1, prec, {"RBX_GEMM_KB": KB})
        run(256, 400, 624, 0, 0, 3, {"RBX_GEMM_WRITE_HI": 1})
        run(256, 400, 624, 0, 0, 3, {"RBX_GEMM_STAGES": 2})
        return
    B = 65536
    for prec in (3, 1):
        for env in ({}, {"RBX_GEMM_KB": 32}, {"RBX_GEMM_STAGES": 2}, {"RBX_GEMM_STAGES": 3}, {"RBX_GEMM_KB": 32, "RBX_GEMM_BN": 128, "RBX_GEMM_STAGES": 2},
                    {"RBX_GEMM_BN": 256}):
            run(B, 400, 624, 0, 0, prec, env, True)
        run(B, 400, 400, 0, 0, prec, {}, True)
        run(B, 624, 400, 0, 1, prec, {}, True)
        run(B, 400, 400, 0, 1, prec, {}, True)
        run(400, 624, B, 1, 1, prec, {}, True)
        run(400, 400, B, 1, 1, prec, {}, True)
        run(400, 624, B, 1, 1, prec, {"RBX_GEMM_KB": 32}, True)


if __name__ == "__main__":
    main()
