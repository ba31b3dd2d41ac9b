"""Host-side cost of one hot-path step through the layer API (FeatureEmbedding + FactorizationMachine modules, autograd
backward) on configs[1], against the same step replayed from a CUDA graph:  python tools/layer_overhead.py [--profile]"""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from recbox_b200 import layers, loader  # noqa: E402
from recbox_b200.features import FeatureMap  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    CFG = bench.CFG
    B, F, Fn, D = CFG["B"], CFG["F"], CFG["Fn"], CFG["D"]
    fmap = bench.criteo_feature_map(FeatureMap)
    torch.manual_seed(1)
    emb = layers.FeatureEmbedding(fmap, D).to(dev)
    fml = layers.FactorizationMachine(fmap).to(dev)
    params = list(emb.parameters()) + list(fml.parameters())
    g = torch.Generator().manual_seed(2)
    M = bench.host_batch_matrix(B, "uniform", 3, g)
    ds = loader.PackedDataset(fmap, M)
    pb = ds.batch(0, B).to(dev)
    dE = (torch.randn(B, F + Fn, D, generator=g) * 1e-3).to(dev)
    d_out = (torch.randn(B, 1, generator=g) * 1e-3).to(dev)

    class _Model(object):
        feature_map, device = fmap, dev

    def step():
        for p in params:
            p.grad = None
        X = layers.get_inputs(_Model, pb)
        E = emb(X)
        y = fml(X, E)
        torch.autograd.backward([E, y], [dE, d_out])
        return y

    for _ in range(10):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("eager: %.1f us/step host issue, %.1f us/step wall, %d parameters" % (t_issue / args.steps * 1e6, t_all / args.steps * 1e6, len(params)))
    if args.profile:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(args.steps):
            step()
        pr.disable()
        torch.cuda.synchronize()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35)
        print(s.getvalue()[:6000])
    from recbox_b200 import graphs
    y_ref = step().detach().clone()               # (a live autograd graph would pin AccumulateGrad nodes to this stream)
    gref = [p.grad.clone() for p in params]
    gs = graphs.GraphedStep(step, warmup=3)          # after the capture the parameters' .grad are the graph's static tensors
    for p in params:
        p.grad.zero_()
    y = gs()
    torch.cuda.synchronize()
    print("graph replay == eager:", torch.equal(y, y_ref), all(torch.allclose(p.grad, r, rtol=1e-5, atol=1e-7) for p, r in zip(params, gref)))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        gs()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print("graph: %.1f us/step host issue, %.1f us/step wall" % (t_issue / args.steps * 1e6, t_all / args.steps * 1e6))


if __name__ == "__main__":
    main()
