#!/usr/bin/env python
"""Model-level number next to bench.py's hot-path number (last bullet): a whole DeepFM train step
on BASELINE configs[1] (26 cat + 13 dense, 26 x 38 462 rows, D = 16, MLP 400-400-400-1 ReLU, B = 65 536) built from the
layer API of recbox_b200.layers exactly the way a RecBox model is:

    zero_grad -> get_inputs -> FeatureEmbedding -> FactorizationMachine + MLP -> sigmoid -> BCE(mean) -> backward
    -> clip_grad_norm_(all params, 10) -> Adam.step        (RankingModel.train_step, ranking_model.py:191-197)

The embedding / FM part runs on the fused kernels; the MLP either on cuBLAS (nn.Linear, the round-1 state) or on the tcgen05
GEMM chain of csrc/gemm.cu behind layers.MLP_Block (a13); the optimizer is torch.optim.Adam over every parameter (the
reference's dense semantics); the `graph` variants replay the whole step from a CUDA graph.
The CPU arm is oracle.DeepFMOracle (the reference train step restated) on the host cores, a few steps.

  python tools/train_step_bench.py [tag]  -> gpurun_out/<tag>_train_step.json
"""
import json
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from recbox_b200 import layers  # noqa: E402
from recbox_b200.features import FeatureMap  # noqa: E402
from recbox_b200.loader import PackedDataLoader  # noqa: E402

B, F, Fn, D, V = 65536, 26, 13, 16, 38462
tag = sys.argv[1] if len(sys.argv) > 1 else "ts"
dev = torch.device("cuda")


def feature_map():
    fm = FeatureMap("criteo_synth", ".")
    for i in range(Fn):
        fm.features["I%d" % (i + 1)] = {"source": "", "type": "numeric"}
    for i in range(F):
        fm.features["C%d" % (i + 1)] = {"source": "", "type": "categorical", "vocab_size": V, "padding_idx": 0}
    fm.finalize(["label"])
    fm.default_emb_dim = D
    return fm


class DeepFM(nn.Module):
    def __init__(self, fm, ours=False):
        super().__init__()
        self.feature_map, self.device = fm, dev
        self.embedding_layer = layers.FeatureEmbedding(fm, D)
        self.fm_layer = layers.FactorizationMachine(fm)
        if ours:          # a13 on the tcgen05 GEMM chain (csrc/gemm.cu) behind the reference's MLP_Block constructor
            self.mlp = layers.MLP_Block(input_dim=fm.sum_emb_out_dim(), hidden_units=[400, 400, 400], hidden_activations="ReLU",
                                        output_dim=1)
            return
        mods, d = [], fm.sum_emb_out_dim()
        for h in (400, 400, 400):
            mods += [nn.Linear(d, h), nn.ReLU()]
            d = h
        mods.append(nn.Linear(d, 1))
        self.mlp = nn.Sequential(*mods)

    def forward(self, inputs):
        X = layers.get_inputs(self, inputs)
        E = self.embedding_layer(X)
        y = self.fm_layer(X, E) + self.mlp(E.flatten(start_dim=1))
        return torch.sigmoid(y)


def make_batches(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        out.append(np.concatenate([rng.random((B, Fn)), rng.integers(1, V, (B, F)).astype(np.float64),
                                   (rng.random((B, 1)) < 0.5).astype(np.float64)], 1))
    return out


def run_gpu(tf32, packed, steps=30, warmup=5, ours=False, precision=3, graph=False, fused_adam=False):
    """ours: MLP_Block on the tcgen05 GEMM (precision 3 = 3xTF32 fp32-level, 1 = plain TF32) instead of nn.Linear on cuBLAS;
    graph: the whole train step (device batch -> loss -> backward -> clip -> Adam) replayed from a CUDA graph."""
    from recbox_b200 import graphs, ops
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    ops.GEMM_PRECISION = precision
    torch.manual_seed(0)
    fm = feature_map()
    model = DeepFM(fm, ours=ours).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=graph, fused=True if fused_adam else None)
    host = make_batches(4, 1)
    if packed:
        feed = [list(PackedDataLoader(fm, h, batch_size=B).bind(model.embedding_layer))[0] for h in host]
    else:
        feed = [torch.from_numpy(h).pin_memory() for h in host]

    def train(batch):
        y_true = layers.get_labels(model, batch)
        opt.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy(model(batch), y_true, reduction="mean")
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        return loss

    if graph:
        static = feed[0].to(dev)                                 # the graph reads its batch from these device blocks
        gs = graphs.GraphedStep(lambda: train(static), warmup=3)

        def step(i):
            feed[i % len(feed)].copy_into(static)                # H2D of the step's packed batch, then the replay
            return gs()
    else:
        def step(i):
            return train(feed[i % len(feed)])

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        loss = step(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    ops.GEMM_PRECISION = 3
    return {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "input": "packed" if packed else "float64",
            "mlp": ("tcgen05 GEMM chain, %s" % ("3xTF32 (fp32-level)" if precision == 3 else "plain TF32")) if ours else
                   ("nn.Linear / cuBLAS, %s" % ("TF32" if tf32 else "fp32")),
            "optimizer": "torch.optim.Adam(fused=True)" if fused_adam else "torch.optim.Adam (foreach, the reference's default)",
            "issue": "cuda-graph replay" if graph else "eager", "loss": float(loss), "steps": steps}


def run_cpu(steps=2):
    from collections import OrderedDict as OD
    from helpers import oracle
    torch.set_num_threads(os.cpu_count() or 1)
    feats = OD()
    for i in range(Fn):
        feats["I%d" % (i + 1)] = {"source": "", "type": "numeric"}
    for i in range(F):
        feats["C%d" % (i + 1)] = {"source": "", "type": "categorical", "vocab_size": V, "padding_idx": 0}
    m = oracle.DeepFMOracle(feats, ["label"], D, hidden=(400, 400, 400), seed=0)
    batches = [torch.from_numpy(h) for h in make_batches(2, 1)]
    m.train_step(batches[0])
    t0 = time.perf_counter()
    for i in range(steps):
        m.train_step(batches[i % 2])
    dt = (time.perf_counter() - t0) / steps
    return {"ms_per_step": dt * 1e3, "samples_per_s": B / dt, "cores": torch.get_num_threads(), "steps": steps,
            "what": "oracle.DeepFMOracle.train_step (the reference's dense train step restated), B=65536"}


if __name__ == "__main__":
    out = OrderedDict()
    out["config"] = "BASELINE configs[1] DeepFM: 26 cat + 13 dense, 26x38462 rows, D=16, MLP 400-400-400-1, B=65536, dense Adam + clip 10"
    variants = [("cublas_fp32_mlp_float64_batches", dict(tf32=False, packed=False)),
                ("cublas_fp32_mlp_packed_batches", dict(tf32=False, packed=True)),
                ("cublas_tf32_mlp_packed_batches", dict(tf32=True, packed=True)),
                ("ours_3xtf32_mlp_packed_batches", dict(tf32=False, packed=True, ours=True, precision=3)),
                ("ours_3xtf32_mlp_packed_batches_graph", dict(tf32=False, packed=True, ours=True, precision=3, graph=True)),
                ("ours_3xtf32_mlp_fused_adam_graph", dict(tf32=False, packed=True, ours=True, precision=3, graph=True, fused_adam=True)),
                ("ours_tf32_mlp_fused_adam_graph", dict(tf32=False, packed=True, ours=True, precision=1, graph=True, fused_adam=True))]
    only = os.environ.get("TS_ONLY")            # one variant, few steps (launch lists under ncu)
    if only:
        variants = [(n, dict(kw, steps=int(os.environ.get("TS_STEPS", "3")), warmup=2)) for n, kw in variants if n == only]
    for name, kw in variants:
        try:
            out[name] = run_gpu(**kw)
        except Exception as e:
            out[name] = {"error": repr(e)[:300]}
            torch.cuda.synchronize()
    if not only:
        out["cpu_reference_port"] = run_cpu()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", tag + "_train_step.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, v in out.items():
        print(k, v)
