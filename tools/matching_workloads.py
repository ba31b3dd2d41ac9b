"""BASELINE configs[2] (two-tower DSSM, 10 M items, in-batch / sampled negatives) and configs[4] (SASRec, L = 200, 1 M items)
as bench workloads: `python bench.py --workload dssm|sasrec [--gpus N]` (cfg 3 / cfg 5).

One step = the whole train step of the matching model: embedding lookups (user table, pooled history, item table -- the hot
path's kernels: rbx_gather_rows, rbx_pooled_gather_fwd/bwd, rbx_rowdot, rbx_scatter_add_rows), the towers / attention block
(every Linear on the tcgen05 GEMM), loss, autograd backward with the table gradients scattered into PERSISTENT dense gradient
tables, and the touched-rows update that consumes and clears them (f1; no O(table) memset or dense optimizer pass).  The
launch-bound step (B = 8 192 / 1 024 sequences) is replayed from a CUDA graph (recbox_b200.graphs.GraphedStep).

Multi-GPU: replicas only (the 2.6 GB / 256 MB tables fit every GPU): each rank runs its own batch shard and the
replicas stay ONE model: the touched rows' gradients are exchanged as (row id, gradient row) blocks (replica.SparseRowExchange:
rbx_unique_ids -> rbx_gather_rows -> all_gather -> rbx_scatter_add_rows), the dense tower / block gradients by NCCL.

Reference shapes: rechub DSSM third_party/rechub/models/matching/dssm.py:39-65, YoutubeSBC youtube_sbc.py:58-84 (in-batch
negatives), SASRec sasrec.py:65-107; towers are rechub MLP (basic/layers.py:233-266) without batch-norm / dropout."""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ids(n, shape, kind, seed, lo=1):
    rng = np.random.default_rng(seed)
    if kind == "zipf":
        x = np.minimum(rng.zipf(1.05, size=shape), n - 1)
    else:
        x = rng.integers(lo, n, size=shape)
    return torch.from_numpy(x.astype(np.int32))


class _TableRows(torch.autograd.Function):
    """Lookup whose backward scatter-adds into a persistent dense gradient table (consumed + cleared by the touched-rows
    optimizer) instead of materialising a fresh [rows, D] gradient per use of the table."""

    @staticmethod
    def forward(ctx, anchor, table, g_table, ids, pad, pool_mode):
        from recbox_b200 import ops
        ctx.g_table, ctx.pad, ctx.pool_mode = g_table, pad, pool_mode
        if pool_mode is None:
            ctx.save_for_backward(ids)
            return ops.gather_rows(table, ids)
        out, cnt = ops.pooled_gather_fwd(table, ids, pool_mode)
        ctx.save_for_backward(ids, cnt)
        return out

    @staticmethod
    def backward(ctx, g):
        from recbox_b200 import ops
        if ctx.pool_mode is None:
            (ids,) = ctx.saved_tensors
            ops.scatter_add_rows(g.contiguous(), ids, ctx.pad, ctx.g_table)
        else:
            ids, cnt = ctx.saved_tensors
            ops.pooled_gather_bwd(g.contiguous(), ids, cnt, ctx.pad, ctx.g_table, ctx.pool_mode)
        return None, None, None, None, None, None


def _time_kernel(fn, reps=20):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _timed_steps(step, K, W, world, graph=True):
    """-> (ms per step as max over ranks, 'graph' | 'eager')."""
    import torch.distributed as dist
    from recbox_b200 import graphs
    mode = "eager"
    run = step
    if graph:
        try:
            gs = graphs.GraphedStep(step, warmup=3)
            run, mode = gs, "cuda-graph replay"
        except Exception as e:                       # capture is an optimisation, not the product: report and run eagerly
            sys.stderr.write("GraphedStep failed, running eagerly: %r\n" % (e,))
            torch.cuda.synchronize()
    for _ in range(max(W, 3)):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), mode


# =====================================================================================================================
# configs[2]: two-tower DSSM
# =====================================================================================================================
def run_dssm(args, rank, world, dev, bench):
    from recbox_b200 import layers, ops, optim
    from recbox_b200.blocks import linear
    B, D, L, negs = 8192, 64, 20, 10
    n_items, n_users = 10_000_000, 1_000_000
    pad = n_items                                             # padding row = last row of the item table (SURVEY 8d cfg 3)
    gen = torch.Generator(device=dev).manual_seed(20240 + 3 + rank)
    item_table = torch.empty(n_items + 1, D, device=dev).normal_(0, 0.01, generator=gen)
    item_table[pad].zero_()
    user_table = torch.empty(n_users, D, device=dev).normal_(0, 0.01, generator=gen)
    g_item, g_user = torch.zeros_like(item_table), torch.zeros_like(user_table)
    torch.manual_seed(20240 + 3)
    user_mlp = layers.MLP_Layer(2 * D, None, [256, 128, 64], dropout_rates=0.0).to(dev)
    item_mlp = layers.MLP_Layer(D, None, [256, 128, 64], dropout_rates=0.0).to(dev)
    tower_params = list(user_mlp.parameters()) + list(item_mlp.parameters())
    tower_opt = torch.optim.SGD(tower_params, lr=1e-3)
    opt_item = optim.TouchedRowsOptimizer([(item_table, g_item)], kind="sgd", lr=1e-3)
    opt_user = optim.TouchedRowsOptimizer([(user_table, g_user)], kind="sgd", lr=1e-3)
    ex_item = ex_user = None
    if world > 1:                     # replicas: sparse (row id, gradient row) exchange of the touched rows, dense tower grads by NCCL
        from recbox_b200 import replica
        ex_item = replica.SparseRowExchange(n_items + 1, D, B * (L + 1 + negs), dev)
        ex_user = replica.SparseRowExchange(n_users, D, B, dev)
    NB = 4
    batches = []
    for i in range(NB):
        s = 1000 * rank + i
        hist = _ids(n_items, (B, L), args.ids, s, lo=0)
        lens = torch.from_numpy(np.random.default_rng(s + 7).integers(1, L + 1, size=(B, 1)))
        hist = torch.where(torch.arange(L)[None, :] < lens, hist, torch.full_like(hist, pad))
        batches.append(dict(user=_ids(n_users, (B,), "uniform", s + 1, lo=0).to(dev), hist=hist.to(dev),
                            pos=_ids(n_items, (B,), args.ids, s + 2, lo=0).to(dev),
                            neg=_ids(n_items, (B, negs), "uniform", s + 3, lo=0).to(dev)))
    anchor = torch.zeros(1, device=dev, requires_grad=True)
    target_diag = torch.arange(B, device=dev)
    target_zero = torch.zeros(B, dtype=torch.long, device=dev)
    static = {k: v.clone() for k, v in batches[0].items()}     # the graph's input tensors; each step's batch is copied in

    def make_step(variant):
        def step():
            b = static
            ue = _TableRows.apply(anchor, user_table, g_user, b["user"], -1, None)
            hp = _TableRows.apply(anchor, item_table, g_item, b["hist"], pad, 1)          # masked average pooling
            u = F.normalize(user_mlp(torch.cat([ue, hp], 1)), p=2, dim=1)
            if variant == "inbatch":
                v = F.normalize(item_mlp(_TableRows.apply(anchor, item_table, g_item, b["pos"], pad, None)), p=2, dim=1)
                scores = linear(u, v)                                                     # [B, B] = u v^T on the tensor cores
                loss = F.cross_entropy(scores / 0.05, target_diag)
                touched = torch.cat([b["hist"].reshape(-1), b["pos"]])
            else:
                items = torch.cat([b["pos"].view(-1, 1), b["neg"]], 1)                    # [B, 1 + negs], positive first
                v = F.normalize(item_mlp(_TableRows.apply(anchor, item_table, g_item, items.reshape(-1), pad, None)), p=2, dim=1)
                scores = layers.two_tower_score(u, v.view(B, 1 + negs, D))                # [B, 11] row dots
                loss = F.cross_entropy(scores / 0.05, target_zero)
                touched = torch.cat([b["hist"].reshape(-1), items.reshape(-1)])
            for p in tower_params:
                p.grad = None
            loss.backward()
            users = b["user"]
            if world > 1:
                layers.sync_replica_gradients(tower_params, average=True)
                touched, users = ex_item.exchange(g_item, touched), ex_user.exchange(g_user, users)
            tower_opt.step()
            opt_item.step(touched)
            opt_user.step(users)
            return loss
        return step

    out = {}
    K, W = min(args.steps, 50), args.warmup
    for variant in ("inbatch", "sampled"):
        ms, mode = _timed_steps(make_step(variant), K, W, world, graph=world == 1)     # (collectives stay outside CUDA graphs)
        out[variant] = {"ms_per_step": ms, "value": B * world / (ms * 1e-3), "issue": mode}
    # the hot-path kernels alone (one launch each, inputs of the first batch), against the HBM roofline
    peak, peak_src = bench.peaks()
    b = batches[0]
    gE = torch.randn(B, D, device=dev)
    u = torch.randn(B, D, device=dev)
    v = torch.randn(B, 1 + negs, D, device=dev)
    dy = torch.randn(B, 1 + negs, device=dev)
    _, cnt = ops.pooled_gather_fwd(item_table, b["hist"], 1)
    items11 = torch.cat([b["pos"].view(-1, 1), b["neg"]], 1).reshape(-1)
    g11 = torch.randn(B * (1 + negs), D, device=dev)
    kern = {}

    def add(name, fn, nbytes):
        ms = _time_kernel(fn)
        kern[name] = {"ms": ms, "alg_bytes": nbytes, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    add("pooled_gather_fwd [B,20] x D=64", lambda: ops.pooled_gather_fwd(item_table, b["hist"], 1), B * (L * (4 + 4 * D) + 4 * D))
    add("pooled_gather_bwd", lambda: ops.pooled_gather_bwd(gE, b["hist"], cnt, pad, g_item, 1), B * (L * (4 + 8 * D) + 4 * D))
    add("gather_rows [B*11] x D=64", lambda: ops.gather_rows(item_table, items11), B * 11 * (4 + 8 * D))
    add("scatter_add_rows [B*11]", lambda: ops.scatter_add_rows(g11, items11, pad, g_item), B * 11 * (4 + 12 * D))
    add("rowdot_fwd [B,11,64]", lambda: ops.rowdot_fwd(u, v), B * (4 * D * (1 + 11) + 4 * 11))
    add("rowdot_bwd", lambda: ops.rowdot_bwd(u, v, dy), B * (4 * D * (1 + 11) * 2 + 4 * 11))
    g_item.zero_()
    dom = max(kern, key=lambda k: kern[k]["ms"])
    line = None
    if rank == 0:
        line = {
            "metric": "samples/sec on synthetic two-tower batches (whole train step: lookups + towers + loss + backward + touched-rows update)",
            "value": out["inbatch"]["value"], "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": out["inbatch"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: two-tower DSSM, 10 000 001-row item table shared by the pooled user history (L=20, masked "
                                   "average) and the item tower, 1 000 000 users, D=64, towers 256-128-64, B=8192 per GPU, in-batch negatives "
                                   "([B,B] score GEMM, softmax CE); `sampled` = 10 sampled negatives ([B,11] row dots)",
                       "global_batch": B * world, "ids": args.ids, "batches_rotated": 1,
                       "l2": "item table 2.56 GB exceeds the 126 MB L2",
                       "parallelism": "1 GPU" if world == 1 else "dp%d replicas of the tables; sparse all_gather of the touched (row id, gradient row) blocks + "
                                                                   "NCCL all-reduce of the tower gradients inside the step" % world},
            "variants": out, "gpu_launches_issue": out["inbatch"]["issue"],
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"],
                         "traffic": None, "peak_source": peak_src},
            "kernels": kern,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = _cpu_dssm(B, D, L, n_items, n_users, args.ids)
    return line


def _cpu_dssm(B, D, L, n_items, n_users, ids_kind):
    """The same in-batch step in torch CPU ops with the reference's dense-gradient nn.Embedding tables, on all host cores."""
    torch.set_num_threads(os.cpu_count() or 1)
    item = torch.nn.Embedding(n_items + 1, D, padding_idx=n_items)
    user = torch.nn.Embedding(n_users, D)
    mk = lambda d: torch.nn.Sequential(torch.nn.Linear(d, 256), torch.nn.ReLU(), torch.nn.Linear(256, 128), torch.nn.ReLU(),
                                       torch.nn.Linear(128, 64), torch.nn.ReLU())
    um, im = mk(2 * D), mk(D)
    hist = _ids(n_items, (B, L), ids_kind, 1, lo=0).long()
    uid, pos = _ids(n_users, (B,), "uniform", 2, lo=0).long(), _ids(n_items, (B,), ids_kind, 3, lo=0).long()
    params = list(item.parameters()) + list(user.parameters()) + list(um.parameters()) + list(im.parameters())

    def step():
        for p in params:
            p.grad = None
        h = item(hist)
        m = (h.sum(-1) != 0).float()
        hp = h.sum(1) / (m.sum(-1, keepdim=True) + 1e-12)
        u = F.normalize(um(torch.cat([user(uid), hp], 1)), p=2, dim=1)
        v = F.normalize(im(item(pos)), p=2, dim=1)
        F.cross_entropy(u @ v.t() / 0.05, torch.arange(B)).backward()
    step()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    return {"value": B / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of the full B=%d in-batch step, forward + backward only (dense nn.Embedding gradients as the reference; %.0f ms/step)"
                      % (n, B, dt * 1e3)}


# =====================================================================================================================
# configs[4]: SASRec
# =====================================================================================================================
def run_sasrec(args, rank, world, dev, bench):
    from recbox_b200 import ops, optim
    from recbox_b200.blocks import linear
    B, L, D, blocks = 1024, 200, 64, 2
    n_items = 1_000_000                                         # + pad row 0
    gen = torch.Generator(device=dev).manual_seed(20240 + 5 + rank)
    table = torch.empty(n_items + 1, D, device=dev).normal_(0, 0.01, generator=gen)
    table[0].zero_()
    g_table = torch.zeros_like(table)
    torch.manual_seed(20240 + 5)
    pos_emb = torch.nn.Embedding(L, D).to(dev)
    ln = lambda: torch.nn.LayerNorm(D, eps=1e-8).to(dev)
    attn_ln, fwd_ln, last_ln = [ln() for _ in range(blocks)], [ln() for _ in range(blocks)], ln()
    lin = lambda i, o: torch.nn.Linear(i, o).to(dev)
    qkv, proj = [lin(D, 3 * D) for _ in range(blocks)], [lin(D, D) for _ in range(blocks)]
    ff1, ff2 = [lin(D, D) for _ in range(blocks)], [lin(D, D) for _ in range(blocks)]
    mods = [pos_emb, last_ln] + attn_ln + fwd_ln + qkv + proj + ff1 + ff2
    dense_params = [p for m in mods for p in m.parameters()]
    dense_opt = torch.optim.SGD(dense_params, lr=1e-3)
    opt_tab = optim.TouchedRowsOptimizer([(table, g_table)], kind="sgd", lr=1e-3)
    ex_tab = None
    if world > 1:
        from recbox_b200 import layers, replica
        ex_tab = replica.SparseRowExchange(n_items + 1, D, 3 * B * L, dev)
    rng = np.random.default_rng(100 + rank)
    lens = torch.from_numpy(rng.integers(20, L + 1, size=(B, 1)))
    keep = torch.arange(L)[None, :] >= (L - lens)               # left-padded
    mk = lambda s: torch.where(keep, _ids(n_items + 1, (B, L), args.ids, s, lo=1), torch.zeros(B, L, dtype=torch.int32)).to(dev)
    seq, pos, neg = mk(1000 * rank + 1), mk(1000 * rank + 2), mk(1000 * rank + 3)
    ids3 = torch.stack([seq, pos, neg], 0).contiguous()         # the three shared-table lookups of sasrec.py:99-100
    anchor = torch.zeros(1, device=dev, requires_grad=True)
    positions = torch.arange(L, device=dev)
    tmask = (seq != 0).unsqueeze(-1).float()
    valid_w = (pos != 0).float()
    n_valid = valid_w.sum().clamp_min(1.0)

    def step():
        e3 = _TableRows.apply(anchor, table, g_table, ids3, 0, None)             # [3, B, L, D] in one launch
        x = (e3[0] * D ** 0.5 + pos_emb(positions)) * tmask
        for i in range(blocks):
            q_in = attn_ln[i](x)
            q = linear(q_in, qkv[i].weight[:D], qkv[i].bias[:D])                 # queries from the normed input, keys / values from x
            kv = linear(x, qkv[i].weight[D:], qkv[i].bias[D:])                   # (nn.MultiheadAttention(Q, x, x), sasrec.py:83-86)
            a = F.scaled_dot_product_attention(q.unsqueeze(1), kv[..., :D].unsqueeze(1), kv[..., D:].unsqueeze(1), is_causal=True).squeeze(1)
            x = q_in + linear(a, proj[i].weight, proj[i].bias)
            x = fwd_ln[i](x)
            x = (x + linear(linear(x, ff1[i].weight, ff1[i].bias, relu=True), ff2[i].weight, ff2[i].bias)) * tmask
        out = last_ln(x)
        pos_logits, neg_logits = (out * e3[1]).sum(-1), (out * e3[2]).sum(-1)   # token dots (sasrec.py:104-105)
        # BCE over the non-padding positions (rechub's trainer masks with pos != 0); written as a weighted mean so that nothing
        # sizes itself on the host and the whole step can be captured into a CUDA graph
        pl, nl = F.logsigmoid(pos_logits), F.logsigmoid(-neg_logits)
        loss = -((pl + nl) * valid_w).sum() / n_valid
        for p in dense_params:
            p.grad = None
        loss.backward()
        touched = ids3
        if world > 1:
            layers.sync_replica_gradients(dense_params, average=True)
            touched = ex_tab.exchange(g_table, ids3)
        dense_opt.step()
        opt_tab.step(touched)
        return loss

    K, W = min(args.steps, 30), args.warmup
    ms, mode = _timed_steps(step, K, W, world, graph=world == 1)
    peak, peak_src = bench.peaks()
    g3 = torch.randn(3, B, L, D, device=dev)
    kern = {}

    def add(name, fn, nbytes):
        t = _time_kernel(fn)
        kern[name] = {"ms": t, "alg_bytes": nbytes, "gbs": nbytes / t / 1e6, "frac": nbytes / t / 1e6 / peak}
    add("gather_rows 3 x [B,L] x D=64", lambda: ops.gather_rows(table, ids3), 3 * B * L * (4 + 8 * D))
    add("scatter_add_rows 3 x [B,L]", lambda: ops.scatter_add_rows(g3, ids3, 0, g_table), 3 * B * L * (4 + 12 * D))
    g_table.zero_()
    dom = max(kern, key=lambda k: kern[k]["ms"])
    line = None
    if rank == 0:
        line = {
            "metric": "sequences/sec on synthetic SASRec batches (whole train step)",
            "value": B * world / (ms * 1e-3), "unit": "sequences/s", "tokens_per_s": B * L * world / (ms * 1e-3), "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: SASRec, 1 000 001-row item table shared by the sequence / positive / negative lookups, "
                                   "L=200 (lengths U[20,200], left-padded), D=64, 2 blocks, 1 head, B=1024 sequences per GPU; projections / FFN on "
                                   "the tcgen05 GEMM, attention core = torch SDPA (library)",
                       "global_batch": B * world, "ids": args.ids, "issue": mode,
                       "parallelism": "1 GPU" if world == 1 else "dp%d replicas (pure data parallel sweep); sparse all_gather of the touched (row id, gradient row) "
                                                                   "blocks + NCCL all-reduce of the dense gradients inside the step" % world},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"],
                         "traffic": None, "peak_source": peak_src},
            "kernels": kern,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = _cpu_sasrec(L, D, blocks, n_items, args.ids)
    return line


def _cpu_sasrec(L, D, blocks, n_items, ids_kind, B=128):
    """The same step in torch CPU ops with a dense-gradient nn.Embedding table (as rechub's SASRec), forward + backward, on a
    bounded sample of B sequences."""
    torch.set_num_threads(os.cpu_count() or 1)
    emb = torch.nn.Embedding(n_items + 1, D, padding_idx=0)
    pos_emb = torch.nn.Embedding(L, D)
    lns = [torch.nn.LayerNorm(D, eps=1e-8) for _ in range(2 * blocks + 1)]
    attn = [torch.nn.MultiheadAttention(D, 1, 0.0) for _ in range(blocks)]
    ff = [(torch.nn.Linear(D, D), torch.nn.Linear(D, D)) for _ in range(blocks)]
    mods = [emb, pos_emb] + lns + attn + [m for pair in ff for m in pair]
    params = [p for m in mods for p in m.parameters()]
    seq, pos, neg = (_ids(n_items + 1, (B, L), ids_kind, s, lo=1).long() for s in (1, 2, 3))
    causal = ~torch.tril(torch.ones(L, L, dtype=torch.bool))

    def step():
        for p in params:
            p.grad = None
        x = emb(seq) * D ** 0.5 + pos_emb(torch.arange(L))
        for i in range(blocks):
            xt = x.transpose(0, 1)
            q = lns[2 * i](xt)
            a, _ = attn[i](q, xt, xt, attn_mask=causal)
            x = (q + a).transpose(0, 1)
            x = lns[2 * i + 1](x)
            x = x + ff[i][1](torch.relu(ff[i][0](x)))
        out = lns[-1](x)
        pl, nl = (out * emb(pos)).sum(-1), (out * emb(neg)).sum(-1)
        (F.binary_cross_entropy_with_logits(pl, torch.ones_like(pl)) + F.binary_cross_entropy_with_logits(nl, torch.zeros_like(nl))).backward()
    step()
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        step()
    dt = (time.perf_counter() - t0) / n
    return {"value": B / dt, "unit": "sequences/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d steps of B=%d sequences (L=%d), forward + backward only, dense nn.Embedding gradients as the reference (%.0f ms/step)"
                      % (n, B, L, dt * 1e3)}
