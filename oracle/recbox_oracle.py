"""CPU oracle for the RecBox embedding + feature-interaction hot path.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module; recbox_b200/ never does (the product
path raises when its CUDA library is missing, it does not fall back to this file).

What it is: a function-by-function CPU restatement of the reference's algorithm for every row of
SURVEY.md section 8(a).  The reference is 100 % Python whose arithmetic lives in PyTorch ATen, so
the restatement uses the same ATen CPU ops (torch, fp32) for floating-point rows and numpy for the
integer / index rows.  Each function cites the reference file:line it follows.

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference itself,
SURVEY.md section 8c).  The oracle is therefore pinned against outputs of the UNMODIFIED reference
run in the dev container: oracle/make_golden.py drives /root/reference through oracle/ref_shim.py
and writes tests/golden/*.npz; tests/test_oracle_golden.py checks every function below against
those files (and against the integer known-answer vectors of SURVEY.md section 4).
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# a1  batch matrix -> per-feature columns
# --------------------------------------------------------------------------------------------
def column_index(features, labels=()):
    """recbox/ranking/features.py:106-120 (FeatureMap.set_column_index): consecutive columns in
    feature order, `max_len` columns for a sequence feature, labels after all features."""
    out, pos = {}, 0
    for name, spec in features.items():
        if "max_len" in spec:
            out[name] = list(range(pos, pos + spec["max_len"]))
            pos += spec["max_len"]
        else:
            out[name] = pos
            pos += 1
    for lab in labels:
        out[lab] = pos
        pos += 1
    return out


def get_inputs(batch, features, labels=()):
    """recbox/ranking/pytorch/models/ranking_model.py:106-116: slice one column (or max_len
    columns) of the [B, n_cols] float64 batch per non-meta feature."""
    cols = column_index(features, labels)
    return {name: batch[:, cols[name]] for name, spec in features.items() if spec["type"] != "meta"}


def get_labels(batch, features, labels):
    """ranking_model.py:118-122: label column -> float32 [B,1]."""
    cols = column_index(features, labels)
    return batch[:, cols[labels[0]]].float().view(-1, 1)


def split_batch_numpy(batch, cat_cols, num_cols, label_col=None):
    """Integer/byte view of a1 + the casts of feature_embedding.py:201,204: ids are the float64
    column truncated to int64 (`.long()`), numeric columns are rounded to float32 (`.float()`)."""
    batch = np.asarray(batch, dtype=np.float64)
    ids = batch[:, cat_cols].astype(np.int64)
    dense = batch[:, num_cols].astype(np.float32)
    label = batch[:, label_col].astype(np.float32) if label_col is not None else None
    return ids, dense, label


# --------------------------------------------------------------------------------------------
# a9  sequence pooling
# --------------------------------------------------------------------------------------------
def masked_average_pooling(emb, mask=None):
    """recbox/ranking/pytorch/layers/pooling.py:26-31 and recbox/core/pytorch/layers/sequence.py:8-12
    (identical arithmetic): sum over L divided by the number of rows whose element-sum is non-zero
    (+1e-12)."""
    s = torch.sum(emb, dim=1)
    if mask is None:
        mask = emb.sum(dim=-1) != 0
    return s / (mask.float().sum(-1, keepdim=True) + 1e-12)


def masked_sum_pooling(emb):
    """pooling.py:38-40 / sequence.py:19-20."""
    return torch.sum(emb, dim=1)


# --------------------------------------------------------------------------------------------
# a2 / a3  multi-slot embedding
# --------------------------------------------------------------------------------------------
def embed_feature(x, spec, weight, padding_idx=None):
    """One feature of FeatureEmbeddingDict.forward (feature_embedding.py:199-209) /
    EmbeddingDictLayer.forward (core/pytorch/layers/embedding.py:124-134):
    numeric -> x.float().view(-1,1) @ W^T with W = nn.Linear(1,D,bias=False).weight [D,1];
    categorical / sequence -> nn.Embedding lookup of x.long()."""
    if spec["type"] == "numeric":
        return F.linear(x.float().view(-1, 1), weight)
    if spec["type"] in ("categorical", "sequence"):
        return F.embedding(x.long(), weight, padding_idx=padding_idx)
    raise NotImplementedError(spec["type"])


def embed_dict(X, features, weights, encoders=None, feature_source=(), feature_type=()):
    """FeatureEmbeddingDict.forward feature_embedding.py:188-214.  `weights[name]` is the
    feature's weight (shared features map to the same tensor), `encoders[name]` an optional
    callable applied to the looked-up rows (feature_encoder / embedding_callback)."""
    encoders = encoders or {}
    out = OrderedDict()
    for name, spec in features.items():
        if feature_source and spec.get("source") not in feature_source:
            continue
        if feature_type and spec["type"] not in feature_type:
            continue
        if name not in weights:
            continue
        e = embed_feature(X[name], spec, weights[name], spec.get("padding_idx"))
        if name in encoders:
            e = encoders[name](e)
        out[name] = e
    return out


def dict2tensor(emb_dict, dynamic_emb_dim=False, core_semantics=False):
    """feature_embedding.py:169-186 (stack on dim 1, or cat on the last dim when
    dynamic_emb_dim); core/pytorch/layers/embedding.py:109-114 returns the bare [B,D] tensor
    when exactly one feature is selected (core_semantics=True)."""
    vals = list(emb_dict.values())
    if core_semantics and len(vals) == 1:
        return vals[0]
    if dynamic_emb_dim:
        return torch.cat(vals, dim=-1)
    return torch.stack(vals, dim=1)


def gather_rows_numpy(table, ids):
    """aten::embedding == index_select of rows (a5 forward); the bit-exact path."""
    return np.asarray(table)[np.asarray(ids, dtype=np.int64)]


def embedding_dense_backward_numpy(grad_out, ids, num_rows, padding_idx=None):
    """aten::embedding_dense_backward (a5): zero [V,D] then accumulate grad rows by id, skipping
    padding_idx.  Accumulates in float64 so the result is the correctly rounded reference sum."""
    grad_out = np.asarray(grad_out, dtype=np.float64)
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    g = grad_out.reshape(ids.shape[0], -1)
    out = np.zeros((num_rows, g.shape[1]), dtype=np.float64)
    keep = np.ones_like(ids, dtype=bool) if padding_idx is None else ids != padding_idx
    np.add.at(out, ids[keep], g[keep])
    return out


# --------------------------------------------------------------------------------------------
# a6  FM / inner-product interaction
# --------------------------------------------------------------------------------------------
INTERACTION_MODES = ("product_sum", "bi_interaction", "inner_product", "elementwise_product")


def inner_product_interaction(E, mode="product_sum"):
    """recbox/ranking/pytorch/layers/interactions/inner_product.py:40-56."""
    if mode not in INTERACTION_MODES:
        raise ValueError("InnerProductInteraction output={} is not supported.".format(mode))
    nf = E.shape[1]
    if mode in ("product_sum", "bi_interaction"):
        sum_sq = torch.sum(E, dim=1) ** 2          # inner_product.py:42
        sq_sum = torch.sum(E ** 2, dim=1)          # :43
        bi = (sum_sq - sq_sum) * 0.5               # :44
        return bi if mode == "bi_interaction" else bi.sum(dim=-1, keepdim=True)
    if mode == "inner_product":
        gram = torch.bmm(E, E.transpose(1, 2))     # :50
        mask = torch.triu(torch.ones(nf, nf), 1).bool()
        return torch.masked_select(gram, mask).view(-1, nf * (nf - 1) // 2)
    iu = torch.triu_indices(nf, nf, offset=1)      # :38, :54-56
    return torch.index_select(E, 1, iu[0]) * torch.index_select(E, 1, iu[1])


# --------------------------------------------------------------------------------------------
# a7 / a8  first-order term and FM block
# --------------------------------------------------------------------------------------------
def logistic_regression(X, features, lr_weights, bias=None, encoders=None):
    """recbox/ranking/pytorch/layers/blocks/logistic_regression.py:30-35: a D=1 multi-slot lookup
    (sequence features MaskedSumPooling-ed, feature_embedding.py:72-75) summed over fields + bias."""
    enc = dict(encoders or {})
    for name, spec in features.items():
        if spec["type"] == "sequence" and name in lr_weights:
            enc[name] = masked_sum_pooling
    w = dict2tensor(embed_dict(X, features, lr_weights, enc))
    out = w.sum(dim=1)
    if bias is not None:
        out = out + bias
    return out


def factorization_machine(X, E, features, lr_weights, bias=None):
    """recbox/ranking/pytorch/layers/blocks/factorization_machine.py:30-34."""
    return inner_product_interaction(E, "product_sum") + logistic_regression(X, features, lr_weights, bias)


# --------------------------------------------------------------------------------------------
# a10  two-tower scores
# --------------------------------------------------------------------------------------------
def two_tower_score(u, v):
    """y[b,j] = <u_b, v_{b,j}>, j = 0 the positive (layout fixed by
    recbox/matching/pytorch/models/match_model.py:71-75 and losses/softmax_crossentropy_loss.py:19-21).
    u [B,D]; v [B*(1+negs), D] as produced by collate_fn (h5_generator.py:61-69)."""
    B, D = u.shape
    return torch.bmm(v.view(B, -1, D), u.unsqueeze(-1)).squeeze(-1)


def dssm_score(u, v):
    """third_party/rechub/models/matching/dssm.py:48 (after F.normalize at :57,:65)."""
    return torch.mul(u, v).sum(dim=1)


def inbatch_scores(u, v):
    """third_party/rechub/models/matching/youtube_sbc.py:67: B x B cosine similarity."""
    return torch.cosine_similarity(u.unsqueeze(1), v, dim=2)


def softmax_cross_entropy_loss(y_pred):
    """recbox/core/pytorch/losses/softmax_crossentropy_loss.py:14-21."""
    return -torch.log(F.softmax(y_pred, dim=1)[:, 0]).mean()


# --------------------------------------------------------------------------------------------
# a11  SASRec gather + per-token dot
# --------------------------------------------------------------------------------------------
def sasrec_gather(table, seq, pos, neg, padding_idx=0):
    """third_party/rechub/models/matching/sasrec.py:99-100: three lookups in ONE shared item table
    -> [B,3,L,D]."""
    rows = [F.embedding(t.long(), table, padding_idx=padding_idx) for t in (seq, pos, neg)]
    return torch.stack(rows, dim=1)


def token_dot(seq_out, item_emb):
    """sasrec.py:104-105: (seq_output * pos_embed).sum(-1)."""
    return (seq_out * item_emb).sum(dim=-1)


# --------------------------------------------------------------------------------------------
# a13  dense tail
# --------------------------------------------------------------------------------------------
def mlp(x, layers):
    """recbox/ranking/pytorch/layers/blocks/mlp_block.py:43-61 with ReLU hidden activations, no
    batch-norm / dropout: `layers` = [(W,b), ...]; ReLU after every layer but the last."""
    for i, (W, b) in enumerate(layers):
        x = F.linear(x, W, b)
        if i + 1 < len(layers):
            x = torch.relu(x)
    return x


# --------------------------------------------------------------------------------------------
# f4  GEMM-shaped consumers of [B, F, D]
# --------------------------------------------------------------------------------------------
def cross_net_v2(x0, layers):
    """recbox/ranking/pytorch/layers/interactions/cross_net.py:48-59: X_{i+1} = X_i + X_0 * (W_i X_i + b_i)."""
    xi = x0
    for W, b in layers:
        xi = xi + x0 * F.linear(xi, W, b)
    return xi


def cross_net(x0, layers):
    """cross_net.py:23-45: X_{i+1} = X_i + (w_i . X_i) X_0 + b_i; layers = [(w [1, d], b [d]), ...]."""
    xi = x0
    for w, b in layers:
        xi = xi + (F.linear(xi, w) * x0 + b)
    return xi


def cin(E, convs, fc):
    """compressed_interaction_net.py:35-48: outer product over the field axes, 1x1 convolution over the H*M channels,
    sum-pool over the embedding dim, Linear over the concatenated pools.  convs = [(W [out, H*M], b [out])], fc = (W, b)."""
    B, _, D = E.shape
    x0, xi, pools = E, E, []
    for W, b in convs:
        had = torch.einsum("bhd,bmd->bhmd", x0, xi).reshape(B, -1, D)
        xi = torch.einsum("oc,bcd->bod", W, had) + b.view(1, -1, 1)
        pools.append(xi.sum(dim=-1))
    return F.linear(torch.cat(pools, dim=-1), fc[0], fc[1])


def din_attention(target, hist, mask, mlp_layers, use_softmax=False):
    """ranking/pytorch/layers/attentions/target_attention.py:47-66 with a ReLU scoring MLP (layers as in `mlp`)."""
    L = hist.shape[1]
    t = target.unsqueeze(1).expand(-1, L, -1)
    a = torch.cat([t, hist, t - hist, t * hist], dim=-1)
    w = mlp(a.reshape(-1, a.shape[-1]), mlp_layers).view(-1, L)
    if mask is not None:
        w = w * mask.float()
    if use_softmax:
        if mask is not None:
            w = w + -1.e9 * (1 - mask.float())
        w = w.softmax(dim=-1)
    return (w.unsqueeze(-1) * hist).sum(dim=1)


def multi_head_target_attention(target, hist, mask, Wq, Wk, Wv, Wo, num_heads):
    """target_attention.py:92-121 (use_qkvo, use_scale): one query per sample against its history."""
    B, L, _ = hist.shape
    hd = Wq.shape[0] // num_heads
    q = F.linear(target, Wq).view(B, 1, num_heads, hd).transpose(1, 2)
    k = F.linear(hist, Wk).view(B, L, num_heads, hd).transpose(1, 2)
    v = F.linear(hist, Wv).view(B, L, num_heads, hd).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-1, -2)) / hd ** 0.5
    if mask is not None:
        scores = scores.masked_fill(mask.view(B, 1, 1, L).expand(-1, num_heads, -1, -1).float() == 0, -1.e9)
    out = torch.matmul(scores.softmax(dim=-1), v).transpose(1, 2).reshape(B, num_heads * hd)
    return F.linear(out, Wo)


# --------------------------------------------------------------------------------------------
# a12  clip + optimizer
# --------------------------------------------------------------------------------------------
def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ as called at ranking_model.py:195 / match_model.py:197:
    global L2 over all grads; scale by min(1, max_norm / (norm + 1e-6)).  Returns (norm, coef)."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads]), 2.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam single-tensor update (the dense optimizer of ranking_model.py:62-64 /
    match_model.py:48-51), in the operation order of torch/optim/adam.py::_single_tensor_adam."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------------------------
# a14  batch-level id dedup
# --------------------------------------------------------------------------------------------
def unique_items(item_indexes):
    """recbox/matching/pytorch/dataloaders/h5_generator.py:45-53 (collate_fn_unique): sorted unique
    ids, inverse map, and the FIRST flat position of every unique id."""
    flat = np.asarray(item_indexes).reshape(-1)
    uniq, first, inverse = np.unique(flat, return_index=True, return_inverse=True)
    return uniq, first, inverse.reshape(-1)


def collate_unique_reference(item_indexes):
    """collate_fn_unique's item-id arithmetic, line for line in its own torch calls
    (recbox/matching/pytorch/dataloaders/h5_generator.py:49-52,58): returns (unique, unique_indexes,
    inverse_indexes) exactly as the reference does -- including that the returned inverse is the flipped one."""
    item_indexes = torch.as_tensor(item_indexes)
    unique, inverse_indexes = torch.unique(item_indexes.flatten(), return_inverse=True, sorted=True)
    perm = torch.arange(inverse_indexes.size(0), dtype=inverse_indexes.dtype)
    inverse_indexes, perm = inverse_indexes.flip([0]), perm.flip([0])
    unique_indexes = inverse_indexes.new_empty(unique.size(0)).scatter_(0, inverse_indexes, perm)
    return unique, unique_indexes, inverse_indexes


# --------------------------------------------------------------------------------------------
# (f4)  InteractionMachine
# --------------------------------------------------------------------------------------------
def power_sums(E, order):
    """p_k = (X ** k).sum(dim=1) as interaction_machine.py:47-65 forms them (Q = Q * X, then Q.sum(dim=1)) -> [B, order, D]."""
    out, Q = [], E
    for k in range(order):
        if k:
            Q = Q * E
        out.append(Q.sum(dim=1))
    return torch.stack(out, dim=1)


def interaction_machine(E, order, fc_weight, fc_bias):
    """InteractionMachine.forward without batch norm (ranking/pytorch/layers/interactions/interaction_machine.py:29-70)."""
    P = power_sums(E, order)
    p = [P[:, k] for k in range(order)]
    out = [p[0]]
    if order >= 2:
        out.append((p[0].pow(2) - p[1]) / 2)
    if order >= 3:
        out.append((p[0].pow(3) - 3 * p[0] * p[1] + 2 * p[2]) / 6)
    if order >= 4:
        out.append((p[0].pow(4) - 6 * p[0].pow(2) * p[1] + 3 * p[1].pow(2) + 8 * p[0] * p[2] - 6 * p[3]) / 24)
    if order == 5:
        out.append((p[0].pow(5) - 10 * p[0].pow(3) * p[1] + 20 * p[0].pow(2) * p[2] - 30 * p[0] * p[3]
                    - 20 * p[1] * p[2] + 15 * p[0] * p[1].pow(2) + 24 * p[4]) / 120)
    return torch.nn.functional.linear(torch.cat(out, dim=-1), fc_weight, fc_bias)


# --------------------------------------------------------------------------------------------
# (f2)  negative sampling
# --------------------------------------------------------------------------------------------
def sampling_block(num_items, block_query_indexes, num_negs, user2items_dict, ignore_pos_items=False, seed=None):
    """recbox/matching/pytorch/dataloaders/h5_generator.py:72-95 (sampling_block, uniform sampling_probs): uniform
    draws with replacement; with ignore_pos_items the query user's items get probability 0 and the rest is
    renormalised.  numpy's global MT19937 stream, as in the reference."""
    if seed is not None:
        np.random.seed(seed)
    if ignore_pos_items:
        rows = []
        for q in block_query_indexes:
            probs = np.ones(num_items) / num_items
            probs[user2items_dict[q]] = 0
            probs = probs / np.sum(probs)
            rows.append(np.random.choice(num_items, size=num_negs, replace=True, p=probs))
        return np.array(rows)
    return np.random.choice(num_items, size=(len(block_query_indexes), num_negs), replace=True)


# --------------------------------------------------------------------------------------------
# (f3)  retrieval evaluation
# --------------------------------------------------------------------------------------------
def flat_ip_search(query_vecs, corpus_vecs, topk):
    """faiss.IndexFlatIP(dim).search (recbox/utils/ann/faiss.py:8-14; faiss is a third-party dependency, un-pinned in
    setup.py / requirements.txt and absent from this image): exact inner products in float32, the topk largest per
    query in descending order; ties resolved to the smaller index here; (-inf, -1) padding when the corpus is smaller
    than topk."""
    q = np.asarray(query_vecs, dtype=np.float32)
    c = np.asarray(corpus_vecs, dtype=np.float32)
    scores = q @ c.T
    U, N = scores.shape
    k = min(topk, N)
    order = np.lexsort((np.broadcast_to(np.arange(N), scores.shape), -scores), axis=1)[:, :k]
    top_s = np.take_along_axis(scores, order, axis=1)
    if k < topk:
        top_s = np.concatenate([top_s, np.full((U, topk - k), -np.inf, np.float32)], 1)
        order = np.concatenate([order, np.full((U, topk - k), -1, order.dtype)], 1)
    return top_s, order.astype(np.int64)


class _Metric(object):
    def __init__(self, k=1):
        self.topk = k


class Recall(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:71-80
        return len(set(true_items) & set(topk_items[:self.topk])) / (len(true_items) + 1e-12)


class nRecall(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:83-92
        return len(set(true_items) & set(topk_items[:self.topk])) / min(self.topk, len(true_items) + 1e-12)


class Precision(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:95-104
        return len(set(true_items) & set(topk_items[:self.topk])) / (self.topk + 1e-12)


class F1(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:107-116
        p, r = Precision(self.topk)(topk_items, true_items), Recall(self.topk)(topk_items, true_items)
        return 2 * p * r / (p + r + 1e-12)


class DCG(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:119-132
        true_items = set(true_items)
        return sum(1 / np.log(2 + i) for i, item in enumerate(topk_items[:self.topk]) if item in true_items)


class NDCG(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:135-145
        dcg_fn = DCG(k=self.topk)
        return dcg_fn(topk_items[:self.topk], true_items) / (dcg_fn(true_items[:self.topk], true_items) + 1e-12)


class MRR(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:148-160
        true_items = set(true_items)
        return sum(1 / (i + 1.0) for i, item in enumerate(topk_items[:self.topk]) if item in true_items)


class HitRate(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:163-171
        return 1 if len(set(true_items) & set(topk_items[:self.topk])) > 0 else 0


class MAP(_Metric):
    def __call__(self, topk_items, true_items):          # core/metrics.py:174-190
        true_items = set(true_items)
        pos, precision = 0, 0
        for i, item in enumerate(topk_items[:self.topk]):
            if item in true_items:
                pos += 1
                precision += pos / (i + 1.0)
        return precision / (pos + 1e-12)


def evaluate_block(user_embs, corpus_vecs, query_indices, train_user2items, valid_user2items, metric_funcs, max_topk,
                   search_topk=500):
    """core/metrics.py:52-68: top-500 search, train items pushed down by -1e9, argsort, first max_topk, metrics."""
    scores, indices = flat_ip_search(user_embs, corpus_vecs, search_topk)
    ok = indices >= 0
    mask = np.zeros((len(user_embs), len(corpus_vecs)))
    for i, q in enumerate(query_indices):
        mask[i, train_user2items[q]] = 1
    mask = np.where(ok, np.take_along_axis(mask, np.maximum(indices, 0), axis=1), 0)
    scores = scores + (-1e9 * mask).astype(np.float32)
    sorted_idxs = np.argsort(-scores, axis=1, kind="stable")
    topk_items = np.take_along_axis(indices, sorted_idxs, axis=1)[:, 0:max_topk]
    true_items = [valid_user2items[q] for q in query_indices]
    return topk_items, [[f(preds, labels) for f in metric_funcs] for preds, labels in zip(topk_items, true_items)]


def evaluate_metrics(user_embs, item_embs, train_user2items, valid_user2items, query_indices, metrics, search_topk=500):
    """core/metrics.py:11-50 (single worker): {metric string: mean over users}."""
    ns = {"Recall": Recall, "nRecall": nRecall, "Precision": Precision, "F1": F1, "DCG": DCG, "NDCG": NDCG, "MRR": MRR,
          "HitRate": HitRate, "MAP": MAP}
    funcs, max_topk = [], 0
    for m in metrics:
        try:
            funcs.append(eval(m, {}, ns))
            max_topk = max(max_topk, int(m.split("k=")[-1].strip(")")))
        except Exception:
            raise NotImplementedError("metrics={} not implemented.".format(m))
    results = []
    for idx in range(0, len(user_embs), 1000):
        _, r = evaluate_block(user_embs[idx:idx + 1000], item_embs, query_indices[idx:idx + 1000], train_user2items,
                              valid_user2items, funcs, max_topk, search_topk)
        results += r
    return dict(zip(metrics, np.average(np.array(results), axis=0).tolist()))


# --------------------------------------------------------------------------------------------
# (f1)  touched-rows optimizers: the published torch.optim algorithms run on CPU tensors
# --------------------------------------------------------------------------------------------
def touched_rows_step(kind, w, g, state, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip=1.0):
    """One update of table w [R, D] with dense gradient g restricted to its non-zero ("touched") rows.
    kind: "sgd" / "adagrad" (torch.optim.SGD / Adagrad: exact on a dense gradient, untouched rows do not move),
    "adam_rows" (adam_step of this file on the touched rows only), "sparse_adam" (torch.optim.SparseAdam).
    state: dict with "m", "v" tensors (updated in place).  Returns the new w."""
    w = w.clone()
    g = g * clip
    rows = torch.nonzero(g.reshape(g.shape[0], -1).abs().sum(1) > 0).reshape(-1)
    if kind == "sgd":
        w[rows] = w[rows] - lr * g[rows]
    elif kind == "adagrad":
        p = torch.nn.Parameter(w.clone())
        opt = torch.optim.Adagrad([p], lr=lr, eps=eps)
        opt.state[p]["sum"] = state["v"].clone()
        opt.state[p]["step"] = torch.tensor(float(step - 1))
        p.grad = g.clone()
        opt.step()
        state["v"].copy_(opt.state[p]["sum"])
        w = p.detach().clone()
    elif kind == "adam_rows":
        wr, mr, vr = w[rows].clone(), state["m"][rows].clone(), state["v"][rows].clone()
        adam_step(wr, g[rows].clone(), mr, vr, step, lr, beta1, beta2, eps)
        w[rows], state["m"][rows], state["v"][rows] = wr, mr, vr
    elif kind == "sparse_adam":
        p = torch.nn.Parameter(w.clone())
        opt = torch.optim.SparseAdam([p], lr=lr, betas=(beta1, beta2), eps=eps)
        opt.state[p]["step"] = step - 1
        opt.state[p]["exp_avg"] = state["m"].clone()
        opt.state[p]["exp_avg_sq"] = state["v"].clone()
        p.grad = torch.sparse_coo_tensor(rows.unsqueeze(0), g[rows], size=g.shape).coalesce()
        opt.step()
        state["m"].copy_(opt.state[p]["exp_avg"])
        state["v"].copy_(opt.state[p]["exp_avg_sq"])
        w = p.detach().clone()
    else:
        raise ValueError(kind)
    return w


# --------------------------------------------------------------------------------------------
# (e)  row-shard routing around the all-to-all  (new; defined in SURVEY.md section 8e)
# --------------------------------------------------------------------------------------------
def shard_route(global_rows, world):
    """Bucket global row ids by owner (row % world), keeping first-come order inside a bucket.
    Returns (send_local_rows grouped by owner, counts[world], pos) with
    send[pos[i]] == global_rows.flat[i] // world."""
    r = np.asarray(global_rows, dtype=np.int64).reshape(-1)
    owner = r % world
    order = np.argsort(owner, kind="stable")
    counts = np.bincount(owner, minlength=world).astype(np.int64)
    pos = np.empty_like(order)
    pos[order] = np.arange(r.shape[0])
    return (r // world)[order], counts, pos


def shard_unroute(recv_rows, pos):
    """Inverse permutation: rows come back in send order; out[i] = recv[pos[i]]."""
    return np.asarray(recv_rows)[np.asarray(pos)]


# --------------------------------------------------------------------------------------------
# Assembled DeepFM (configs 1, 2, 4): the reference train step restated end to end
# --------------------------------------------------------------------------------------------
class DeepFMOracle:
    """DeepFM / FM assembled from the reference layers the way FuxiCTR's model zoo does
    (SURVEY.md Appendix A): E = FeatureEmbedding(X); y = FM(X,E) [+ MLP(E.flatten(1))];
    sigmoid; BCE(mean); clip_grad_norm_(10); Adam.  Parameters are plain leaf tensors keyed by the
    reference's state_dict names so weights can be exchanged with the reference and the product."""

    def __init__(self, features, labels, D, hidden=(400, 400, 400), use_mlp=True, lr=1e-3,
                 max_grad_norm=10.0, seed=0):
        self.features, self.labels, self.D = features, list(labels), D
        self.use_mlp, self.lr, self.max_grad_norm = use_mlp, lr, max_grad_norm
        g = torch.Generator().manual_seed(seed)
        P = OrderedDict()
        nf = 0
        for name, spec in features.items():
            if spec["type"] == "meta":
                continue
            nf += 1
            for prefix, d in (("embedding_layer.embedding_layer.embedding_layers.", D),
                              ("fm_layer.lr_layer.embedding_layer.embedding_layer.embedding_layers.", 1)):
                if spec["type"] == "numeric":   # xavier_normal_ (ranking_model.py:93-99)
                    w = torch.randn(d, 1, generator=g) * math.sqrt(2.0 / (1 + d))
                else:                           # normal(std=1e-4) on rows 1.. (feature_embedding.py:126-137)
                    w = torch.randn(spec["vocab_size"], d, generator=g) * 1e-4
                    if spec.get("padding_idx") is not None:
                        w[spec["padding_idx"]] = 0
                P[prefix + name + ".weight"] = w
        P["fm_layer.lr_layer.bias"] = torch.zeros(1)
        if use_mlp:
            dims = [nf * D] + list(hidden) + [1]
            for i in range(len(dims) - 1):
                std = math.sqrt(2.0 / (dims[i] + dims[i + 1]))
                P["mlp.mlp.%d.weight" % (2 * i)] = torch.randn(dims[i + 1], dims[i], generator=g) * std
                P["mlp.mlp.%d.bias" % (2 * i)] = torch.zeros(dims[i + 1])
        self.params = OrderedDict((k, v.requires_grad_(True)) for k, v in P.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
        self.step = 0

    def _weights(self, prefix):
        n = len(prefix)
        return {k[n:-len(".weight")]: v for k, v in self.params.items() if k.startswith(prefix)}

    def forward(self, batch):
        X = get_inputs(batch, self.features, self.labels)
        E = dict2tensor(embed_dict(X, self.features,
                                   self._weights("embedding_layer.embedding_layer.embedding_layers.")))
        y = factorization_machine(
            X, E, self.features,
            self._weights("fm_layer.lr_layer.embedding_layer.embedding_layer.embedding_layers."),
            self.params["fm_layer.lr_layer.bias"])
        if self.use_mlp:
            names = sorted((k for k in self.params if k.startswith("mlp.mlp.") and k.endswith("weight")),
                           key=lambda s: int(s.split(".")[2]))
            layers = [(self.params[k], self.params[k[:-6] + "bias"]) for k in names]
            y = y + mlp(E.flatten(start_dim=1), layers)
        return torch.sigmoid(y)

    def loss(self, batch):
        y_true = get_labels(batch, self.features, self.labels)
        return F.binary_cross_entropy(self.forward(batch), y_true, reduction="mean")

    def train_step(self, batch):
        """ranking_model.py:191-197."""
        for p in self.params.values():
            p.grad = None
        loss = self.loss(batch)
        loss.backward()
        ps = [p for p in self.params.values()]
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in ps]
        _, coef = clip_grad_norm(grads, self.max_grad_norm)
        self.step += 1
        with torch.no_grad():
            for (k, p), g in zip(self.params.items(), grads):
                adam_step(p, g * coef, self.m[k], self.v[k], self.step, lr=self.lr)
        return loss.detach()
