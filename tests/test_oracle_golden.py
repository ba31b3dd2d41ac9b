"""Pin the CPU oracle (oracle/recbox_oracle.py) against the golden vectors minted from the
UNMODIFIED reference by oracle/make_golden.py (tests/golden/*.npz) and against the integer
known-answer vectors of SURVEY.md section 4.  CPU only."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from helpers import assert_close, oracle

def same_sum(a, b, what=""):
    """Floating-point SUMS (reductions over the batch / fields): bit-exact when the host runs ATen with the
    reduction split the goldens were minted with, else within 1e-6 (ATen's mm / sum partial order
    depends on the host's thread count).  Gathers stay torch.equal."""
    if not torch.equal(a, b):
        assert_close(a, b, rtol=1e-6, atol_scale=1e-6, what=what)
    return True


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(z[k]) if z[k].dtype != np.bool_ else torch.from_numpy(z[k]) for k in z.files}


def ranking_features(tag):
    seq = tag.endswith("seq")
    share = tag.endswith("share")
    f = OrderedDict()
    f["I1"] = {"source": "", "type": "numeric"}
    f["C1"] = {"source": "", "type": "categorical", "vocab_size": 11, "padding_idx": 0}
    f["I2"] = {"source": "", "type": "numeric"}
    f["C2"] = {"source": "", "type": "categorical", "vocab_size": 7, "padding_idx": 0}
    f["C3"] = {"source": "", "type": "categorical", "vocab_size": 13, "padding_idx": 0}
    f["C4"] = {"source": "", "type": "categorical", "vocab_size": 11, "padding_idx": 0}
    if share:
        f["C4"]["share_embedding"] = "C1"
    if seq:
        f["S1"] = {"source": "", "type": "sequence", "vocab_size": 9, "padding_idx": 0, "max_len": 5,
                   "feature_encoder": "layers.MaskedAveragePooling()"}
    return f


EMB = "emb.embedding_layer.embedding_layers."
LRP = "fm.lr_layer.embedding_layer.embedding_layer.embedding_layers."


def weights_from(gold, prefix, features, share_ok=True):
    W = OrderedDict()
    for name, spec in features.items():
        key = prefix + name + ".weight"
        if share_ok and "share_embedding" in spec:      # one module registered under two names
            W[name] = W[spec["share_embedding"]]
        elif key in gold:
            W[name] = gold[key].clone().requires_grad_(True)
    return W


# ------------------------------------------------------------------------------------------- KATs
def test_known_answers_survey_section4():
    E = torch.arange(24, dtype=torch.float32).view(2, 3, 4)
    assert oracle.inner_product_interaction(E, "product_sum").reshape(-1).tolist() == [314, 3626]
    assert oracle.inner_product_interaction(E, "bi_interaction").reshape(-1).tolist() == [32, 59, 92, 131, 752, 851, 956, 1067]
    assert oracle.inner_product_interaction(E, "inner_product").reshape(-1).tolist() == [38, 62, 214, 950, 1166, 1510]
    assert oracle.inner_product_interaction(E, "elementwise_product")[0].reshape(-1).tolist() == [0, 5, 12, 21, 0, 9, 20, 33, 32, 45, 60, 77]
    Ev = E.clone().requires_grad_(True)
    oracle.inner_product_interaction(Ev, "product_sum").sum().backward()
    assert Ev.grad[0].tolist() == [[12, 14, 16, 18], [8, 10, 12, 14], [4, 6, 8, 10]]
    with pytest.raises(ValueError):
        oracle.inner_product_interaction(E, "nope")
    W = np.arange(10, dtype=np.float32).reshape(5, 2)
    idx = np.array([[1, 1, 0], [4, 0, 0]])
    g = oracle.embedding_dense_backward_numpy(np.ones((2, 3, 2), np.float32), idx, 5, padding_idx=0)
    assert g.tolist() == [[0, 0], [2, 2], [0, 0], [0, 0], [1, 1]]
    emb = torch.from_numpy(oracle.gather_rows_numpy(W, idx))
    assert_close(oracle.masked_average_pooling(emb), torch.tensor([[4 / 3, 7 / 3], [8 / 3, 11 / 3]]), what="masked avg")
    # DSSM-style softmax loss at (near-)zero scores with 10 negatives = ln 11; SASRec BCE at init = 2 ln 2
    assert abs(float(oracle.softmax_cross_entropy_loss(torch.zeros(4, 11))) - np.log(11)) < 1e-6


# ---------------------------------------------------------------------------- golden: interaction
@pytest.mark.parametrize("tag", ["kat", "rnd"])
@pytest.mark.parametrize("mode", oracle.INTERACTION_MODES)
def test_interaction_golden(tag, mode):
    g = load("interaction")
    E = g[tag + ".E"].clone().requires_grad_(True)
    y = oracle.inner_product_interaction(E, mode)
    assert same_sum(y, g["%s.%s.out" % (tag, mode)])
    (y * g["%s.%s.w" % (tag, mode)]).sum().backward()
    assert same_sum(E.grad, g["%s.%s.dE" % (tag, mode)])


def test_pooling_golden():
    g = load("pooling")
    assert same_sum(oracle.masked_average_pooling(g["emb"]), g["ranking_avg"])
    assert same_sum(oracle.masked_average_pooling(g["emb"]), g["core_avg"])
    assert same_sum(oracle.masked_average_pooling(g["emb"], g["mask"]), g["ranking_avg_mask"])
    assert same_sum(oracle.masked_sum_pooling(g["emb"]), g["ranking_sum"])
    assert same_sum(oracle.masked_sum_pooling(g["emb"]), g["core_sum"])


# ------------------------------------------------------------------------- golden: ranking layers
@pytest.mark.parametrize("tag", ["ranking_layers_d8", "ranking_layers_d10", "ranking_layers_share", "ranking_layers_seq"])
def test_ranking_layers_golden(tag):
    g = load(tag)
    feats = ranking_features(tag)
    W = weights_from(g, EMB, feats)
    X = oracle.get_inputs(g["batch"], feats, ["label"])
    enc = {"S1": oracle.masked_average_pooling} if "S1" in feats else None
    E = oracle.dict2tensor(oracle.embed_dict(X, feats, W, enc))
    assert torch.equal(E, g["E"])
    loss = (E * g["wE"]).sum()
    W1 = None
    if "y" in g:
        W1 = weights_from(g, LRP, feats, share_ok=False)
        bias = g["fm.lr_layer.bias"].clone().requires_grad_(True)
        lr = oracle.logistic_regression(X, feats, W1, bias)
        fm = oracle.inner_product_interaction(E, "product_sum")
        y = oracle.factorization_machine(X, E, feats, W1, bias)
        assert same_sum(lr, g["lr_out"]) and same_sum(fm, g["fm_out"]) and same_sum(y, g["y"])
        loss = loss + (y * g["wy"]).sum()
    loss.backward()
    for name in feats:
        k = "grad." + EMB + name + ".weight"
        if k in g:
            assert same_sum(W[name].grad, g[k]), k
        k1 = "grad." + LRP + name + ".weight"
        if W1 is not None and k1 in g:
            assert same_sum(W1[name].grad, g[k1]), k1
    if W1 is not None:
        assert same_sum(bias.grad, g["grad.fm.lr_layer.bias"])


def test_core_layers_golden():
    g = load("core_layers")
    feats = OrderedDict([
        ("item_id", {"type": "categorical", "source": "item", "vocab_size": 23, "padding_idx": 22}),
        ("item_cat", {"type": "categorical", "source": "item", "vocab_size": 6}),
        ("user_id", {"type": "categorical", "source": "user", "vocab_size": 17}),
        ("user_age", {"type": "numeric", "source": "user"}),
        ("user_hist", {"type": "sequence", "source": "user", "vocab_size": 23, "padding_idx": 22,
                       "share_embedding": "item_id"}),
    ])
    W = weights_from(g, EMB, feats)
    X = {k[2:]: v for k, v in g.items() if k.startswith("X.")}
    enc = {"user_hist": oracle.masked_average_pooling}
    U = oracle.dict2tensor(oracle.embed_dict(X, feats, W, enc, feature_source=("user",)), core_semantics=True)
    V = oracle.dict2tensor(oracle.embed_dict(X, feats, W, enc, feature_source=("item",)), core_semantics=True)
    assert same_sum(U, g["U"]) and same_sum(V, g["V"])
    ((U * g["wU"]).sum() + (V * g["wV"]).sum()).backward()
    for name in ("item_id", "item_cat", "user_id", "user_age"):
        assert same_sum(W[name].grad, g["grad." + EMB + name + ".weight"]), name
    one = oracle.dict2tensor(oracle.embed_dict(X, OrderedDict(user_id=feats["user_id"]), W), core_semantics=True)
    assert one.dim() == 2 and torch.equal(one, g["single"])


def test_two_tower_golden():
    g = load("two_tower")
    u, v = g["u"].clone().requires_grad_(True), g["v"].clone().requires_grad_(True)
    y = oracle.two_tower_score(u, v)
    assert same_sum(y, g["y"])
    loss = oracle.softmax_cross_entropy_loss(y)
    assert same_sum(loss, g["loss"])
    loss.backward()
    assert same_sum(u.grad, g["du"]) and same_sum(v.grad, g["dv"])
    assert same_sum(oracle.dssm_score(g["dssm_u"], g["dssm_v"]), g["dssm_y"])


# ------------------------------------------------------------------- golden: assembled train step
def _deepfm_oracle(g, feats, D, hidden, use_mlp):
    m = oracle.DeepFMOracle(feats, ["label"], D, hidden=hidden, use_mlp=use_mlp)
    P = OrderedDict()
    for k in m.params:
        P[k] = g["init." + k].clone().requires_grad_(True)
    m.params = P
    m.m = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    m.v = OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    return m


def test_deepfm_train_golden():
    g = load("deepfm_train")
    feats = ranking_features("ranking_layers_d8")
    m = _deepfm_oracle(g, feats, 8, (16, 8), True)
    assert sorted(m.params) == sorted(k[5:] for k in g if k.startswith("init.")), "state_dict keys"
    losses = []
    for step in range(3):
        losses.append(float(m.train_step(g["batch%d" % step])))
        if step == 0:
            pass
    assert_close(torch.tensor(losses), g["losses"].float(), rtol=1e-6, atol_scale=0, what="losses")
    for k, v in m.params.items():
        assert_close(v, g["final." + k], rtol=1e-5, atol_scale=1e-6, what=k)
    assert_close(m.forward(g["batch0"]), g["pred_final"], rtol=1e-5, what="pred")


def test_config1_fm_golden():
    """BASELINE configs[0]: FM on the 1k-row Criteo-shaped CSV the reference preprocessed."""
    g = load("config1_fm")
    feats = OrderedDict()
    for i in range(1, 14):
        feats["I%d" % i] = {"source": "", "type": "numeric"}
    for i, V in enumerate(g["vocab_sizes"].tolist(), 1):
        feats["C%d" % i] = {"source": "", "type": "categorical", "vocab_size": int(V), "padding_idx": 0}
    m = _deepfm_oracle(g, feats, 10, (), False)
    batch = g["batch"]
    assert_close(m.forward(batch[:128]), g["pred_init"], rtol=1e-6, what="pred_init")
    losses = [float(m.train_step(batch[s * 128:(s + 1) * 128])) for s in range(4)]
    assert_close(torch.tensor(losses), g["losses"].float(), rtol=1e-6, atol_scale=0, what="losses")
    assert_close(m.forward(batch[:128]), g["pred_final"], rtol=1e-5, what="pred_final")


# ----------------------------------------------------------------------------- integer / index rows
def test_split_batch_and_column_index():
    feats = ranking_features("ranking_layers_seq")
    cols = oracle.column_index(feats, ["label"])
    assert cols["I1"] == 0 and cols["S1"] == [6, 7, 8, 9, 10] and cols["label"] == 11
    batch = np.array([[0.25, 3.0, 0.5, 2.0, 12.0, 10.0, 1, 2, 0, 0, 0, 1.0]])
    ids, dense, label = oracle.split_batch_numpy(batch, [1, 3, 4, 5], [0, 2], 11)
    assert ids.tolist() == [[3, 2, 12, 10]] and dense.dtype == np.float32 and label.tolist() == [1.0]


def test_unique_items_first_occurrence():
    uniq, first, inv = oracle.unique_items(np.array([[5, 3, 5], [9, 3, 1]]))
    assert uniq.tolist() == [1, 3, 5, 9] and first.tolist() == [5, 1, 0, 3] and inv.tolist() == [2, 1, 2, 3, 1, 0]


@pytest.mark.parametrize("world", [1, 2, 8])
def test_shard_route_roundtrip(world):
    rng = np.random.default_rng(world)
    rows = rng.integers(0, 10_000, size=1000)
    send, counts, pos = oracle.shard_route(rows, world)
    assert counts.sum() == 1000 and np.array_equal(np.sort(pos), np.arange(1000))
    starts = np.concatenate([[0], np.cumsum(counts)])
    for w in range(world):
        seg = slice(starts[w], starts[w + 1])
        orig = rows[rows % world == w]
        assert np.array_equal(send[seg] * world + w, orig), "bucket keeps first-come order"
    payload = rng.standard_normal((1000, 4))
    sent = np.empty_like(payload)
    sent[pos] = payload
    assert np.array_equal(oracle.shard_unroute(sent, pos), payload)


def test_clip_and_adam_match_torch():
    g = torch.Generator().manual_seed(0)
    p = torch.randn(50, generator=g, requires_grad=True)
    q = p.detach().clone()
    opt = torch.optim.Adam([p], lr=1e-3)
    m, v = torch.zeros(50), torch.zeros(50)
    for step in range(1, 4):
        grad = torch.randn(50, generator=g) * 5
        p.grad = grad.clone()
        total = torch.nn.utils.clip_grad_norm_([p], 10.0)
        norm, coef = oracle.clip_grad_norm([grad], 10.0)
        assert torch.equal(norm, total)
        opt.step()
        oracle.adam_step(q, grad * coef, m, v, step)
        assert_close(q, p, rtol=1e-6, atol_scale=1e-7, what="adam")


def test_collate_unique_matches_reference_golden():
    """a14: collate_unique.npz holds what the reference's own collate_fn_unique returned
    (oracle/make_golden.py::golden_collate_unique)."""
    z = np.load(os.path.join(GOLD, "collate_unique.npz"))
    for k in range(int(z["n_cases"])):
        item = z["item_indexes_%d" % k]
        uniq, uidx, inv = oracle.collate_unique_reference(item)
        assert np.array_equal(uniq.numpy(), z["unique_%d" % k])
        assert np.array_equal(uidx.numpy(), z["unique_indexes_%d" % k])
        assert np.array_equal(inv.numpy(), z["inverse_indexes_%d" % k])
        u2, f2, i2 = oracle.unique_items(item)            # the numpy form the GPU tests compare with
        assert np.array_equal(u2, z["unique_%d" % k]) and np.array_equal(f2, z["unique_indexes_%d" % k])
        assert np.array_equal(i2[::-1], z["inverse_indexes_%d" % k]), "the reference returns the flipped inverse"


@pytest.mark.parametrize("kind", ["sgd", "adagrad"])
def test_touched_rows_exact_kinds_equal_dense_torch_optimizers(kind):
    """f1: for SGD / Adagrad a touched-rows update IS the dense torch optimizer step (zero-gradient rows do not move)."""
    g = torch.Generator().manual_seed(1)
    w = torch.randn(30, 4, generator=g)
    grad = torch.zeros(30, 4)
    grad[[2, 9, 17]] = torch.randn(3, 4, generator=g)
    p = torch.nn.Parameter(w.clone())
    opt = (torch.optim.SGD if kind == "sgd" else torch.optim.Adagrad)([p], lr=0.05)
    state = {"m": torch.zeros(30, 4), "v": torch.zeros(30, 4)}
    q = w
    for step in range(1, 4):
        p.grad = grad.clone() * step
        opt.step()
        q = oracle.touched_rows_step(kind, q, grad * step, state, step, lr=0.05, eps=1e-10)
        assert_close(q, p.detach(), rtol=1e-6, atol_scale=1e-7, what=kind)


def test_touched_rows_adam_is_lazy_not_dense():
    """f1: "adam_rows" leaves untouched rows alone -- documented difference from the reference's dense Adam."""
    w = torch.ones(6, 2)
    grad = torch.zeros(6, 2)
    grad[1] = 0.5
    state = {"m": torch.full((6, 2), 0.1), "v": torch.full((6, 2), 0.01)}
    q = oracle.touched_rows_step("adam_rows", w, grad, state, 2)
    assert torch.equal(q[0], w[0]) and not torch.equal(q[1], w[1])
    assert float(state["m"][0, 0]) == pytest.approx(0.1)     # dense Adam would have decayed it to 0.09


# ------------------------------------------------------------------------------- f4 InteractionMachine
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_interaction_machine_golden(order):
    """oracle.interaction_machine / power_sums against the reference's InteractionMachine (no batch norm)."""
    g = load("interaction_machine")
    tag = "o%d_bn0." % order
    X = g[tag + "X"].clone().requires_grad_(True)
    y = oracle.interaction_machine(X, order, g[tag + "sd.fc.weight"], g[tag + "sd.fc.bias"])
    same_sum(y, g[tag + "y"], "y")
    (y * g[tag + "w"]).sum().backward()
    same_sum(X.grad, g[tag + "dX"], "dX")
    P = oracle.power_sums(g[tag + "X"], order)
    assert P.shape == (23, order, 8)
    same_sum(P[:, 0], g[tag + "X"].sum(1), "p1")


# ------------------------------------------------------------------------------- a11 SASRec gathers + token dots
def test_sasrec_gather_and_token_dot_golden():
    """oracle.sasrec_gather / token_dot against rechub's SASRec (third_party/rechub/models/matching/sasrec.py:98-107):
    the [B,3,L,D] shared-table lookup is bit-exact, the logits are the token dots with the model's own sequence
    output, and the lookups route the gradient back into one table (pad row 0 of `seq` gets none)."""
    g = load("sasrec_gather")
    table = g["table"].clone().requires_grad_(True)
    emb = oracle.sasrec_gather(table, g["seq"], g["pos"], g["neg"], padding_idx=0)
    assert emb.shape == g["emb"].shape and torch.equal(emb, g["emb"])
    same_sum(oracle.token_dot(g["seq_out"], emb[:, 1]), g["pos_logits"], "pos_logits")
    same_sum(oracle.token_dot(g["seq_out"], emb[:, 2]), g["neg_logits"], "neg_logits")
    # rows never looked up (and the pad row) carry no gradient in the reference either
    used = torch.zeros(table.shape[0], dtype=torch.bool)
    for k in ("seq", "pos", "neg"):
        used[g[k].reshape(-1).long()] = True
    used[0] = False
    assert torch.equal(g["table_grad"][~used], torch.zeros_like(g["table_grad"][~used]))
    assert float(g["table_grad"][used].abs().sum()) > 0


def test_inbatch_scores_and_mlp_restatements():
    """a10 in-batch score matrix (youtube_sbc.py:67: cosine similarity of every user with every item, positives on the
    diagonal) and the a13 dense tail used by DeepFMOracle."""
    g = torch.Generator().manual_seed(3)
    u, v = torch.randn(5, 4, generator=g), torch.randn(5, 4, generator=g)
    s = oracle.inbatch_scores(u, v)
    assert s.shape == (5, 5)
    for i in range(5):
        for j in range(5):
            assert abs(float(s[i, j]) - float((u[i] * v[j]).sum() / (u[i].norm() * v[j].norm()))) < 1e-6
    x = torch.randn(3, 4, generator=g)
    W1, b1, W2, b2 = torch.randn(6, 4, generator=g), torch.randn(6, generator=g), torch.randn(1, 6, generator=g), torch.randn(1, generator=g)
    want = torch.relu(x @ W1.t() + b1) @ W2.t() + b2
    assert torch.allclose(oracle.mlp(x, [(W1, b1), (W2, b2)]), want, atol=1e-6)


def test_mlp_restatement_matches_reference_mlp_block():
    """a13: oracle.mlp (ReLU hidden layers, linear head) against the reference's own MLP_Block / MLP_Layer outputs."""
    g = load("mlp_block")
    for tag, idx in (("plain", (0, 2, 4)), ("core", (0, 2, 4))):
        layers_ = [(g["%s.init.mlp.%d.weight" % (tag, i)], g["%s.init.mlp.%d.bias" % (tag, i)]) for i in idx]
        x = g["x"].clone().requires_grad_(True)
        ws = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in layers_]
        y = oracle.mlp(x, ws)
        assert torch.allclose(y, g[tag + ".y"], rtol=1e-6, atol=1e-6)
        (y * g[tag + ".w"]).sum().backward()
        assert torch.allclose(x.grad, g[tag + ".dx"], rtol=1e-5, atol=1e-6)
        for (W, b), i in zip(ws, idx):
            assert torch.allclose(W.grad, g["%s.grad.mlp.%d.weight" % (tag, i)], rtol=1e-5, atol=1e-6)
            assert torch.allclose(b.grad, g["%s.grad.mlp.%d.bias" % (tag, i)], rtol=1e-5, atol=1e-6)


def test_f4_block_restatements_match_reference():
    """f4: CrossNet / CrossNetV2 / CIN / DIN attention / multi-head target attention restated in the oracle against the
    outputs of the reference's own modules (tests/golden/blocks.npz)."""
    g = load("blocks")
    E, target, hist, mask = g["E"], g["target"], g["hist"], g["mask"]
    flat = E.flatten(1)
    p = lambda tag, k: g["%s.init.%s" % (tag, k)]
    y = oracle.cross_net_v2(flat, [(p("crossnetv2", "cross_layers.%d.weight" % i), p("crossnetv2", "cross_layers.%d.bias" % i)) for i in range(3)])
    assert torch.allclose(y, g["crossnetv2.y"], rtol=1e-5, atol=1e-6)
    y = oracle.cross_net(flat, [(p("crossnet", "cross_net.%d.weight.weight" % i), p("crossnet", "cross_net.%d.bias" % i)) for i in range(2)])
    assert torch.allclose(y, g["crossnet.y"], rtol=1e-5, atol=1e-6)
    convs = [(p("cin", "cin_layer.layer_%d.weight" % i).squeeze(-1), p("cin", "cin_layer.layer_%d.bias" % i)) for i in (1, 2)]
    y = oracle.cin(E, convs, (p("cin", "fc.weight"), p("cin", "fc.bias")))
    assert torch.allclose(y, g["cin.y"], rtol=1e-5, atol=1e-5)
    for tag, idx, sm in (("din", (0, 2), False), ("din_softmax", (0, 2, 4), True)):
        ls = [(p(tag, "attention_layer.mlp.%d.weight" % i), p(tag, "attention_layer.mlp.%d.bias" % i)) for i in idx]
        y = oracle.din_attention(target, hist, mask, ls, use_softmax=sm)
        assert torch.allclose(y, g[tag + ".y"], rtol=1e-5, atol=1e-6), tag
    y = oracle.multi_head_target_attention(target, hist, mask, p("mhta", "W_q.weight"), p("mhta", "W_k.weight"), p("mhta", "W_v.weight"),
                                           p("mhta", "W_o.weight"), 2)
    assert torch.allclose(y, g["mhta.y"], rtol=1e-5, atol=1e-6)
