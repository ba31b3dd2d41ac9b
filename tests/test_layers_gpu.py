"""The layer API (recbox_b200.layers) on cuda:0 against the golden vectors minted from the UNMODIFIED
reference (tests/golden/*.npz, oracle/make_golden.py): same feature maps, same weights (loaded
through state_dict under the reference's parameter names), same inputs -> same outputs and grads.
Gathered rows bit-exact; sums within 1e-5 relative (fp32, BASELINE.json north_star)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch
from torch import nn

from helpers import assert_close
from test_layers_host import feature_map
from test_oracle_golden import load

from recbox_b200 import layers
from recbox_b200.features import MatchingFeatureMap

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sub(g, prefix):
    return OrderedDict((k[len(prefix):], v) for k, v in g.items() if k.startswith(prefix))


def _X(fm, batch, packed):
    batch = batch.to(DEV)
    if packed:
        return layers.PackedInputs(fm, batch)
    return {f: batch[:, fm.get_column_index(f)] for f, s in fm.features.items() if s["type"] != "meta"}


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("tag,D", [("ranking_layers_d8", 8), ("ranking_layers_d10", 10), ("ranking_layers_share", 16)])
def test_ranking_embedding_fm_lr(tag, D, packed):
    g = load(tag)
    fm = feature_map(tag, D)
    emb = layers.FeatureEmbedding(fm, D)
    fml = layers.FactorizationMachine(fm)
    emb.load_state_dict(_sub(g, "emb."))
    fml.load_state_dict(_sub(g, "fm."))
    emb.to(DEV)
    fml.to(DEV)
    X = _X(fm, g["batch"], packed)
    E = emb(X)
    assert E.shape == g["E"].shape and torch.equal(E.cpu(), g["E"]), "E must be bit-exact"
    lr_out = fml.lr_layer(X)
    fm_out = fml.fm_layer(E)
    y = fml(X, E)
    assert_close(lr_out, g["lr_out"], what="lr_out")
    assert_close(fm_out, g["fm_out"], atol_scale=2e-5, what="fm_out")
    assert_close(y, g["y"], atol_scale=2e-5, what="y")
    loss = (E * g["wE"].to(DEV)).sum() + (y * g["wy"].to(DEV)).sum()
    loss.backward()
    for prefix, mod in (("grad.emb.", emb), ("grad.fm.", fml)):
        want = _sub(g, prefix)
        got = {k: p.grad for k, p in mod.named_parameters()}
        assert sorted(got) == sorted(want)
        for k in want:
            assert got[k] is not None, k
            assert_close(got[k], want[k], atol_scale=2e-5, what=prefix + k)
    # second call: FactorizationMachine's first-order term now rides in the embedding launch
    for p in list(emb.parameters()) + list(fml.parameters()):
        p.grad = None
    X2 = _X(fm, g["batch"], packed)
    E2 = emb(X2)
    y2 = fml(X2, E2)
    assert torch.equal(E2.cpu(), g["E"])
    assert_close(y2, g["y"], atol_scale=2e-5, what="y (fused launch)")
    ((E2 * g["wE"].to(DEV)).sum() + (y2 * g["wy"].to(DEV)).sum()).backward()
    for prefix, mod in (("grad.emb.", emb), ("grad.fm.", fml)):
        want = _sub(g, prefix)
        for k, p in mod.named_parameters():
            assert_close(p.grad, want[k], atol_scale=2e-5, what=prefix + k + " (fused launch)")


def test_ranking_sequence_feature_encoder():
    g = load("ranking_layers_seq")
    fm = feature_map("ranking_layers_seq", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    emb.load_state_dict(_sub(g, "emb."))
    emb.to(DEV)
    X = _X(fm, g["batch"], False)
    E = emb(X)
    assert E.shape == g["E"].shape
    assert torch.equal(E[:, :6].cpu(), g["E"][:, :6])
    assert_close(E[:, 6], g["E"][:, 6], what="pooled sequence slot")
    (E * g["wE"].to(DEV)).sum().backward()
    want = _sub(g, "grad.emb.")
    for k, p in emb.named_parameters():
        assert_close(p.grad, want[k], atol_scale=2e-5, what=k)


def test_feature_source_and_type_selection():
    g = load("ranking_layers_d8")
    fm = feature_map("ranking_layers_d8", 8)
    for i, (k, s) in enumerate(fm.features.items()):
        s["source"] = "user" if i % 2 == 0 else "item"
    emb = layers.FeatureEmbedding(fm, 8)
    emb.load_state_dict(_sub(g, "emb."))
    emb.to(DEV)
    X = _X(fm, g["batch"], False)
    names = list(fm.features)
    sel = emb(X, feature_source="user")
    idx = [i for i, n in enumerate(names) if fm.features[n]["source"] == "user"]
    assert torch.equal(sel.cpu(), g["E"][:, idx])
    cat = emb(X, feature_type="categorical")
    idx = [i for i, n in enumerate(names) if fm.features[n]["type"] == "categorical"]
    assert torch.equal(cat.cpu(), g["E"][:, idx])
    flat = emb(X, dynamic_emb_dim=True)
    assert torch.equal(flat.cpu(), g["E"].flatten(1))


def test_core_embedding_layer_two_towers():
    g = load("core_layers")
    fmap = MatchingFeatureMap(feature_specs=OrderedDict([
        ("item_id", {"type": "categorical", "source": "item", "vocab_size": 23, "padding_idx": 22}),
        ("item_cat", {"type": "categorical", "source": "item", "vocab_size": 6}),
        ("user_id", {"type": "categorical", "source": "user", "vocab_size": 17}),
        ("user_age", {"type": "numeric", "source": "user"}),
        ("user_hist", {"type": "sequence", "source": "user", "vocab_size": 23, "padding_idx": 22,
                       "share_embedding": "item_id", "embedding_callback": "layers.MaskedAveragePooling()"}),
    ]))
    layer = layers.EmbeddingLayer(fmap, 8)
    layer.load_state_dict(_sub(g, "emb."))
    layer.to(DEV)
    X = {k[2:]: v.to(DEV) for k, v in g.items() if k.startswith("X.")}
    U = layer(X, feature_source="user")
    V = layer(X, feature_source="item")
    assert U.shape == g["U"].shape and V.shape == g["V"].shape
    assert torch.equal(V.cpu(), g["V"])
    assert torch.equal(U[:, :2].cpu(), g["U"][:, :2])
    assert_close(U[:, 2], g["U"][:, 2], what="pooled user_hist")
    ((U * g["wU"].to(DEV)).sum() + (V * g["wV"].to(DEV)).sum()).backward()
    want = _sub(g, "grad.emb.")
    for k, p in layer.named_parameters():
        assert_close(p.grad, want[k], atol_scale=2e-5, what=k)
    one = layers.EmbeddingLayer(fmap, 8, required_feature_columns=["user_id"])
    one.load_state_dict({"embedding_layer.embedding_layers.user_id.weight":
                         g["emb.embedding_layer.embedding_layers.user_id.weight"]})
    one.to(DEV)
    out = one(X)
    assert out.dim() == 2 and torch.equal(out.cpu(), g["single"])       # embedding.py:110-111


def test_two_tower_score_and_softmax_loss():
    g = load("two_tower")
    u = g["u"].to(DEV).requires_grad_(True)
    v = g["v"].to(DEV).requires_grad_(True)
    y = layers.two_tower_score(u, v)
    assert_close(y, g["y"], what="y")
    # SoftmaxCrossEntropyLoss (core/pytorch/losses/softmax_crossentropy_loss.py:14-21), positives in column 0
    loss = -torch.log_softmax(y, dim=1)[:, 0].mean()
    assert_close(loss, g["loss"], what="loss")
    loss.backward()
    assert_close(u.grad, g["du"], atol_scale=2e-5, what="du")
    assert_close(v.grad, g["dv"], atol_scale=2e-5, what="dv")
    yd = layers.two_tower_score(g["dssm_u"].to(DEV), g["dssm_v"].to(DEV))
    assert_close(yd.view(-1), g["dssm_y"], what="dssm")


def test_pooling_modules_on_materialised_tensor():
    g = load("pooling")
    emb = g["emb"].to(DEV).requires_grad_(True)
    assert_close(layers.MaskedAveragePooling()(emb), g["ranking_avg"], what="avg")
    assert_close(layers.MaskedAveragePooling()(emb, g["mask"].to(DEV)), g["ranking_avg_mask"], what="avg mask")
    assert_close(layers.MaskedSumPooling()(emb), g["ranking_sum"], what="sum")
    assert_close(layers.CoreMaskedAveragePooling()(emb), g["core_avg"], what="core avg")
    ref = g["emb"].clone().requires_grad_(True)
    w = torch.randn(6, 8, generator=torch.Generator().manual_seed(0))
    m = (ref.sum(-1) != 0)
    ((ref.sum(1) / (m.float().sum(-1, keepdim=True) + 1e-12)) * w).sum().backward()
    (layers.MaskedAveragePooling()(emb) * w.to(DEV)).sum().backward()
    assert_close(emb.grad, ref.grad, what="pool grad")


def _MLP(input_dim, hidden):
    """MLP_Block(hidden_units, ReLU, output_dim=1) of blocks/mlp_block.py:23-61 (parameter names mlp.0, mlp.2, ...): the fused
    tcgen05 GEMM chain of csrc/gemm.cu behind the reference's constructor."""
    return layers.MLP_Block(input_dim=input_dim, hidden_units=list(hidden), hidden_activations="ReLU", output_dim=1)


class _DeepFM(nn.Module):
    def __init__(self, fm, D, hidden, use_mlp=True):
        super().__init__()
        self.feature_map = fm
        self.embedding_layer = layers.FeatureEmbedding(fm, D)
        self.fm_layer = layers.FactorizationMachine(fm)
        self.mlp = _MLP(fm.sum_emb_out_dim(), hidden) if use_mlp else None

    def forward(self, batch):
        X = layers.PackedInputs(self.feature_map, batch)
        E = self.embedding_layer(X)
        y = self.fm_layer(X, E)
        if self.mlp is not None:
            y = y + self.mlp(E.flatten(start_dim=1))
        return torch.sigmoid(y)


def _train_step(model, opt, batch):
    """RankingModel.train_step (ranking_model.py:191-197): zero_grad, BCE(mean), backward,
    clip_grad_norm_(all params, 10), Adam step."""
    y_true = batch[:, -1].float().view(-1, 1)
    opt.zero_grad()
    loss = torch.nn.functional.binary_cross_entropy(model(batch), y_true, reduction="mean")
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
    opt.step()
    return float(loss)


def test_deepfm_three_train_steps_match_reference():
    g = load("deepfm_train")
    fm = feature_map("ranking_layers_d8", 8)
    model = _DeepFM(fm, 8, (16, 8))
    init = _sub(g, "init.")
    assert sorted(model.state_dict()) == sorted(init)
    model.load_state_dict(init)
    model.to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = [_train_step(model, opt, g["batch%d" % s].to(DEV)) for s in range(3)]
    assert_close(torch.tensor(losses), g["losses"].float(), rtol=2e-6, atol_scale=0, what="losses")
    final = _sub(g, "final.")
    for k, v in model.state_dict().items():
        assert_close(v, final[k], rtol=1e-4, atol_scale=1e-5, what=k)
    model.eval()
    assert_close(model(g["batch0"].to(DEV)), g["pred_final"], rtol=1e-5, what="pred_final")


def test_config1_fm_on_the_reference_preprocessed_csv():
    """BASELINE configs[0]: FM (D = 10 -> scalar kernel path) on the 1k-row Criteo-shaped CSV that
    the reference's own FeatureProcessor tokenised (tests/golden/config1_fm.npz)."""
    g = load("config1_fm")
    from recbox_b200.features import FeatureMap
    fm = FeatureMap("criteo_1k", ".")
    for i in range(1, 14):
        fm.add_numeric("I%d" % i)
    for i, V in enumerate(g["vocab_sizes"].tolist(), 1):
        fm.add_categorical("C%d" % i, int(V), padding_idx=0)
    fm.finalize(["label"])
    fm.default_emb_dim = 10
    model = _DeepFM(fm, 10, (), use_mlp=False)
    init = _sub(g, "init.")
    assert sorted(model.state_dict()) == sorted(init)
    model.load_state_dict(init)
    model.to(DEV)
    batch = g["batch"].to(DEV)
    model.eval()
    assert_close(model(batch[:128]), g["pred_init"], rtol=1e-6, what="pred_init")
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = [_train_step(model, opt, batch[s * 128:(s + 1) * 128]) for s in range(4)]
    assert_close(torch.tensor(losses), g["losses"].float(), rtol=2e-6, atol_scale=0, what="losses")
    model.eval()
    assert_close(model(batch[:128]), g["pred_final"], rtol=1e-5, what="pred_final")
    # Adam divides by sqrt(v): where a gradient entry is ~0 (|g| ~ 1e-9 at std-1e-4 init) its rounding
    # noise is amplified to a fraction of lr per step, so weights carry an absolute floor of
    # ~1e-4 * max|w| after 4 steps; the logits above are the 1e-5 contract.
    final = _sub(g, "final.")
    for k, v in model.state_dict().items():
        assert_close(v, final[k], rtol=1e-4, atol_scale=2e-4, what=k)


def test_module_to_and_deepcopy_keep_the_fused_storage():
    import copy
    g = load("ranking_layers_d8")
    fm = feature_map("ranking_layers_d8", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    emb.load_state_dict(_sub(g, "emb."))
    emb2 = copy.deepcopy(emb).to(DEV)
    emb.to(DEV)
    X = _X(fm, g["batch"], True)
    assert torch.equal(emb(X).cpu(), g["E"]) and torch.equal(emb2(X).cpu(), g["E"])
    with torch.no_grad():
        emb2.embedding_layer.embedding_layers["C1"].weight.mul_(2.0)
    assert torch.equal(emb(X).cpu(), g["E"])                      # the copy owns its own table
    assert not torch.equal(emb2(X).cpu(), g["E"])


@pytest.mark.parametrize("bn", [False, True])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_interaction_machine_matches_reference(order, bn):
    """f4: layers.InteractionMachine (one-pass power sums) against the reference module's forward / backward."""
    g = load("interaction_machine")
    tag = "o%d_bn%d." % (order, int(bn))
    m = layers.InteractionMachine(8, order=order, batch_norm=bn)
    m.load_state_dict(_sub(g, tag + "sd."))
    m.to(DEV).train()
    X = g[tag + "X"].to(DEV).requires_grad_(True)
    y = m(X)
    assert_close(y, g[tag + "y"], atol_scale=2e-5, what="y")
    (y * g[tag + "w"].to(DEV)).sum().backward()
    assert_close(X.grad, g[tag + "dX"], atol_scale=5e-5, what="dX")
    for k, p_ in m.named_parameters():
        assert_close(p_.grad, g[tag + "grad." + k], atol_scale=5e-5, what=k)


def test_sasrec_lookups_and_token_dots_match_rechub():
    """a11 against the golden minted from rechub's SASRec (sasrec.py:98-107): the three shared-table lookups are one
    rbx_gather_rows call (bit-exact), the per-token logits one rbx_rowdot call each."""
    from recbox_b200 import ops
    g = load("sasrec_gather")
    table = g["table"].to(DEV)
    ids = torch.stack([g["seq"], g["pos"], g["neg"]], 1).to(torch.int32).to(DEV)          # [B,3,L]
    emb = ops.gather_rows(table, ids)
    assert torch.equal(emb.cpu(), g["emb"])
    B, _, L, D = emb.shape
    so = g["seq_out"].to(DEV).reshape(B * L, D).contiguous()
    for k, name in ((1, "pos_logits"), (2, "neg_logits")):
        y = ops.rowdot_fwd(so, emb[:, k].reshape(B * L, 1, D).contiguous())
        assert_close(y.view(B, L), g[name], atol_scale=2e-5, what=name)


@pytest.mark.parametrize("tag", ["plain", "mixed", "nohead", "core"])
def test_mlp_block_matches_reference(tag):
    """a13: our MLP_Block / MLP_Layer (Linear layers on the tcgen05 GEMM chain, bias + ReLU + ReLU-mask fused) against the
    outputs and gradients of the reference's own modules on the same seeded init and input (tests/golden/mlp_block.npz)."""
    g = load("mlp_block")
    mk = {"plain": lambda: layers.MLP_Block(input_dim=104, hidden_units=[96, 48], hidden_activations="ReLU", output_dim=1),
          "mixed": lambda: layers.MLP_Block(input_dim=104, hidden_units=[64, 32], hidden_activations=["relu", "tanh"], output_dim=1,
                                            output_activation="sigmoid", batch_norm=True, use_bias=False),
          "nohead": lambda: layers.MLP_Block(input_dim=104, hidden_units=[48], hidden_activations="ReLU"),
          "core": lambda: layers.MLP_Layer(input_dim=104, output_dim=1, hidden_units=[32, 16], hidden_activations="ReLU",
                                           final_activation=None, dropout_rates=[0.0, 0.0])}[tag]
    torch.manual_seed(5)
    m = mk()
    init = _sub(g, tag + ".init.")
    assert sorted(m.state_dict()) == sorted(init)
    for k, v in m.state_dict().items():          # same constructor, same seed -> the reference's own initial weights
        assert torch.equal(v, init[k]), k
    m.to(DEV).train()
    x = g["x"].to(DEV).requires_grad_(True)
    y = m(x)
    assert_close(y, g[tag + ".y"], rtol=1e-5, atol_scale=1e-5, what="y")
    (y * g[tag + ".w"].to(DEV)).sum().backward()
    assert_close(x.grad, g[tag + ".dx"], rtol=1e-5, atol_scale=1e-5, what="dx")
    for k, p in m.named_parameters():
        assert_close(p.grad, g["%s.grad.%s" % (tag, k)], rtol=1e-5, atol_scale=1e-5, what=k)
