"""Host-side logic of the layer API (no GPU): construction, parameter naming, seeded init,
fused-storage aliasing, install() rebinding, and the "no CPU path" contract."""
from collections import OrderedDict

import numpy as np
import pytest
import torch
from torch import nn

from test_oracle_golden import load, ranking_features

from recbox_b200 import RbxError, layers
from recbox_b200.features import FeatureMap, MatchingFeatureMap


def feature_map(tag, D):
    fm = FeatureMap("golden", ".")
    for k, v in ranking_features(tag).items():
        fm.features[k] = dict(v)
    fm.finalize(["label"])
    fm.default_emb_dim = D
    return fm


def test_state_dict_keys_and_seeded_init_match_reference():
    """Same seed -> same initial weights, same parameter names (golden minted from the reference:
    oracle/make_golden.py golden_init)."""
    g = load("init_seed2024")
    torch.manual_seed(2024)
    fm = feature_map("ranking_layers_share", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    fml = layers.FactorizationMachine(fm)
    got = OrderedDict(("emb." + k, v) for k, v in emb.state_dict().items())
    got.update(("fm." + k, v) for k, v in fml.state_dict().items())
    assert sorted(got) == sorted(g)
    for k, v in got.items():
        assert torch.equal(v, g[k]), k


def test_per_feature_parameters_alias_one_fused_table():
    fm = feature_map("ranking_layers_d8", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    d = emb.embedding_layer
    grp = d._store.groups[8]
    assert grp.R == 11 + 7 + 13 + 11 and grp.table.shape == (grp.R, 8) and grp.dense_w.shape == (2, 8)
    for name in ("C1", "C2", "C3", "C4"):
        m = d.embedding_layers[name]
        assert type(m) == nn.Embedding                       # match_model.py:92-103 scans for exactly this
        off = grp.emb_off[id(m)]
        assert m.weight.data_ptr() == grp.table[off].data_ptr()
        assert float(m.weight[0].abs().sum()) == 0.0          # padding row stays zero
    # writes through the per-feature parameter land in the fused table (checkpoint load path)
    with torch.no_grad():
        d.embedding_layers["C2"].weight.fill_(3.0)
    o = grp.emb_off[id(d.embedding_layers["C2"])]
    assert float(grp.table[o:o + 7].min()) == 3.0
    # load_state_dict keeps the aliasing
    sd = {k: torch.full_like(v, 2.0) for k, v in emb.state_dict().items()}
    emb.load_state_dict(sd)
    d._store.ensure()
    assert float(d._store.groups[8].table.min()) == 2.0 and float(d._store.groups[8].dense_w.max()) == 2.0


def test_shared_embedding_registers_one_module_under_two_names():
    fm = feature_map("ranking_layers_share", 16)
    d = layers.FeatureEmbedding(fm, 16).embedding_layer
    assert d.embedding_layers["C4"] is d.embedding_layers["C1"]
    assert d._store.groups[16].R == 11 + 7 + 13                 # shared table stored once


def test_lr_layer_uses_dim_one_and_sum_pooling_for_sequences():
    fm = feature_map("ranking_layers_seq", 8)
    lr = layers.LogisticRegression(fm)
    d = lr.embedding_layer.embedding_layer
    assert d.embedding_layers["C1"].embedding_dim == 1 and d.embedding_layers["I1"].out_features == 1
    assert type(d.feature_encoders["S1"]) is layers.MaskedSumPooling      # feature_embedding.py:66-68
    emb = layers.FeatureEmbedding(fm, 8).embedding_layer
    assert type(emb.feature_encoders["S1"]) is layers.MaskedAveragePooling  # from the "layers.X()" spec string


def test_unknown_modes_raise_like_the_reference():
    with pytest.raises(ValueError):
        layers.InnerProductInteraction(5, output="nope")
    fm = feature_map("ranking_layers_d8", 8)
    fm.features["C1"]["feature_encoder"] = "layers.DoesNotExist()"
    with pytest.raises(ValueError):
        layers.FeatureEmbedding(fm, 8)


def test_no_cpu_path():
    fm = feature_map("ranking_layers_d8", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    X = {k: torch.zeros(4, dtype=torch.float64) for k in fm.features}
    with pytest.raises(RbxError):
        emb(X)
    with pytest.raises(RbxError):
        layers.InnerProductInteraction(3, "bi_interaction")(torch.zeros(2, 3, 4))
    with pytest.raises(RbxError):
        layers.two_tower_score(torch.zeros(2, 4), torch.zeros(2, 3, 4))
    with pytest.raises(RbxError):
        layers.MaskedAveragePooling()(torch.zeros(2, 3, 4))


def test_core_layer_construction():
    fmap = MatchingFeatureMap(feature_specs=OrderedDict([
        ("item_id", {"type": "categorical", "source": "item", "vocab_size": 23, "padding_idx": 22}),
        ("user_id", {"type": "categorical", "source": "user", "vocab_size": 17}),
        ("user_age", {"type": "numeric", "source": "user"}),
        ("user_hist", {"type": "sequence", "source": "user", "vocab_size": 23, "padding_idx": 22,
                       "share_embedding": "item_id", "embedding_callback": "layers.MaskedAveragePooling()"}),
    ]))
    layer = layers.EmbeddingLayer(fmap, 8)
    d = layer.embedding_layer
    assert d.embedding_layers["user_hist"] is d.embedding_layers["item_id"]
    assert type(d.embedding_callbacks["user_hist"]) is layers.CoreMaskedAveragePooling
    assert sorted(layer.state_dict()) == sorted("embedding_layer.embedding_layers.%s.weight" % n
                                                for n in ("item_id", "user_id", "user_age", "user_hist"))


@pytest.mark.reference
def test_install_rebinds_reference_symbols():
    """Unmodified reference model code picks the fused modules up after recbox_b200.install()."""
    import tempfile

    import recbox_b200
    from oracle import ref_shim
    L = ref_shim.install()
    ref_cls = L.FeatureEmbedding
    try:
        done = recbox_b200.install()
        assert ("recbox.ranking.pytorch.layers", "FeatureEmbedding") in done
        assert L.FeatureEmbedding is layers.FeatureEmbedding and L.FactorizationMachine is layers.FactorizationMachine
        assert L.InteractionMachine is layers.InteractionMachine
        import fuxictr.pytorch.layers as FL
        assert FL.LogisticRegression is layers.LogisticRegression
        import recbox.core.pytorch.layers as CL
        assert CL.EmbeddingLayer is layers.EmbeddingLayer
        from recbox.ranking.features import FeatureMap as RefFeatureMap
        from recbox.ranking.pytorch.models.ranking_model import RankingModel
        tmp = tempfile.mkdtemp()
        fm = RefFeatureMap("golden", tmp)
        for k, v in ranking_features("ranking_layers_d8").items():
            fm.features[k] = dict(v)
        fm.labels = ["label"]
        fm.num_fields = fm.get_num_fields()
        fm.set_column_index()
        fm.default_emb_dim = 8

        class DeepFM(RankingModel):           # FuxiCTR-style model body, verbatim against the reference API
            def __init__(self, feature_map, **kw):
                super(DeepFM, self).__init__(feature_map, **kw)
                self.embedding_layer = L.FeatureEmbedding(feature_map, 8)
                self.fm_layer = L.FactorizationMachine(feature_map)
                self.mlp = L.MLP_Block(input_dim=feature_map.sum_emb_out_dim(), output_dim=1, hidden_units=[16, 8])
                self.compile("adam", "binary_cross_entropy", 1e-3)
                self.reset_parameters()
                self.model_to_device()
        torch.manual_seed(5)
        m = DeepFM(fm, model_id="m", gpu=-1, verbose=0, model_root=tmp, metrics=["AUC"])
        assert type(m.embedding_layer) is layers.FeatureEmbedding
        g = load("deepfm_train")
        init = {k[5:]: v for k, v in g.items() if k.startswith("init.")}
        assert sorted(m.state_dict()) == sorted(init)
        # reference's reset_parameters (xavier on the Linear(1,D), ranking_model.py:93-104) ran through our modules
        assert all(m.state_dict()[k].shape == v.shape for k, v in init.items())
    finally:
        recbox_b200.uninstall()
    assert L.FeatureEmbedding is ref_cls


def test_interaction_machine_parameter_names_match_reference():
    """state_dict keys of layers.InteractionMachine equal the reference module's (golden minted from it), and the CPU
    call is refused (no CPU path)."""
    g = load("interaction_machine")
    for order, bn in ((2, 0), (5, 1)):
        tag = "o%d_bn%d.sd." % (order, bn)
        want = sorted(k[len(tag):] for k in g if k.startswith(tag))
        m = layers.InteractionMachine(8, order=order, batch_norm=bool(bn))
        assert sorted(m.state_dict()) == want
        with pytest.raises(RbxError):
            m(torch.zeros(2, 3, 8))
    with pytest.raises(AssertionError):
        layers.InteractionMachine(8, order=6)


def test_embdict_drops_its_cached_stack_when_mutated():
    """dict2tensor stacks the CURRENT dict values (feature_embedding.py:169-186): a model that replaces an entry after the
    fused forward (DIN-style `feature_emb_dict[seq_field] = pooled`) must not get the stale cached [B,F,D] back."""
    E = torch.arange(24, dtype=torch.float32).view(2, 3, 4)
    d = layers._EmbDict()
    d.names = ("a", "b", "c")
    for i, n in enumerate(d.names):
        d[n] = E[:, i, :]
    d.stacked = E                                   # what the fused forward does last
    assert d.stacked is E
    d["b"] = torch.zeros(2, 4)                      # replace an entry -> cache dropped
    assert d.stacked is None
    d2 = layers._EmbDict()
    d2.names = ("a", "b")
    d2["a"], d2["b"] = E[:, 0, :], E[:, 1, :]
    d2.stacked = E
    d2.pop("a")
    assert d2.stacked is None
    d3 = layers._EmbDict()
    d3["a"] = E[:, 0, :]
    d3.stacked = E
    E[:, 0, :] += 1                                 # in-place write through a view bumps the version -> cache dropped
    assert d3.stacked is None
    st = layers._Stash()
    E2 = torch.zeros(2, 3, 4)
    st.version = E2._version
    E2._rbx_stash = st
    assert layers._stash_of(E2) is st
    E2.mul_(2)
    assert layers._stash_of(E2) is None             # the launch's by-products no longer describe E


def test_refused_store_invalidates_cached_plans_and_survives_pickle():
    import copy
    import pickle
    fm = feature_map("ranking_layers_d8", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    d = emb.embedding_layer
    key, names = d._select([], [])
    plan = d._plan(key, names)
    assert d._calls and plan["call"] is not None
    old_param = d.embedding_layers[names[-1]].weight
    d.embedding_layers[names[-1]].weight = nn.Parameter(torch.ones_like(old_param))      # pretrained weights loaded late
    d._sync_store()
    assert not d._calls, "plans made before the re-fusion hold the orphaned Parameter"
    call, _ = d._plan(key, names)["call"]
    assert all(p is not old_param for p in call.params)
    assert any(p is d.embedding_layers[names[-1]].weight for p in call.params)
    # pickling rebuilds the store (its offsets are keyed by id(module))
    emb2 = pickle.loads(pickle.dumps(emb))
    d2 = emb2.embedding_layer
    for n in names:
        assert torch.equal(d2.embedding_layers[n].weight, d.embedding_layers[n].weight)
    g = next(iter(d2._store.groups.values()))
    assert all(id(m) in g.emb_off for m in g.emb)
    assert copy.deepcopy(emb).state_dict().keys() == emb.state_dict().keys()


def test_mlp_block_constructor_matches_reference_init():
    """a13 host side: same constructor arguments, same `mlp` Sequential (state_dict keys), same RNG call order as the
    reference's MLP_Block / MLP_Layer (golden minted from the reference under torch.manual_seed(5))."""
    from test_oracle_golden import load
    g = load("mlp_block")
    cases = {"plain": lambda: layers.MLP_Block(input_dim=104, hidden_units=[96, 48], hidden_activations="ReLU", output_dim=1),
             "mixed": lambda: layers.MLP_Block(input_dim=104, hidden_units=[64, 32], hidden_activations=["relu", "tanh"], output_dim=1,
                                               output_activation="sigmoid", batch_norm=True, use_bias=False),
             "core": lambda: layers.MLP_Layer(input_dim=104, output_dim=1, hidden_units=[32, 16], hidden_activations="ReLU",
                                              final_activation=None, dropout_rates=[0.0, 0.0])}
    for tag, mk in cases.items():
        torch.manual_seed(5)
        m = mk()
        want = {k[len(tag) + 6:]: v for k, v in g.items() if k.startswith(tag + ".init.")}
        assert sorted(m.state_dict()) == sorted(want)
        for k, v in m.state_dict().items():
            assert torch.equal(v, want[k]), (tag, k)
    with pytest.raises(layers.RbxError):              # no CPU path for the Linear layers either
        cases["plain"]()(torch.zeros(2, 104))


def test_dense_tail_ops_have_no_cpu_path():
    """a13 / f4 host side: the GEMM wrappers and the blocks built on them refuse CPU tensors loudly."""
    from recbox_b200 import blocks, ops
    with pytest.raises(layers.RbxError):
        ops.gemm(torch.zeros(4, 8), torch.zeros(2, 8))
    with pytest.raises(layers.RbxError):
        ops.colsum(torch.zeros(4, 8))
    with pytest.raises(layers.RbxError):
        blocks.linear(torch.zeros(4, 8), torch.zeros(2, 8))
    with pytest.raises(layers.RbxError):
        blocks.CrossNetV2(8, 2)(torch.zeros(4, 8))
    m = blocks.CompressedInteractionNet(3, [4, 2])          # constructor parity: the reference's parameter names
    assert sorted(m.state_dict()) == ["cin_layer.layer_1.bias", "cin_layer.layer_1.weight", "cin_layer.layer_2.bias",
                                      "cin_layer.layer_2.weight", "fc.bias", "fc.weight"]
    assert tuple(m.cin_layer["layer_1"].weight.shape) == (4, 9, 1) and tuple(m.cin_layer["layer_2"].weight.shape) == (2, 12, 1)
