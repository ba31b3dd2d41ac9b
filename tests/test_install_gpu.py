"""recbox_b200.install() on the GPU: a model written against the REFERENCE's API -- a `RankingModel` subclass (the
reference's own base class: get_inputs / get_labels / add_loss / train_step, ranking_model.py:30-197, imported unmodified
from the git-ignored copy baseline/_ref/ that baseline/fetch_ref.py makes and that ships with the snapshot) whose layers
are looked up as `recbox.ranking.pytorch.layers.*` -- trains on cuda:0 through the fused modules after the rebinding and
walks the trajectory the golden recorded from the all-reference run of the same class (tests/golden/deepfm_train.npz:
losses, first-step gradients, final weights, predictions)."""
import os
import tempfile

import pytest
import torch

from helpers import ROOT, assert_close
from test_oracle_golden import load, ranking_features

REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "recbox", "ranking", "pytorch")),
                                 reason="baseline/_ref (python baseline/fetch_ref.py) not present")]


def test_unmodified_ranking_model_trains_on_the_fused_modules():
    import recbox_b200
    from recbox_b200 import layers
    from oracle import ref_shim
    if not ref_shim.available():                      # the GPU box: no /root/reference, the shipped copy instead
        ref_shim.REFERENCE_ROOT = REF
    L = ref_shim.install()
    ref_cls = L.FeatureEmbedding
    try:
        done = recbox_b200.install()
        assert ("recbox.ranking.pytorch.layers", "FeatureEmbedding") in done
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

        class DeepFM(RankingModel):                   # the class oracle/make_golden.py trained all-reference, verbatim
            def __init__(self, feature_map, **kw):
                super(DeepFM, self).__init__(feature_map, **kw)
                self.embedding_layer = L.FeatureEmbedding(feature_map, 8)
                self.fm_layer = L.FactorizationMachine(feature_map)
                self.mlp = L.MLP_Block(input_dim=feature_map.sum_emb_out_dim(), output_dim=1, hidden_units=[16, 8])
                self.compile("adam", "binary_cross_entropy", 1e-3)
                self.reset_parameters()
                self.model_to_device()

            def forward(self, inputs):
                X = self.get_inputs(inputs)
                E = self.embedding_layer(X)
                y = self.fm_layer(X, E)
                y = y + self.mlp(E.flatten(start_dim=1))
                return {"y_pred": self.output_activation(y)}

        m = DeepFM(fm, model_id="m", gpu=0, verbose=0, model_root=tmp, metrics=["AUC"])
        m._max_gradient_norm = 10.
        assert type(m.embedding_layer) is layers.FeatureEmbedding and type(m.mlp) is layers.MLP_Block
        assert next(m.parameters()).is_cuda
        g = load("deepfm_train")
        init = {k[5:]: v for k, v in g.items() if k.startswith("init.")}
        assert sorted(m.state_dict()) == sorted(init)
        m.load_state_dict(init)
        losses = []
        for s in range(3):
            m.train()
            losses.append(float(m.train_step(g["batch%d" % s])))      # host float64 batch, as the reference's loader yields
            if s == 0:
                for k, p in m.named_parameters():
                    assert_close(p.grad, g["grad0." + k], atol_scale=2e-5, what="grad0." + k)
        assert_close(torch.tensor(losses), g["losses"].float(), rtol=2e-6, atol_scale=0, what="losses")
        for k, v in m.state_dict().items():
            assert_close(v, g["final." + k], rtol=1e-4, atol_scale=1e-5, what="final." + k)
        m.eval()
        with torch.no_grad():
            assert_close(m.forward(g["batch0"])["y_pred"], g["pred_final"], rtol=1e-5, what="pred_final")
    finally:
        recbox_b200.uninstall()
    assert L.FeatureEmbedding is ref_cls
