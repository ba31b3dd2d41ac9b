"""recbox_b200.graphs.GraphedStep: a whole train step (packed batch -> FeatureEmbedding + FM + MLP_Block -> BCE -> backward ->
clip -> Adam) replayed from a CUDA graph must walk the same parameter trajectory as the same step issued eagerly."""
import copy

import numpy as np
import pytest
import torch
from torch import nn

from helpers import assert_close
from test_layers_host import feature_map

from recbox_b200 import graphs, layers
from recbox_b200.loader import PackedDataset

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _DeepFM(nn.Module):
    def __init__(self, fm, D):
        super().__init__()
        self.feature_map, self.device = fm, torch.device(DEV)
        self.embedding_layer = layers.FeatureEmbedding(fm, D)
        self.fm_layer = layers.FactorizationMachine(fm)
        self.mlp = layers.MLP_Block(input_dim=fm.sum_emb_out_dim(), hidden_units=[32, 16], hidden_activations="ReLU", output_dim=1)

    def forward(self, batch):
        X = layers.get_inputs(self, batch)
        E = self.embedding_layer(X)
        return torch.sigmoid(self.fm_layer(X, E) + self.mlp(E.flatten(start_dim=1)))


def _batches(fm, n, B, seed):
    rng = np.random.default_rng(seed)
    cols = []
    for name, spec in fm.features.items():
        cols.append(rng.random((n * B, 1)) if spec["type"] == "numeric" else rng.integers(0, spec["vocab_size"], (n * B, 1)).astype(np.float64))
    cols.append((rng.random((n * B, 1)) < 0.5).astype(np.float64))
    ds = PackedDataset(fm, np.concatenate(cols, 1))
    return [ds.batch(i * B, (i + 1) * B) for i in range(n)]


def test_graphed_train_step_walks_the_eager_trajectory():
    D, B, steps = 8, 256, 4
    fm = feature_map("ranking_layers_d8", D)
    torch.manual_seed(3)
    eager = _DeepFM(fm, D).to(DEV)
    graphed = copy.deepcopy(eager)
    feed = _batches(fm, steps, B, 5)

    def make(model, capturable):
        # plain SGD: the update is linear in the gradient, so the trajectories may be compared at 1e-4 (Adam divides by
        # sqrt(v) and turns the round-off of the atomically accumulated gradients into O(lr) differences)
        opt = torch.optim.SGD(model.parameters(), lr=0.05)

        def train(batch):
            opt.zero_grad()
            loss = torch.nn.functional.binary_cross_entropy(model(batch), layers.get_labels(model, batch), reduction="mean")
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
            opt.step()
            return loss
        return train, opt

    train_e, _ = make(eager, False)
    losses_e = [float(train_e(b.to(DEV))) for b in feed]

    train_g, opt_g = make(graphed, True)
    static = feed[0].to(DEV)
    state = copy.deepcopy(graphed.state_dict())
    gs = graphs.GraphedStep(lambda: train_g(static), warmup=2)
    # the warm-up steps moved the parameters and Adam's state: rewind both IN PLACE (the graph holds these tensors)
    graphed.load_state_dict(state)
    assert not opt_g.state                                   # SGD without momentum keeps no state to rewind
    losses_g = []
    for b in feed:
        b.copy_into(static)
        losses_g.append(float(gs()))
    assert_close(torch.tensor(losses_g), torch.tensor(losses_e), rtol=1e-5, atol_scale=0, what="losses")
    for (k, a), (_, b) in zip(graphed.state_dict().items(), eager.state_dict().items()):
        assert_close(a, b, rtol=1e-4, atol_scale=1e-5, what=k)
