"""The packed input pipeline (recbox_b200/loader.py, f2) feeding the fused layers on cuda:0: a PackedBatch gives
bit-identical embeddings / logits / gradients to the float64 batch matrix of the reference's loader, both with
local ids (one pack launch) and with the loader bound to the layer (ids block = the kernels' rows, no launch)."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from helpers import assert_close
from test_layers_host import feature_map
from test_loader_host import _data
from test_oracle_golden import load

from recbox_b200 import layers
from recbox_b200.loader import PackedDataLoader, PackedDataset

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _M(object):
    def __init__(self, fm):
        self.feature_map, self.device = fm, torch.device(DEV)


def _run(emb, fml, X, wE, wy):
    for p in list(emb.parameters()) + list(fml.parameters()):
        p.grad = None
    E = emb(X)
    y = fml(X, E)
    ((E * wE).sum() + (y * wy).sum()).backward()
    grads = OrderedDict((k, p.grad.clone()) for k, p in list(emb.named_parameters()) + list(fml.named_parameters()))
    return E.detach().clone(), y.detach().clone(), grads


@pytest.mark.parametrize("compact", [False, "auto"])
@pytest.mark.parametrize("bind", [False, True])
@pytest.mark.parametrize("tag,D", [("ranking_layers_d8", 8), ("ranking_layers_d10", 10)])
def test_packed_batch_equals_float64_batch(tag, D, bind, compact):
    g = load(tag)
    fm = feature_map(tag, D)
    emb = layers.FeatureEmbedding(fm, D)
    fml = layers.FactorizationMachine(fm)
    emb.load_state_dict(OrderedDict((k[4:], v) for k, v in g.items() if k.startswith("emb.")))
    fml.load_state_dict(OrderedDict((k[3:], v) for k, v in g.items() if k.startswith("fm.")))
    emb.to(DEV)
    fml.to(DEV)
    model = _M(fm)
    wE, wy = g["wE"].to(DEV), g["wy"].to(DEV)
    ref = _run(emb, fml, layers.get_inputs(model, g["batch"]), wE, wy)
    assert torch.equal(ref[0].cpu(), g["E"])
    dl = PackedDataLoader(fm, PackedDataset(fm, g["batch"], compact=compact), batch_size=len(g["batch"]))
    assert dl.dataset.compact == (compact == "auto")         # every vocabulary of the golden config is below 65 536
    if bind:
        dl.bind(emb)
    (pb,) = list(dl)
    X = layers.get_inputs(model, pb)
    got = _run(emb, fml, X, wE, wy)
    assert torch.equal(got[0], ref[0]), "E"
    # the second call computes the first-order term inside the embedding launch (other summation order)
    assert_close(got[1], ref[1], rtol=1e-6, atol_scale=1e-6, what="y")
    for k in ref[2]:
        assert_close(got[2][k], ref[2][k], rtol=1e-6, atol_scale=1e-6, what=k)     # atomics reorder sums
    assert torch.equal(layers.get_labels(model, pb).cpu(), g["batch"][:, -1].float().view(-1, 1))
    assert torch.equal(layers.get_labels(model, g["batch"]).cpu(), g["batch"][:, -1].float().view(-1, 1))


@pytest.mark.parametrize("compact", [False, "auto"])
def test_device_prefetch_epoch_shuffled(compact):
    fm = feature_map("ranking_layers_d8", 8)
    arr = _data(5000, fm, 5)
    arr[:, -1] = np.arange(5000)
    dl = PackedDataLoader(fm, PackedDataset(fm, arr, compact=compact), batch_size=512, shuffle=True, seed=1, device=DEV)
    ds = dl.dataset
    labs = []
    for b in dl:
        got, want = (b.ids16, ds.ids16) if ds.compact else (b.ids, ds.ids)
        assert got.is_cuda and b.dense.is_cuda and b.labels.is_cuda
        lab = b.labels.long().cpu()
        assert torch.equal(got.cpu(), want[lab]) and torch.equal(b.dense.cpu(), ds.dense[lab])
        labs.append(lab)
    assert sorted(torch.cat(labs).tolist()) == list(range(5000))


@pytest.mark.parametrize("compact", [False, "auto"])
def test_feature_source_subset_with_bound_offsets(compact):
    """A bound loader's ids carry the main layer's row offsets; a layer that selects a subset of the features (or
    the D=1 LR table with other offsets) must still read the right rows."""
    g = load("ranking_layers_d8")
    fm = feature_map("ranking_layers_d8", 8)
    emb = layers.FeatureEmbedding(fm, 8)
    fml = layers.FactorizationMachine(fm)
    emb.load_state_dict(OrderedDict((k[4:], v) for k, v in g.items() if k.startswith("emb.")))
    fml.load_state_dict(OrderedDict((k[3:], v) for k, v in g.items() if k.startswith("fm.")))
    emb.to(DEV)
    fml.to(DEV)
    model = _M(fm)
    (pb,) = list(PackedDataLoader(fm, PackedDataset(fm, g["batch"], compact=compact), batch_size=len(g["batch"])).bind(emb))
    X = layers.get_inputs(model, pb)
    assert_close(fml.lr_layer(X), g["lr_out"], what="lr_out (own offsets)")
    sub = emb(X, feature_type="categorical")
    full = emb(layers.get_inputs(model, g["batch"]))
    cat_pos = [i for i, (n, s) in enumerate(fm.features.items()) if s["type"] == "categorical"]
    assert torch.equal(sub, full[:, cat_pos, :])
