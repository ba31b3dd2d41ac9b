"""Host logic of the packed input pipeline (recbox_b200/loader.py): the one-time
conversion equals the reference's per-batch casts (ranking_model.py:106-116 slices + feature_embedding.py:201,204
`.long()` / `.float()`, restated by oracle.get_inputs), the loader surface mirrors h5_dataloader.py:50-59, and
a shuffled epoch visits every sample once.  CPU only."""
import numpy as np
import pytest
import torch

from helpers import oracle
from test_layers_host import feature_map
from test_oracle_golden import load, ranking_features

from recbox_b200 import RbxError
from recbox_b200.loader import PackedBatch, PackedColumns, PackedDataLoader, PackedDataset


def _data(N, fm, seed=0):
    rng = np.random.default_rng(seed)
    cols = []
    for n, s in fm.features.items():
        if s["type"] == "numeric":
            cols.append(rng.lognormal(0, 1, N).round(4))
        else:
            cols.append(rng.integers(0, s["vocab_size"], N).astype(np.float64))
    cols.append((rng.random(N) < 0.5).astype(np.float64))
    return np.stack(cols, 1)


def test_conversion_matches_reference_casts():
    fm = feature_map("ranking_layers_d8", 8)
    arr = _data(1000, fm)
    ds = PackedDataset(fm, arr, pin=False, compact=False)
    X = oracle.get_inputs(torch.from_numpy(arr), ranking_features("ranking_layers_d8"), ["label"])
    for i, n in enumerate(ds.cat_names):
        assert torch.equal(ds.ids[:, i].long(), X[n].long()), n
    for i, n in enumerate(ds.num_names):
        assert torch.equal(ds.dense[:, i], X[n].float()), n
    assert torch.equal(ds.labels, torch.from_numpy(arr[:, -1]).float())
    assert ds.bytes_per_sample == 4 * (4 + 2 + 1) and len(ds) == 1000
    # dict-of-columns form (load_h5) gives the same blocks
    names = list(fm.features.keys()) + ["label"]
    ds2 = PackedDataset(fm, {n: arr[:, i] for i, n in enumerate(names)}, pin=False, compact=False)
    assert torch.equal(ds2.ids, ds.ids) and torch.equal(ds2.dense, ds.dense) and torch.equal(ds2.labels, ds.labels)


def test_golden_batch_through_packed_dataset():
    g = load("ranking_layers_d8")
    fm = feature_map("ranking_layers_d8", 8)
    ds = PackedDataset(fm, g["batch"], pin=False)
    b = ds.batch(0, len(ds))
    X = b.columns()
    assert isinstance(X, PackedColumns) and list(X.keys()) == ds.cat_names + ds.num_names
    for n in ds.cat_names:
        assert torch.equal(X[n].long(), g["batch"][:, fm.get_column_index(n)].long())


def test_loader_surface_and_epoch_coverage():
    fm = feature_map("ranking_layers_d8", 8)
    arr = _data(1003, fm, 1)
    arr[:, -1] = np.arange(1003)                 # label = sample number, to track coverage
    dl = PackedDataLoader(fm, PackedDataset(fm, arr, compact=False), batch_size=128, shuffle=False)
    assert dl.num_samples == 1003 and dl.num_batches == 8 and len(dl) == 8
    seen = torch.cat([b.labels for b in dl])
    assert torch.equal(seen, torch.arange(1003).float())
    dl = PackedDataLoader(fm, PackedDataset(fm, arr, compact=False), batch_size=128, shuffle=True, seed=3)
    got, sizes = [], []
    for b in dl:
        assert isinstance(b, PackedBatch)
        got.append((b.labels.clone(), b.ids.clone(), b.dense.clone()))
        sizes.append(len(b))
    assert sizes == [128] * 7 + [1003 - 7 * 128]
    lab = torch.cat([x[0] for x in got]).long()
    assert sorted(lab.tolist()) == list(range(1003)) and lab.tolist() != list(range(1003))
    ids = torch.cat([x[1] for x in got])
    assert torch.equal(ids, dl.dataset.ids[lab])             # rows stay together under the shuffle
    assert torch.equal(torch.cat([x[2] for x in got]), dl.dataset.dense[lab])


def test_row_offsets_and_errors():
    fm = feature_map("ranking_layers_d8", 8)
    arr = _data(50, fm, 2)
    ds = PackedDataset(fm, arr, pin=False, compact=False)
    local = ds.ids.clone()
    ds.add_row_offsets([0, 11, 18, 31])
    assert torch.equal(ds.ids, local + torch.tensor([0, 11, 18, 31], dtype=torch.int32))
    with pytest.raises(RbxError):
        ds.add_row_offsets([0, 11, 18, 31])
    bad = arr.copy()
    bad[3, 1] = 99                                # C1 has vocab_size 11
    with pytest.raises(RbxError):
        PackedDataset(fm, bad, pin=False)
    with pytest.raises(RbxError):
        PackedDataset(fm, arr[:, :-1], pin=False)
    fm2 = feature_map("ranking_layers_seq", 8)
    with pytest.raises(RbxError):
        PackedDataset(fm2, np.zeros((4, 12)), pin=False)


def test_compact_uint16_blocks():
    """Every vocabulary below 65 536 -> the ids are stored (and shipped) as uint16: same values, half the bytes; the
    shuffled loader and the dict-of-columns view keep working; a vocabulary above 65 536 falls back to int32."""
    fm = feature_map("ranking_layers_d8", 8)
    arr = _data(777, fm, 5)
    wide = PackedDataset(fm, arr, pin=False, compact=False)
    ds = PackedDataset(fm, arr, pin=False)                       # compact="auto"
    assert ds.compact and ds.ids is None and ds.ids16.dtype == torch.int16
    assert torch.equal(ds.ids16.to(torch.int32) & 0xFFFF, wide.ids)
    assert ds.bytes_per_sample == 2 * 4 + 4 * (2 + 1) and wide.bytes_per_sample == 4 * (4 + 2 + 1)
    ds.add_row_offsets([0, 11, 18, 31])                          # compact ids stay local: the layer adds the offsets
    assert ds.offsets is None
    b = ds.batch(10, 60)
    assert b.ids is None and b.nbytes == 50 * ds.bytes_per_sample and len(b) == 50
    X = b.columns()
    for i, n in enumerate(ds.cat_names):
        assert torch.equal(X[n], wide.ids[10:60, i])
    dl = PackedDataLoader(fm, ds, batch_size=100, shuffle=True, seed=1)
    seen = torch.cat([(bb.ids16.to(torch.int32) & 0xFFFF) for bb in dl])
    assert sorted(map(tuple, seen.tolist())) == sorted(map(tuple, wide.ids.tolist()))
    # large ids survive the uint16 round trip (bit pattern, not sign)
    fm.features[ds.cat_names[0]]["vocab_size"] = 65536
    big = arr.copy()
    big[:, fm.get_column_index(ds.cat_names[0])] = np.arange(777) * 84 % 65536
    d2 = PackedDataset(fm, big, pin=False)
    assert d2.compact and torch.equal(d2.ids16[:, 0].to(torch.int32) & 0xFFFF, torch.from_numpy((np.arange(777) * 84 % 65536).astype(np.int32)))
    fm.features[ds.cat_names[0]]["vocab_size"] = 65537
    d3 = PackedDataset(fm, big, pin=False)
    assert not d3.compact and d3.ids is not None
    with pytest.raises(RbxError):
        PackedDataset(fm, big, pin=False, compact=True)
