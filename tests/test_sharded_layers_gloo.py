"""Row-sharded tables behind the layer API (layers.sharded_tables / _ShardedStore / _ShardedEmbedFn) on CPU: the
host-side logic -- which rows a rank keeps, the per-feature parameter slices, the zero / fence / backward protocol of the
autograd node, gradient views, replicated-vs-sharded gradient handling, gather_state_dict -- with tests/helpers.CpuKern
standing in for the CUDA kernels and a torch stand-in for the id packing (the product has no CPU path; the GPU run of the
same protocol is tests/test_sharded_layers_gpu.py).  World 1 in-process, world 2 over gloo."""
import os
import socket
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import CpuKern, Problem, assert_close

from recbox_b200 import RbxError, layers, sharded
from recbox_b200.features import FeatureMap


def _cpu_pack(self, inputs, cats, nums, offs, vocab=None):
    rows = torch.stack([inputs[n].long() + o for n, o in zip(cats, offs)], 1).int() if cats else None
    dense_x = torch.stack([inputs[n].float() for n in nums], 1) if nums else None
    return rows, dense_x


def _fmap(pb):
    fm = FeatureMap("sharded", ".")
    for k, v in pb.features.items():
        fm.features[k] = dict(v)
    fm.finalize(["label"])
    fm.default_emb_dim = pb.D
    return fm


def _build(pb):
    """FeatureEmbedding + FactorizationMachine of the reference's DeepFM front, built under the switch, loaded with the
    problem's FULL per-feature weights under the reference's names (as load_state_dict of a reference checkpoint does)."""
    fm = _fmap(pb)
    with layers.sharded_tables(mode="a2a", kern=CpuKern):
        emb = layers.FeatureEmbedding(fm, pb.D)
        fml = layers.FactorizationMachine(fm)
    sd = OrderedDict(("embedding_layer.embedding_layers.%s.weight" % n, pb.W[n].clone()) for n in pb.features)
    emb.load_state_dict(sd)
    sd1 = OrderedDict(("lr_layer.embedding_layer.embedding_layer.embedding_layers.%s.weight" % n, pb.W1[n].clone())
                      for n in pb.features)
    sd1["lr_layer.bias"] = pb.bias.clone()
    fml.load_state_dict(sd1)
    return fm, emb, fml, sd, sd1


def _run(pb, rank, world, monkey_target):
    B = pb.B // world
    sl = slice(rank * B, (rank + 1) * B)
    g = torch.Generator().manual_seed(4)
    Ft = pb.F + pb.Fn
    dE = torch.randn(pb.B, Ft, pb.D, generator=g)
    d_y = torch.randn(pb.B, generator=g)
    fm, emb, fml, sd, sd1 = _build(pb)
    assert not emb.embedding_layer._store.sharded                   # still the full tables on the host
    layers.shard_now(emb)
    layers.shard_now(fml)
    st = emb.embedding_layer._store
    assert st.sharded and fml.lr_layer.embedding_layer.embedding_layer._store.sharded
    # every per-feature parameter is this rank's rows of the reference's table, under the reference's name
    grp = st.groups[pb.D]
    for n, o in zip(pb.cat_names, pb.field_off):
        w = emb.embedding_layer.embedding_layers[n].weight
        assert layers.is_sharded(w) and torch.equal(w.data, pb.W[n][(rank - o) % world::world])
        assert w.data_ptr() == grp.sem.table[sharded.local_rows(o, world, rank)].data_ptr()
    X = OrderedDict((n, pb.X[n][sl]) for n in pb.features)
    params = list(emb.parameters()) + list(fml.parameters())

    def step():
        E = emb(X)
        y = fml(X, E)
        ((E * dE[sl]).sum() + (y.view(-1) * d_y[sl]).sum()).backward()
        layers.sync_replica_gradients(params)
        return E, y

    E, y = step()
    Er, fmr, lrr, *_ = pb.oracle_forward()
    assert torch.equal(E, Er[sl])
    assert_close(y.view(-1), (fmr.reshape(-1) + lrr.reshape(-1))[sl], atol_scale=1e-4, what="y")
    want = pb.oracle_grads(dE, d_y, d_y)
    got = layers.gather_state_dict(emb, grads=True)
    got1 = layers.gather_state_dict(fml, grads=True)
    gt = torch.cat([got["embedding_layer.embedding_layers.%s.weight" % n] for n in pb.cat_names], 0)
    gt1 = torch.cat([got1["lr_layer.embedding_layer.embedding_layer.embedding_layers.%s.weight" % n].reshape(-1)
                     for n in pb.cat_names], 0)
    gw = torch.stack([got["embedding_layer.embedding_layers.%s.weight" % n].reshape(-1) for n in pb.num_names], 0)
    gw1 = torch.cat([got1["lr_layer.embedding_layer.embedding_layer.embedding_layers.%s.weight" % n].reshape(-1)
                     for n in pb.num_names], 0)
    for a, b, name in zip((gt, gt1, gw, gw1, got1["lr_layer.bias"]), want, ("g_table", "g_table_lr", "g_dense_w", "g_dense_w_lr", "g_bias")):
        assert_close(a, b, atol_scale=2e-5, what=name)
    for p in pb.pad_row:
        assert float(gt[p].abs().sum()) == 0.0 and float(gt1[p]) == 0.0, "padding rows keep a zero gradient"
    # the table gradients handed to autograd are views of the gradient shard (no copy) ...
    w0 = emb.embedding_layer.embedding_layers[pb.cat_names[0]].weight
    assert w0.grad.data_ptr() == grp.sem.g_table.data_ptr()
    # ... and a second step gives the same gradients whether the optimizer dropped them (set_to_none) or kept the views
    keep = gt.clone()
    for p in params:
        p.grad = None
    step()
    again = layers.gather_state_dict(emb, grads=True)
    assert_close(torch.cat([again["embedding_layer.embedding_layers.%s.weight" % n] for n in pb.cat_names], 0), keep,
                 rtol=1e-6, atol_scale=1e-6, what="second step")
    for p in params:
        if p.grad is not None:
            p.grad.zero_()                              # zero_grad(set_to_none=False): .grad keeps aliasing the shard
    step()
    again = layers.gather_state_dict(emb, grads=True)
    assert_close(torch.cat([again["embedding_layer.embedding_layers.%s.weight" % n] for n in pb.cat_names], 0), keep,
                 rtol=1e-6, atol_scale=1e-6, what="third step (grads kept)")
    # clip: global norm over sharded (summed across ranks) + replicated (counted once) gradients
    full = [v for v in list(layers.gather_state_dict(emb, grads=True).values()) +
            list(layers.gather_state_dict(fml, grads=True).values()) if v is not None]
    norm_want = torch.sqrt(sum(v.double().pow(2).sum() for v in full))
    norm = layers.clip_grad_norm_(params, 0.5 * float(norm_want))
    assert_close(norm, norm_want.float(), rtol=1e-5, what="global norm")
    after = [v for v in list(layers.gather_state_dict(emb, grads=True).values()) +
             list(layers.gather_state_dict(fml, grads=True).values()) if v is not None]
    assert_close(torch.sqrt(sum(v.double().pow(2).sum() for v in after)).float(), 0.5 * norm_want.float(), rtol=1e-4,
                 what="clipped norm")
    # checkpoints: the gathered state_dict is the reference's, key for key and row for row
    back = layers.gather_state_dict(emb)
    assert list(back) == list(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    back1 = layers.gather_state_dict(fml)
    assert all(torch.equal(back1[k], sd1[k]) for k in sd1)
    # one forward / backward per dictionary and step: a stale backward fails loudly instead of double counting
    E1 = emb(X)
    E2 = emb(X)
    E2.sum().backward()
    with pytest.raises(RbxError):
        E1.sum().backward()


def test_sharded_layers_world1_host(monkeypatch):
    monkeypatch.setattr(layers._FusedDictBase, "_pack", _cpu_pack)
    pb = Problem(40, "nccncc", 8, vocab=[11, 7, 13, 5], seed=3)
    _run(pb, 0, 1, None)


def test_switch_is_scoped_and_env_driven(monkeypatch):
    pb = Problem(8, "cc", 8, vocab=[5, 4], seed=1)
    fm = _fmap(pb)
    assert type(layers.FeatureEmbedding(fm, 8).embedding_layer._store) is layers._FusedStore
    with layers.sharded_tables():
        assert type(layers.FeatureEmbedding(fm, 8).embedding_layer._store) is layers._ShardedStore
    assert type(layers.FeatureEmbedding(fm, 8).embedding_layer._store) is layers._FusedStore
    monkeypatch.setenv("RECBOX_B200_SHARD", "peer")
    st = layers.FeatureEmbedding(fm, 8).embedding_layer._store
    assert type(st) is layers._ShardedStore and st.cfg.mode == "peer" and not st.sharded


def test_sharded_dictionary_rejects_what_it_does_not_cover(monkeypatch):
    monkeypatch.setattr(layers._FusedDictBase, "_pack", _cpu_pack)
    pb = Problem(8, "cc", 6, vocab=[5, 4], seed=1)              # D = 6: not a row width the sharded kernels cover
    with layers.sharded_tables(mode="a2a", kern=CpuKern):
        emb = layers.FeatureEmbedding(_fmap(pb), 6)
    with pytest.raises(RbxError):
        layers.shard_now(emb)
    pb = Problem(8, "cc", 8, vocab=[5, 4], seed=1)
    with layers.sharded_tables():                                # the product configuration has no CPU path
        emb = layers.FeatureEmbedding(_fmap(pb), 8)
    with pytest.raises(RbxError):
        layers.shard_now(emb)
    with pytest.raises(RbxError):
        emb({"C0": torch.zeros(2), "C1": torch.zeros(2)})
    with layers.sharded_tables(mode="a2a", kern=CpuKern):
        emb = layers.FeatureEmbedding(_fmap(pb), 8)
    layers.shard_now(emb)
    import copy
    with pytest.raises(RbxError):
        copy.deepcopy(emb)
    with pytest.raises(RbxError):                                # re-pointing a parameter after the cut
        emb.embedding_layer.embedding_layers["C0"].weight = torch.nn.Parameter(torch.zeros(5, 8))
        emb({"C0": torch.zeros(2), "C1": torch.zeros(2)})


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layers._FusedDictBase._pack = _cpu_pack
        pb = Problem(24 * world, "nccncc", 8, vocab=[11, 7, 13, 5], seed=3)
        _run(pb, rank, world, None)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_layers_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
