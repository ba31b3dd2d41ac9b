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


def _build(pb, mode="a2a"):
    """FeatureEmbedding + FactorizationMachine of the reference's DeepFM front, built under the switch, loaded with the
    problem's FULL per-feature weights under the reference's names (as load_state_dict of a reference checkpoint does)."""
    fm = _fmap(pb)
    with layers.sharded_tables(mode=mode, kern=CpuKern, max_ids=pb.B * pb.F, slack=3.0):
        emb = layers.FeatureEmbedding(fm, pb.D)
        fml = layers.FactorizationMachine(fm)
    sd = OrderedDict(("embedding_layer.embedding_layers.%s.weight" % n, pb.W[n].clone()) for n in pb.features)
    emb.load_state_dict(sd)
    sd1 = OrderedDict(("lr_layer.embedding_layer.embedding_layer.embedding_layers.%s.weight" % n, pb.W1[n].clone())
                      for n in pb.features)
    sd1["lr_layer.bias"] = pb.bias.clone()
    fml.load_state_dict(sd1)
    return fm, emb, fml, sd, sd1


def _run(pb, rank, world, mode="a2a"):
    B = pb.B // world
    sl = slice(rank * B, (rank + 1) * B)
    g = torch.Generator().manual_seed(4)
    Ft = pb.F + pb.Fn
    dE = torch.randn(pb.B, Ft, pb.D, generator=g)
    d_y = torch.randn(pb.B, generator=g)
    fm, emb, fml, sd, sd1 = _build(pb, mode)
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
    _run(pb, 0, 1)


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


def _worker(rank, world, port, out_dir, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layers._FusedDictBase._pack = _cpu_pack
        pb = Problem(24 * world, "nccncc", 8, vocab=[11, 7, 13, 5], seed=3)
        _run(pb, rank, world, mode)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["a2a", "stream"])
def test_sharded_layers_world2_gloo(mode, tmp_path):
    """The layer protocol over the NCCL-style exchange ("a2a") and over the streamed exchange ("stream": peer-visible
    workspaces, flag barriers, parity-double-buffered inboxes -- file-backed stand-ins here)."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _ref_trainer_worker(rank, world, port, out_dir):
    """layers.train_step on a RankingModel subclass (the reference's base class) whose tables are sharded over `world` ranks
    against the ALL-reference model trained by its own train_step on the whole batch."""
    import tempfile
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import recbox_b200
        from oracle import ref_shim
        from test_oracle_golden import ranking_features
        L = ref_shim.install()
        from recbox.ranking.features import FeatureMap as RefFeatureMap
        from recbox.ranking.pytorch.models.ranking_model import RankingModel
        layers._FusedDictBase._pack = _cpu_pack
        tmp = tempfile.mkdtemp()
        fm = RefFeatureMap("golden", tmp)
        for k, v in ranking_features("ranking_layers_d8").items():
            fm.features[k] = dict(v)
        fm.labels = ["label"]
        fm.num_fields = fm.get_num_fields()
        fm.set_column_index()
        fm.default_emb_dim = 8

        def make(LL):
            class FM(RankingModel):                       # FuxiCTR-style FM body against whatever layer set LL is
                def __init__(self, feature_map, **kw):
                    super(FM, self).__init__(feature_map, **kw)
                    self.embedding_layer = LL.FeatureEmbedding(feature_map, 8)
                    self.fm_layer = LL.FactorizationMachine(feature_map)
                    self.compile("adam", "binary_cross_entropy", 1e-3)
                    self.reset_parameters()
                    self.model_to_device()

                def forward(self, inputs):
                    X = self.get_inputs(inputs)
                    y = self.fm_layer(X, self.embedding_layer(X))
                    return {"y_pred": self.output_activation(y)}
            m = FM(fm, model_id="m", gpu=-1, verbose=0, model_root=tmp, metrics=["AUC"])
            m._max_gradient_norm = 10.
            return m

        torch.manual_seed(5)
        ref = make(L)                                     # the reference's own layers
        g = torch.Generator().manual_seed(6)
        with torch.no_grad():                             # O(0.1) weights so gradients and Adam moments are not ~1e-8
            for k, p in ref.named_parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            for mod in ref.modules():
                if isinstance(mod, torch.nn.Embedding) and mod.padding_idx is not None:
                    mod.weight[mod.padding_idx] = 0
        init = {k: v.clone() for k, v in ref.state_dict().items()}
        try:
            recbox_b200.install()
            with layers.sharded_tables(mode="a2a", kern=CpuKern):
                ours = make(L)                            # same class body, rebound layer set, tables to be sharded
        finally:
            recbox_b200.uninstall()
        ours.load_state_dict(init)
        layers.shard_now(ours)
        B = 64
        gb = torch.Generator().manual_seed(7)
        for step in range(3):
            cols = [torch.rand(B, generator=gb, dtype=torch.float64) if s["type"] == "numeric"
                    else torch.randint(0, s["vocab_size"], (B,), generator=gb).double() for s in fm.features.values()]
            batch = torch.stack(cols + [(torch.rand(B, generator=gb) < 0.5).double()], 1)
            ref.train()
            want = float(ref.train_step(batch))
            ours.train()
            sl = slice(rank * (B // world), (rank + 1) * (B // world))
            got = float(layers.train_step(ours, batch[sl]))
            assert abs(got - want) <= 2e-6 * abs(want), (step, got, want)
        final = layers.gather_state_dict(ours)
        for k, v in ref.state_dict().items():
            assert_close(final[k], v, rtol=1e-4, atol_scale=2e-5, what="final." + k)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        if world > 1:
            dist.destroy_process_group()


@pytest.mark.reference
@pytest.mark.parametrize("world", [1, 2])
def test_reference_trainer_step_on_sharded_tables(world, tmp_path):
    if world == 1:
        import multiprocessing
        ctx = multiprocessing.get_context("spawn")           # own process: the worker patches layers and installs the shim
        p = ctx.Process(target=_ref_trainer_worker, args=(0, 1, 0, str(tmp_path)))
        p.start()
        p.join()
        assert p.exitcode == 0
    else:
        mp.spawn(_ref_trainer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
