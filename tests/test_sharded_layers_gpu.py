"""Row-sharded tables behind the layer API on real GPUs (layers.sharded_tables; one process per GPU, NCCL for the
plumbing, the NVLink peer-memory fused kernels for the data path): an UNMODIFIED DeepFM module built under the switch,
loaded with the reference's full checkpoint and moved to its GPU, must walk the reference's own 3-step training
trajectory (tests/golden/deepfm_train.npz, minted from the unmodified reference by oracle/make_golden.py) when every
rank feeds its slice of each batch -- losses, first-step gradients, final weights and predictions -- and must agree with
the replicated layer path on a larger Zipf-skewed problem.  World = min(#GPUs, 8) rounded down to a power of two
(1 on a single-GPU box: the sharded kernels and the whole protocol still run, with one shard)."""
import os
import socket
from collections import OrderedDict

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import Problem, assert_close

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _world():
    n = torch.cuda.device_count()
    w = 1
    while w * 2 <= min(n, 8):
        w *= 2
    return w


def _golden_deepfm(rank, world, dev):
    from test_layers_gpu import _DeepFM, _sub
    from test_layers_host import feature_map
    from test_oracle_golden import load
    from recbox_b200 import layers
    g = load("deepfm_train")
    fm = feature_map("ranking_layers_d8", 8)
    with layers.sharded_tables(mode="peer"):
        model = _DeepFM(fm, 8, (16, 8))                       # the test-suite's DeepFM, untouched
    init = _sub(g, "init.")
    assert sorted(model.state_dict()) == sorted(init)
    model.load_state_dict(init)                               # the reference's FULL tables, reference key names
    model.to(dev)                                             # <- cut here
    d = model.embedding_layer.embedding_layer
    assert d._store.sharded and model.fm_layer.lr_layer.embedding_layer.embedding_layer._store.sharded
    off = 0
    for n in ("C1", "C2", "C3", "C4"):
        w = d.embedding_layers[n].weight
        full = init["embedding_layer.embedding_layer.embedding_layers.%s.weight" % n]
        assert layers.is_sharded(w) and torch.equal(w.detach().cpu(), full[(rank - off) % world::world])
        off += full.shape[0]
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3)
    B = g["batch0"].shape[0]
    Bl = B // world
    sl = slice(rank * Bl, (rank + 1) * Bl)
    losses = []
    for s in range(3):
        batch = g["batch%d" % s][sl].contiguous().to(dev)
        y_true = batch[:, -1].float().view(-1, 1)
        opt.zero_grad()
        # RankingModel.train_step (ranking_model.py:191-197) on this rank's slice: the mean over the GLOBAL batch
        loss = torch.nn.functional.binary_cross_entropy(model(batch), y_true, reduction="sum") / B
        loss.backward()
        layers.sync_replica_gradients(params)                 # replicated parameters only; sharded rows have one owner
        if s == 0:
            got = layers.gather_state_dict(model, grads=True)
            want = _sub(g, "grad0.")
            assert sorted(got) == sorted(want)
            for k in want:
                assert got[k] is not None, k
                assert_close(got[k], want[k], atol_scale=2e-5, what="grad0." + k)
        layers.clip_grad_norm_(params, 10.0)
        opt.step()
        tot = loss.detach().clone()
        if world > 1:
            dist.all_reduce(tot)
        losses.append(float(tot))
    assert_close(torch.tensor(losses), g["losses"].float(), rtol=5e-6, atol_scale=0, what="losses")
    final = _sub(g, "final.")
    got = layers.gather_state_dict(model)
    assert list(got) == list(model.state_dict())
    for k, v in got.items():
        assert tuple(v.shape) == tuple(final[k].shape), k
        assert_close(v, final[k], rtol=1e-4, atol_scale=1e-5, what="final." + k)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()                                        # every rank's last optimizer step has landed
    model.eval()
    with torch.no_grad():
        pred = model(g["batch0"][sl].contiguous().to(dev))
    assert_close(pred, g["pred_final"][sl], rtol=1e-5, what="pred_final")


def _against_replicated(rank, world, dev):
    """Sharded vs replicated layer path: D = 16, 3 numeric + 10 categorical slots, Zipf-skewed ids (many samples of a
    warp hit the same row, rows of one sample live on different ranks), padding rows, three steps."""
    from test_sharded_layers_gloo import _fmap
    from recbox_b200 import layers
    Bl, D = 2048, 16
    vocab = [1001, 17, 50000, 3, 256, 4099, 77, 12345, 640, 31]
    pb = Problem(Bl * world, "ncccncccccncc", D, vocab=vocab, seed=21, zipf=1.2)
    fm = _fmap(pb)
    sd = OrderedDict(("embedding_layer.embedding_layers.%s.weight" % n, pb.W[n].clone()) for n in pb.features)
    sd1 = OrderedDict(("lr_layer.embedding_layer.embedding_layer.embedding_layers.%s.weight" % n, pb.W1[n].clone())
                      for n in pb.features)
    sd1["lr_layer.bias"] = pb.bias.clone()

    def build(shard):
        if shard:
            with layers.sharded_tables(mode="peer"):
                emb, fml = layers.FeatureEmbedding(fm, D), layers.FactorizationMachine(fm)
        else:
            emb, fml = layers.FeatureEmbedding(fm, D), layers.FactorizationMachine(fm)
        emb.load_state_dict(sd)
        fml.load_state_dict(sd1)
        return emb.to(dev), fml.to(dev)

    rep, rep_fm = build(False)
    sh, sh_fm = build(True)
    g = torch.Generator().manual_seed(9)
    wE = torch.randn(pb.B, pb.F + pb.Fn, D, generator=g).to(dev)
    wy = torch.randn(pb.B, 1, generator=g).to(dev)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    Xf = OrderedDict((n, pb.X[n].to(dev)) for n in pb.features)
    Xl = OrderedDict((n, v[sl].contiguous()) for n, v in Xf.items())
    sh_params = list(sh.parameters()) + list(sh_fm.parameters())
    for it in range(3):
        for p in list(rep.parameters()) + list(rep_fm.parameters()) + sh_params:
            p.grad = None
        Er = rep(Xf)
        yr = rep_fm(Xf, Er)
        ((Er * wE).sum() + (yr * wy).sum()).backward()
        E = sh(Xl)
        y = sh_fm(Xl, E)
        ((E * wE[sl]).sum() + (y * wy[sl]).sum()).backward()
        layers.sync_replica_gradients(sh_params)
        assert torch.equal(E, Er[sl]), "gathered rows must be bit-exact"
        assert_close(y, yr[sl], atol_scale=2e-5, what="y")
        for (mod_s, mod_r, tag) in ((sh, rep, "emb."), (sh_fm, rep_fm, "fm.")):
            got = layers.gather_state_dict(mod_s, grads=True)
            want = OrderedDict((k, p.grad) for k, p in mod_r.named_parameters(remove_duplicate=False))
            assert sorted(got) == sorted(want)
            for k in want:
                assert_close(got[k], want[k], atol_scale=2e-5, what="step %d grad %s%s" % (it, tag, k))
        for n, o in zip(pb.cat_names, pb.field_off):           # padding rows (id 0 of every feature) keep a zero gradient
            gw = layers.gather_state_dict(sh, grads=True)["embedding_layer.embedding_layers.%s.weight" % n]
            assert float(gw[0].abs().sum()) == 0.0


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        _golden_deepfm(rank, world, dev)
        _against_replicated(rank, world, dev)
        torch.cuda.synchronize(dev)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_deepfm_walks_the_reference_trajectory(tmp_path):
    world = _world()
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
