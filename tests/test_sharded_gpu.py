"""Row-sharded table on real GPUs (one process per GPU, NCCL): both data paths -- the NVLink
peer-memory fused kernels ("peer") and the NCCL all-to-all baseline ("a2a") -- must reproduce the
single-table fused path (which tests/test_kernels_gpu.py pins to the oracle) on the same inputs:
E bit-exact, sums / gradients within 1e-5.  World = min(#GPUs, 8) rounded down to a power of two
(1 on a single-GPU box: the sharded kernels still run, with one shard)."""
import os
import socket

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


def _worker(rank, world, port, mode, D, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recbox_b200 import ops, sharded
        B = 1000
        pb = Problem(B * world, "nn" + "c" * 9 + "n", D, vocab=[101, 7, 13, 1000, 3, 50, 77, 256, 19], seed=11 + D, zipf=1.3)
        f = pb.fused(dev)
        g = torch.Generator().manual_seed(5)
        Ft = pb.F + pb.Fn
        dE = torch.randn(B * world, Ft, D, generator=g).to(dev)
        d_fm = torch.randn(B * world, generator=g).to(dev)
        d_lr = torch.randn(B * world, generator=g).to(dev)
        sl = slice(rank * B, (rank + 1) * B)
        rows, dx = f["rows"][sl].contiguous(), f["dense_x"][sl].contiguous()

        parts = mode.split("_")                     # mode[_layout[_chunks]]
        mode, layout, chunks = parts[0], (parts[1] if len(parts) > 1 else "split"), (int(parts[2]) if len(parts) > 2 else 1)
        sh = sharded.ShardedEmbeddingFM(pb.R, D, mode=mode, device=dev, max_ids=B * pb.F, slack=3.0, layout=layout,
                                        chunks=chunks)
        sh.load_global(f["table"], f["table_lr"])
        E, S, fm, lr = sh.forward(rows, pb.cat_pos, dx, f["dense_w"], f["dense_w_lr"], pb.num_pos, f["bias"])
        if mode == "stream":       # a second forward lands in the other parity buffers and must agree bit for bit
            E2, S2, fm2, lr2 = sh.forward(rows, pb.cat_pos, dx, f["dense_w"], f["dense_w_lr"], pb.num_pos, f["bias"])
            assert torch.equal(E, E2) and torch.equal(S, S2) and torch.equal(fm, fm2) and torch.equal(lr, lr2)
        gw, gw1, gb = torch.zeros_like(f["dense_w"]), torch.zeros_like(f["dense_w_lr"]), torch.zeros(1, device=dev)
        sh.zero_grad()
        sh.barrier()
        for use_E in (True, False):
            sh.backward(rows, pb.cat_pos, pb.pad_row, dx, f["dense_w"], pb.num_pos, E if use_E else None, S,
                        dE[sl].contiguous(), d_fm[sl].contiguous(), d_lr[sl].contiguous(), gw, gw1, gb)
        sh.barrier()
        sh.check_overflow()
        for t in (gw, gw1, gb):
            dist.all_reduce(t)
        gt, gt1 = sh.gather_global("g_table"), sh.gather_global("g_table_lr")

        # single-table fused path on the whole batch (every rank computes it locally)
        Er, Sr, fmr, lrr = ops.embed_fm_fwd(f["table"], f["table_lr"], f["rows"], pb.cat_pos, f["dense_x"], f["dense_w"],
                                            f["dense_w_lr"], pb.num_pos, f["bias"])
        assert torch.equal(E, Er[sl]), "gathered rows must be bit-exact"
        if mode == "stream":    # the tile kernel sums the slots in a different order than the single-table kernel
            assert_close(S, Sr[sl], what="S[stream]")
            assert_close(fm, fmr[sl], atol_scale=2e-5, what="fm[stream]")
            assert_close(lr, lrr[sl], what="lr[stream]")
        else:
            assert torch.equal(S, Sr[sl]) and torch.equal(fm, fmr[sl]), "same kernel, same order"
        if mode == "stream":
            pass
        elif layout == "rowlr":   # the first-order weights ride in the row: one lane sums them, in slot order
            assert_close(lr, lrr[sl], what="lr[rowlr]")
        else:
            assert torch.equal(lr, lrr[sl]), "same kernel, same order"
        if mode == "push":      # a bucket larger than its slot is flagged, not silently dropped
            tiny = sharded.ShardedEmbeddingFM(pb.R, D, mode="push", device=dev, max_ids=B * pb.F, slack=0.0)
            tiny.slot_cap = 8
            tiny.forward(rows, pb.cat_pos, dx, f["dense_w"], f["dense_w_lr"], pb.num_pos, f["bias"])
            with pytest.raises(sharded.RbxError):
                tiny.check_overflow()
            tiny.close()
        rt, rt1 = torch.zeros_like(f["table"]), torch.zeros_like(f["table_lr"])
        rw, rw1, rb = torch.zeros_like(gw), torch.zeros_like(gw1), torch.zeros_like(gb)
        for use_E in (True, False):
            ops.embed_fm_bwd(f["table"], f["rows"], pb.cat_pos, pb.pad_row, f["dense_x"], f["dense_w"], pb.num_pos,
                             Er if use_E else None, Sr, dE, d_fm, d_lr, rt, rt1, rw, rw1, rb, D, pb.R)
        for got, ref, name in zip((gt, gt1, gw, gw1, gb), (rt, rt1, rw, rw1, rb),
                                  ("g_table", "g_table_lr", "g_dense_w", "g_dense_w_lr", "g_bias")):
            assert_close(got, ref, atol_scale=2e-5, what="%s[%s]" % (name, mode))
        for p in pb.pad_row:
            assert float(gt[p].abs().sum()) == 0.0 and float(gt1[p]) == 0.0, "padding rows keep a zero gradient"
        sh.close()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def _world():
    n = torch.cuda.device_count()
    w = 1
    while w * 2 <= min(n, 8):
        w *= 2
    return w


@pytest.mark.parametrize("D", [16, 64])
@pytest.mark.parametrize("mode", ["stream", "peer", "push", "a2a"])
def test_sharded_matches_single_table(mode, D, tmp_path):
    world = _world()
    mp.spawn(_worker, args=(world, _free_port(), mode, D, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


@pytest.mark.parametrize("D", [4, 16, 32, 128])
@pytest.mark.parametrize("layout", ["rowlr", "rowpad"])
def test_sharded_stream_layouts_match_single_table(layout, D, tmp_path):
    """Streamed exchange with the row + first-order weight in one physical row (2 D or D + 4 floats)."""
    world = _world()
    mp.spawn(_worker, args=(world, _free_port(), "stream_" + layout, D, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


@pytest.mark.parametrize("chunks", [2, 3])
def test_sharded_stream_chunk_lanes_match_single_table(chunks, tmp_path):
    """The batch cut into sample ranges that run the exchange on their own streams / workspaces (overlap of the
    NVLink-bound and the HBM-bound phases) gives the same result as one range."""
    world = _world()
    mp.spawn(_worker, args=(world, _free_port(), "stream_rowlr_%d" % chunks, 16, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


@pytest.mark.parametrize("D", [16, 8, 4])
def test_sharded_rowlr_layout_matches_single_table(D, tmp_path):
    """ROW+LR shard layout (row and first-order weight in one physical row / one NVLink request)."""
    world, mode = _world(), "peer_rowlr"
    mp.spawn(_worker, args=(world, _free_port(), mode, D, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _reducer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recbox_b200 import layers, replica
        n = 1_000_003                                   # not a multiple of 4 * world: ragged slices
        red = replica.ReplicaReducer(n, dev)
        g = torch.Generator().manual_seed(7)
        parts = [torch.randn(red.numel, generator=g) for _ in range(world)]
        want = torch.stack(parts).double().sum(0)
        for it in range(3):                             # repeated use: the barriers re-arm
            red.buffer.copy_(parts[rank].to(dev))
            got = red.all_reduce()
            assert_close(got, want, rtol=1e-6, atol_scale=1e-6, what="all_reduce[%s] pass %d" % (red.path[:24], it))
        t = parts[rank][:777].to(dev).clone()
        red.reduce_tensor(t, average=True)
        assert_close(t, want[:777] / world, rtol=1e-6, atol_scale=1e-6, what="reduce_tensor")
        open(os.path.join(out_dir, "ok%d" % rank), "w").write(red.path)
    finally:
        dist.destroy_process_group()


def test_replica_reducer_in_switch_all_reduce(tmp_path):
    """Replica gradient exchange (SURVEY 8e "replicas only"): rbx_nvls_allreduce_f32 over multicast memory (or the collective
    fallback where the box has no NVSwitch multicast) sums the symmetric buffer of every rank."""
    world = _world()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _sparse_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from recbox_b200 import ops, optim, replica
        R, D, n = 100_003, 64, 5000
        gen = torch.Generator().manual_seed(3)
        rows = [torch.randint(0, R, (n,), generator=gen, dtype=torch.int32) for _ in range(world)]
        grads = [torch.randn(n, D, generator=gen) for _ in range(world)]
        want = torch.zeros(R, D)
        for r_, g_ in zip(rows, grads):
            want.index_add_(0, r_.long(), g_)
        g_table = torch.zeros(R, D, device=dev)
        ops.scatter_add_rows(grads[rank].to(dev), rows[rank].to(dev), -1, g_table)      # the local backward
        ex = replica.SparseRowExchange(R, D, n, dev)
        touched = ex.exchange(g_table, rows[rank].to(dev))
        assert_close(g_table, want, rtol=1e-6, atol_scale=1e-6, what="summed sparse gradient")
        got = set(touched[touched >= 0].cpu().tolist())
        assert got == set(torch.cat(rows).tolist())
        # and the touched-rows optimizer consumes the padded id list (ids of -1 are counted as out of range, not updated)
        w = torch.ones(R, D, device=dev)
        opt = optim.TouchedRowsOptimizer([(w, g_table)], kind="sgd", lr=0.5)
        opt.step(touched)
        assert_close(w, 1.0 - 0.5 * want, rtol=1e-6, atol_scale=1e-6, what="sgd on the union of the touched rows")
        assert float(g_table.abs().sum()) == 0.0                                        # consumed rows are cleared
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sparse_row_exchange_of_replicas(tmp_path):
    """Replicas of a table the batch barely touches: all_gather of (row id, gradient row) blocks; every replica ends
    with the summed gradient on the union of the touched rows, and the touched-rows optimizer applies it."""
    world = _world()
    if world < 2:
        pytest.skip("needs at least two GPUs")
    mp.spawn(_sparse_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
