"""Host-side logic of the row-sharded path (recbox_b200.sharded, mode="a2a") with world_size 2 over
gloo on CPU: split sizes, buffer bookkeeping, owner-side scatter and pad rows -- with the oracle
standing in for the CUDA kernels (tests/helpers.CpuKern).  The sharded result must equal the
single-table oracle result on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import CpuKern, Problem, assert_close

from recbox_b200 import sharded


def test_row_ownership_arithmetic():
    for R in (0, 1, 7, 8, 9, 1000003):
        for world in (1, 2, 3, 8):
            ns = [sharded.local_rows(R, world, r) for r in range(world)]
            assert sum(ns) == R and max(ns) <= sharded.shard_capacity(R, world)
            assert ns == [len(range(r, R, world)) for r in range(world)]
    assert sharded.owned_pad_rows([0, 11, 18, 31, -1], 2, 0) == [0, 9]
    assert sharded.owned_pad_rows([0, 11, 18, 31, -1], 2, 1) == [5, 15]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, mode="a2a", layout="split", chunks=1):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, D = 46 if mode == "stream" else 48, 8       # 46: the stream stand-in's last tile is partial
        pb = Problem(B * world, "nccncc", D, vocab=[11, 7, 13, 5], seed=3)
        f = pb.fused("cpu")
        g = torch.Generator().manual_seed(4)
        dE = torch.randn(B * world, 6, D, generator=g)
        d_fm = torch.randn(B * world, generator=g)
        d_lr = torch.randn(B * world, generator=g)
        sl = slice(rank * B, (rank + 1) * B)
        sh = sharded.ShardedEmbeddingFM(pb.R, D, mode=mode, device="cpu", kern=CpuKern, max_ids=B * pb.F, slack=3.0,
                                        layout=layout, chunks=chunks)
        assert sh.world == world and sh.rank == rank
        sh.load_global(f["table"], f["table_lr"])
        gw, gw1, gb = torch.zeros_like(f["dense_w"]), torch.zeros_like(f["dense_w_lr"]), torch.zeros(1)
        sh.zero_grad()
        # three steps on the same batch (stream mode alternates its parity buffers; gradients accumulate), the last
        # one without the saved E (e re-read from the row buffer)
        steps = 3 if mode == "stream" else 1
        for it in range(steps):
            E, S, fm, lr = sh.forward(f["rows"][sl], pb.cat_pos, f["dense_x"][sl], f["dense_w"], f["dense_w_lr"],
                                      pb.num_pos, f["bias"])
            sh.backward(f["rows"][sl], pb.cat_pos, pb.pad_row, f["dense_x"][sl], f["dense_w"], pb.num_pos,
                        E if it < 2 else None, S, dE[sl], d_fm[sl], d_lr[sl], gw, gw1, gb)
        sh.barrier()
        sh.check_overflow()
        for t in (gw, gw1, gb):
            dist.all_reduce(t)
        gt = sh.gather_global("g_table")
        gt1 = sh.gather_global("g_table_lr")
        # reference: the oracle on the whole batch / whole table
        Er, fmr, lrr, *_ = pb.oracle_forward()
        assert torch.equal(E, Er[sl])
        assert_close(fm, fmr.reshape(-1)[sl], atol_scale=1e-4, what="fm")
        assert_close(lr, lrr.reshape(-1)[sl], what="lr")
        want = pb.oracle_grads(dE, d_fm, d_lr)
        for got, ref, name in zip((gt, gt1, gw, gw1, gb), want, ("g_table", "g_table_lr", "g_dense_w", "g_dense_w_lr", "g_bias")):
            assert_close(got, ref * steps, atol_scale=2e-5, what=name)
        for p in pb.pad_row:
            assert float(gt[p].abs().sum()) == 0.0 and float(gt1[p]) == 0.0, "padding rows keep a zero gradient"
        sh.close()
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_a2a_orchestration_world2_gloo(world, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


@pytest.mark.parametrize("layout,chunks", [("split", 1), ("rowlr", 1), ("rowpad", 1), ("split", 2)])
def test_stream_orchestration_world2_gloo(layout, chunks, tmp_path):
    """mode="stream": slot bookkeeping, parity-double-buffered inboxes, barrier epochs and the three shard layouts, with the
    CPU stand-ins writing into the peers' (file-backed) blocks exactly where the kernels write over NVLink."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), "stream", layout, chunks), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _replica_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recbox_b200 import layers
        # the shape a fused backward leaves behind: per-feature .grad tensors that alias ONE buffer (detached, so not
        # "views" in autograd's sense), plus two gradients that own their storage
        buf = torch.arange(24, dtype=torch.float32) * (rank + 1)
        ps = [torch.nn.Parameter(torch.zeros(4, 4)), torch.nn.Parameter(torch.zeros(2, 4)), torch.nn.Parameter(torch.zeros(3)),
              torch.nn.Parameter(torch.zeros(2, 2))]
        ps[0].grad, ps[1].grad = buf[:16].view(4, 4).detach(), buf[16:24].view(2, 4).detach()
        ps[2].grad = torch.full((3,), float(rank + 1))
        ps[3].grad = torch.full((2, 2), 10.0 * (rank + 1))
        shared, single = layers.fused_grad_buffers(ps)
        assert len(shared) == 1 and shared[0].data_ptr() == buf.data_ptr() and shared[0].numel() == 24 and len(single) == 2
        assert layers.sync_replica_gradients(ps) == 2             # the fused buffer + one bucket for the rest, not four
        tot = sum(r + 1 for r in range(world))
        assert torch.equal(ps[0].grad, (torch.arange(16, dtype=torch.float32) * tot).view(4, 4))
        assert torch.equal(ps[1].grad, (torch.arange(16, 24, dtype=torch.float32) * tot).view(2, 4))
        assert torch.equal(ps[2].grad, torch.full((3,), float(tot)))
        assert torch.equal(ps[3].grad, torch.full((2, 2), 10.0 * tot))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_replica_gradient_sync_world2_gloo(tmp_path):
    """Replica mode (SURVEY 8e "replicas only"): one all-reduce per fused gradient buffer, results land in every
    parameter's .grad view."""
    world = 2
    mp.spawn(_replica_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


def _reducer_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recbox_b200 import layers, replica
        red = replica.ReplicaReducer(10, "cpu")              # no multicast memory on the host: the collective fallback
        assert red.numel == 12 and red.mc == 0 and "all_reduce" in red.path
        red.buffer.copy_(torch.arange(12, dtype=torch.float32) * (rank + 1))
        tot = sum(r + 1 for r in range(world))
        assert torch.equal(red.all_reduce(), torch.arange(12, dtype=torch.float32) * tot)
        t = torch.full((7,), float(rank + 1))
        assert torch.equal(red.reduce_tensor(t, average=True), torch.full((7,), tot / world))
        buf = torch.arange(8, dtype=torch.float32) * (rank + 1)
        ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(4))]
        ps[0].grad, ps[1].grad = buf[:4].detach(), buf[4:].detach()
        assert layers.sync_replica_gradients(ps, reducer=red) == 1
        assert torch.equal(ps[1].grad, torch.arange(4, 8, dtype=torch.float32) * tot)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_replica_reducer_fallback_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))


class _SparseKern(CpuKern):
    """CPU stand-ins for the three kernels SparseRowExchange calls."""

    @staticmethod
    def unique_ids(ids, vocab, want_first=True, want_inverse=True, sync=True):
        u = torch.unique(ids.reshape(-1))
        cap = max(min(ids.numel(), vocab), 1)
        buf = torch.full((cap,), 12345, dtype=ids.dtype)          # capacity buffer, garbage past the count (as on the GPU)
        buf[:u.numel()] = u
        return buf, None, None, torch.tensor([u.numel(), 0])

    @staticmethod
    def gather_rows(table, ids):
        out = torch.zeros(tuple(ids.shape) + (table.shape[1],))
        ok = ids >= 0
        out[ok] = table[ids[ok].long()]
        return out

    @staticmethod
    def scatter_add_rows(g, ids, pad_row, g_table):
        keep = (ids >= 0) & (ids != pad_row)
        g_table.index_add_(0, ids[keep].long(), g[keep])


def _sparse_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recbox_b200 import replica
        R, D, n = 50, 4, 12
        gen = torch.Generator().manual_seed(3)
        rows = [torch.randint(0, R, (n,), generator=gen, dtype=torch.int32) for _ in range(world)]
        grads = [torch.randn(n, D, generator=gen) for _ in range(world)]
        want = torch.zeros(R, D)
        for r_, g_ in zip(rows, grads):
            want.index_add_(0, r_.long(), g_)
        g_table = torch.zeros(R, D)
        g_table.index_add_(0, rows[rank].long(), grads[rank])           # the local backward
        ex = replica.SparseRowExchange(R, D, n, "cpu", kern=_SparseKern)
        touched = ex.exchange(g_table, rows[rank])
        assert torch.allclose(g_table, want, atol=1e-6)
        assert set(touched[touched >= 0].tolist()) == set(torch.cat(rows).tolist())
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sparse_row_exchange_world2_gloo(tmp_path):
    """Replicas of a table the batch barely touches: all_gather of (row id, gradient row) blocks, every replica ends with the
    summed gradient on the union of the touched rows."""
    world = 2
    mp.spawn(_sparse_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
