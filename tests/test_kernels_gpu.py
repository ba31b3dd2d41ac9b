"""Parity of every C-ABI entry point (called through recbox_b200.ops = ctypes) against the CPU
oracle on the same seeded inputs.  Bit-exact for index / gather / byte work, 1e-5 relative (fp32)
for sums (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from helpers import Problem, assert_close, oracle
from recbox_b200 import RbxError

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from recbox_b200 import ops
    return ops


# ----------------------------------------------------------------------------- K1/K2 forward
@pytest.mark.parametrize("D", [4, 8, 16, 32, 64, 128, 10, 1, 40, 200])
@pytest.mark.parametrize("kinds,B", [("nncccccn", 257), ("ccc", 64), ("nn", 33), ("c" * 26 + "n" * 13, 1000), ("cnc", 1)])
def test_embed_fm_fwd(D, kinds, B):
    ops = _ops()
    pb = Problem(B, kinds, D, vocab=37, seed=D + B)
    f = pb.fused(DEV)
    E, S, fm, lr = ops.embed_fm_fwd(f["table"] if pb.F else None, f["table_lr"] if pb.F else None, f["rows"],
                                    pb.cat_pos, f["dense_x"], f["dense_w"], f["dense_w_lr"], pb.num_pos, f["bias"],
                                    B=B)
    Er, fmr, lrr, *_ = pb.oracle_forward()
    assert torch.equal(E.cpu(), Er), "gathered rows / x*w must be bit-exact"
    assert_close(S, Er.sum(1), what="S")
    # FM cancels (sum^2 - sum of squares): tolerance relative to the minuend
    assert_close(fm, fmr.reshape(-1), atol_scale=1e-5 * max(1.0, float((Er.sum(1) ** 2).sum(-1).max() / (fmr.abs().max() + 1e-30))), what="fm")
    assert_close(lr, lrr.reshape(-1), what="lr")


def test_embed_fm_fwd_optional_outputs():
    ops = _ops()
    pb = Problem(100, "ccnc", 16, seed=3)
    f = pb.fused(DEV)
    E, S, fm, lr = ops.embed_fm_fwd(f["table"], None, f["rows"], pb.cat_pos, f["dense_x"], f["dense_w"], None,
                                    pb.num_pos, None, want_E=False, want_S=False, want_fm=True, want_lr=False)
    assert E is None and S is None and lr is None
    _, fmr, *_ = pb.oracle_forward()
    assert_close(fm, fmr.reshape(-1), atol_scale=1e-4, what="fm only")


def test_embed_fm_out_of_range_rows_read_zero():
    ops = _ops()
    pb = Problem(64, "cc", 16, seed=5)
    f = pb.fused(DEV)
    rows = f["rows"].clone()
    rows[3, 1] = pb.R + 7
    rows[9, 0] = -4
    E, S, fm, lr = ops.embed_fm_fwd(f["table"], f["table_lr"], rows, pb.cat_pos, None, None, None, [], f["bias"])
    assert float(E[3, 1].abs().sum()) == 0.0 and float(E[9, 0].abs().sum()) == 0.0


# ----------------------------------------------------------------------------- K3 backward
@pytest.mark.parametrize("D", [4, 16, 64, 128, 10, 40])
@pytest.mark.parametrize("kinds,B", [("nncccccn", 257), ("ccc", 64), ("nn", 33), ("c" * 26 + "n" * 13, 1000)])
@pytest.mark.parametrize("use_E", [True, False])
def test_embed_fm_bwd(D, kinds, B, use_E):
    ops = _ops()
    pb = Problem(B, kinds, D, vocab=23, seed=7 * D + B)
    f = pb.fused(DEV)
    g = torch.Generator().manual_seed(99)
    dE = torch.randn(B, pb.F + pb.Fn, D, generator=g)
    d_fm = torch.randn(B, generator=g)
    d_lr = torch.randn(B, generator=g)
    E, S, fm, lr = ops.embed_fm_fwd(f["table"] if pb.F else None, f["table_lr"] if pb.F else None, f["rows"],
                                    pb.cat_pos, f["dense_x"], f["dense_w"], f["dense_w_lr"], pb.num_pos, f["bias"], B=B)
    gt = torch.zeros_like(f["table"])
    gt1 = torch.zeros_like(f["table_lr"])
    gw = torch.zeros_like(f["dense_w"]) if pb.Fn else None
    gw1 = torch.zeros_like(f["dense_w_lr"]) if pb.Fn else None
    gb = torch.zeros(1, device=DEV)
    ops.embed_fm_bwd(f["table"] if pb.F else None, f["rows"], pb.cat_pos, pb.pad_row, f["dense_x"], f["dense_w"],
                     pb.num_pos, E if use_E else None, S, dE.to(DEV), d_fm.to(DEV), d_lr.to(DEV),
                     gt if pb.F else None, gt1 if pb.F else None, gw, gw1, gb, D, pb.R, B=B)
    r_gt, r_gt1, r_gw, r_gw1, r_gb = pb.oracle_grads(dE, d_fm, d_lr)
    assert_close(gt, r_gt, atol_scale=2e-5, what="g_table")
    assert_close(gt1, r_gt1, atol_scale=2e-5, what="g_table_lr")
    if pb.Fn:
        assert_close(gw, r_gw, atol_scale=2e-5, what="g_dense_w")
        assert_close(gw1, r_gw1, atol_scale=2e-5, what="g_dense_w_lr")
    assert_close(gb, r_gb, atol_scale=2e-5, what="g_bias")
    for f_idx, pr in enumerate(pb.pad_row):
        assert float(gt[pr].abs().sum()) == 0.0 and float(gt1[pr]) == 0.0, "padding row must get zero grad"


def test_embed_fm_bwd_known_answer():
    """W=arange(10).view(5,2), idx=[[1,1,0],[4,0,0]], loss=sum ->
    W.grad=[[0,0],[2,2],[0,0],[0,0],[1,1]] (pad row zero, duplicates accumulate).  Three slots
    sharing one table = three slots with the same field offset."""
    ops = _ops()
    table = torch.arange(10, dtype=torch.float32, device=DEV).view(5, 2)
    rows = torch.tensor([[1, 1, 0], [4, 0, 0]], dtype=torch.int32, device=DEV)
    gt = torch.zeros_like(table)
    dE = torch.ones(2, 3, 2, device=DEV)
    ops.embed_fm_bwd(table, rows, [0, 1, 2], [0, 0, 0], None, None, [], None, None, dE, None, None,
                     gt, None, None, None, None, 2, 5)
    assert gt.cpu().tolist() == [[0, 0], [2, 2], [0, 0], [0, 0], [1, 1]]


# --------------------------------------------------------------------------- a5 gather / scatter
@pytest.mark.parametrize("D", [4, 16, 64, 128, 10, 300])
def test_gather_scatter(D):
    ops = _ops()
    g = torch.Generator().manual_seed(D)
    V, N = 101, 3001
    table = torch.randn(V, D, generator=g)
    ids = torch.randint(0, V, (N,), generator=g)
    out = ops.gather_rows(table.to(DEV), ids.int().to(DEV))
    assert torch.equal(out.cpu(), torch.from_numpy(oracle.gather_rows_numpy(table.numpy(), ids.numpy())))
    grad = torch.randn(N, D, generator=g)
    gt = torch.zeros(V, D, device=DEV)
    ops.scatter_add_rows(grad.to(DEV), ids.int().to(DEV), 0, gt)
    ref = oracle.embedding_dense_backward_numpy(grad.numpy(), ids.numpy(), V, padding_idx=0)
    assert_close(gt, torch.from_numpy(ref), atol_scale=2e-5, what="scatter")
    assert float(gt[0].abs().sum()) == 0.0


# ------------------------------------------------------------------------------- a9 pooled gather
@pytest.mark.parametrize("D", [8, 16, 64, 10])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("L", [1, 7, 20, 50])
def test_pooled_gather(D, mode, L):
    ops = _ops()
    g = torch.Generator().manual_seed(D * 100 + L)
    V, B = 77, 130
    table = torch.randn(V, D, generator=g)
    table[0] = 0
    ids = torch.randint(1, V, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    ids[torch.arange(L)[None, :] >= lens[:, None]] = 0          # right-padded with the pad id
    emb = torch.nn.functional.embedding(ids, table, padding_idx=0)
    ref = oracle.masked_average_pooling(emb) if mode else oracle.masked_sum_pooling(emb)
    out, cnt = ops.pooled_gather_fwd(table.to(DEV), ids.int().to(DEV), mode)
    assert_close(out, ref, what="pooled fwd")
    if mode:
        assert torch.equal(cnt.cpu(), lens.float())
    # backward
    go = torch.randn(B, D, generator=g)
    tl = table.clone().requires_grad_(True)
    e2 = torch.nn.functional.embedding(ids, tl, padding_idx=0)
    (((oracle.masked_average_pooling(e2) if mode else oracle.masked_sum_pooling(e2)) * go).sum()).backward()
    gt = torch.zeros(V, D, device=DEV)
    ops.pooled_gather_bwd(go.to(DEV), ids.int().to(DEV), cnt, 0, gt, mode)
    assert_close(gt, tl.grad, atol_scale=2e-5, what="pooled bwd")


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("D,L", [(4, 1), (8, 7), (64, 200), (128, 33), (16, 20), (10, 5), (1, 9)])
def test_pool_materialised(D, L, mode):
    """rbx_pool_fwd/bwd (Masked{Average,Sum}Pooling called on a [B,L,D] tensor, sequence.py:4-20 / pooling.py:22-40):
    zero rows do not count, an explicit mask overrides the row-sum test, the gradient reaches every position."""
    ops = _ops()
    g = torch.Generator().manual_seed(D * 100 + L + mode)
    B = 301
    emb = torch.randn(B, L, D, generator=g)
    emb[torch.rand(B, L, generator=g) < 0.3] = 0.0           # padded positions
    emb[5] = 0.0                                             # a sample with no valid position at all
    ref = oracle.masked_average_pooling(emb) if mode else oracle.masked_sum_pooling(emb)
    out, cnt = ops.pool_fwd(emb.to(DEV), None, mode)
    assert_close(out, ref, what="pool fwd")
    if mode == 1:
        assert torch.equal(cnt.cpu(), (emb.sum(-1) != 0).float().sum(-1))
        mask = torch.rand(B, L, generator=g) < 0.5
        out_m, cnt_m = ops.pool_fwd(emb.to(DEV), mask.to(DEV), 1)
        assert_close(out_m, oracle.masked_average_pooling(emb, mask), what="pool fwd mask")
        assert torch.equal(cnt_m.cpu(), mask.float().sum(-1))
    go = torch.randn(B, D, generator=g)
    e2 = emb.clone().requires_grad_(True)
    ((oracle.masked_average_pooling(e2) if mode else oracle.masked_sum_pooling(e2)) * go).sum().backward()
    d_emb = ops.pool_bwd(go.to(DEV), cnt, (B, L, D), mode)
    keep = torch.ones(B, dtype=torch.bool)
    keep[5] = False                       # 0/1e-12: the reference's own gradient there is g*1e12, compare the rest
    assert_close(d_emb.cpu()[keep], e2.grad[keep], what="pool bwd")


def test_pooled_known_answer():
    """MaskedAveragePooling on W=arange(10).view(5,2) (row 0 zeroed as the
    reference relies on), idx=[[1,1,0],[4,0,0]] -> [[2,3],[8,9]] (two, resp. one, non-pad rows)."""
    ops = _ops()
    W = torch.arange(10, dtype=torch.float32).view(5, 2)
    W[0] = 0
    ids = torch.tensor([[1, 1, 0], [4, 0, 0]], dtype=torch.int32)
    out, cnt = ops.pooled_gather_fwd(W.to(DEV), ids.to(DEV), 1)
    assert out.cpu().tolist() == [[2.0, 3.0], [8.0, 9.0]] and cnt.cpu().tolist() == [2.0, 1.0]


# ------------------------------------------------------------------------------------ a10 rowdot
@pytest.mark.parametrize("D", [16, 64, 128, 10, 48])
@pytest.mark.parametrize("K", [1, 11, 40])
def test_rowdot(D, K):
    ops = _ops()
    g = torch.Generator().manual_seed(D + K)
    B = 300
    u = torch.randn(B, D, generator=g, requires_grad=True)
    v = torch.randn(B * K, D, generator=g, requires_grad=True)
    y = oracle.two_tower_score(u, v)
    yd = ops.rowdot_fwd(u.detach().to(DEV), v.detach().to(DEV))
    assert_close(yd, y, what="rowdot fwd")
    if K == 1:
        assert_close(yd.reshape(-1), oracle.dssm_score(u, v), what="dssm")
    dy = torch.randn(B, K, generator=g)
    (y * dy).sum().backward()
    du, dv = ops.rowdot_bwd(u.detach().to(DEV), v.detach().to(DEV), dy.to(DEV))
    assert_close(du, u.grad, what="du")
    assert_close(dv.reshape(B * K, D), v.grad, what="dv")


# ----------------------------------------------------------------------------------- a6 interact
KAT = {  # SURVEY.md section 4, e = arange(24).view(2,3,4)
    "product_sum": [314, 3626],
    "bi_interaction": [32, 59, 92, 131, 752, 851, 956, 1067],
    "inner_product": [38, 62, 214, 950, 1166, 1510],
}


@pytest.mark.parametrize("mode", list(KAT))
def test_interact_known_answers(mode):
    ops = _ops()
    E = torch.arange(24, dtype=torch.float32, device=DEV).view(2, 3, 4)
    out = ops.interact_fwd(E, ops.MODES[mode])
    assert out.reshape(-1).cpu().tolist() == [float(x) for x in KAT[mode]]
    if mode == "product_sum":
        dE = ops.interact_bwd(E, torch.ones(2, 1, device=DEV), 0)
        assert dE[0].cpu().tolist() == [[12, 14, 16, 18], [8, 10, 12, 14], [4, 6, 8, 10]]


def test_interact_elementwise_known_answer():
    ops = _ops()
    E = torch.arange(24, dtype=torch.float32, device=DEV).view(2, 3, 4)
    out = ops.interact_fwd(E, 3)
    assert out[0].reshape(-1).cpu().tolist() == [0, 5, 12, 21, 0, 9, 20, 33, 32, 45, 60, 77]


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("F,D", [(39, 16), (5, 10), (2, 64), (13, 128), (7, 1), (8, 4), (40, 8), (33, 32), (4, 16), (70, 16)])
def test_interact_random(mode, F, D):
    ops = _ops()
    name = [k for k, v in ops.MODES.items() if v == mode][0]
    g = torch.Generator().manual_seed(F * D + mode)
    B = 67
    E = torch.randn(B, F, D, generator=g, requires_grad=True)
    ref = oracle.inner_product_interaction(E, name)
    out = ops.interact_fwd(E.detach().to(DEV), mode)
    scale = float((E.detach().sum(1) ** 2).sum(-1).max()) if mode <= 1 else 1.0
    assert_close(out.reshape(ref.shape), ref, atol_scale=1e-5 * max(1.0, scale / float(ref.abs().max())), what=name)
    dout = torch.randn(ref.shape, generator=g)
    (ref * dout).sum().backward()
    dE = ops.interact_bwd(E.detach().to(DEV), dout.to(DEV).contiguous(), mode)
    assert_close(dE, E.grad, atol_scale=2e-5, what=name + " bwd")


@pytest.mark.parametrize("B", [1, 3, 67, 4099])
@pytest.mark.parametrize("F", [25, 32, 33, 39, 40])
def test_interact_inner_product_tensor_core_path(F, B, monkeypatch):
    """mode 2 at D = 16, 25 <= F <= 40 runs as warp-level mma.sync in 3xTF32 (csrc/interact.cu, k_ip_fwd_mma / k_ip_bwd_mma):
    fp32-level accuracy against the oracle (1e-5 contract) and against a float64 product (well inside it), every padded
    row / column shape (F = 25 ... 40), batches that end inside a 16-byte chunk of the gradient stream (B * P % 4 != 0), and
    agreement with the SIMT engine (RBX_IP_ENGINE=0)."""
    ops = _ops()
    g = torch.Generator().manual_seed(100 * F + B)
    E = torch.randn(B, F, 16, generator=g, requires_grad=True)
    ref = oracle.inner_product_interaction(E, "inner_product")
    dout = torch.randn(ref.shape, generator=g)
    (ref * dout).sum().backward()
    Ed, dd = E.detach().to(DEV), dout.to(DEV).contiguous()
    monkeypatch.setenv("RBX_IP_ENGINE", "1")
    out = ops.interact_fwd(Ed, 2)
    dE = ops.interact_bwd(Ed, dd, 2)
    assert_close(out.reshape(ref.shape), ref, what="inner_product (mma)")
    assert_close(dE, E.grad, atol_scale=2e-5, what="inner_product bwd (mma)")
    iu = torch.triu_indices(F, F, 1)
    E64 = E.detach().double()
    want = torch.bmm(E64, E64.transpose(1, 2))[:, iu[0], iu[1]]
    assert float((out.double().cpu() - want).abs().max()) <= 3e-6 * float(want.abs().max()), "3xTF32 keeps fp32-level accuracy"
    monkeypatch.setenv("RBX_IP_ENGINE", "0")
    out0 = ops.interact_fwd(Ed, 2)
    dE0 = ops.interact_bwd(Ed, dd, 2)
    assert_close(out, out0, what="mma vs SIMT engine")
    assert_close(dE, dE0, atol_scale=2e-5, what="mma vs SIMT engine (bwd)")


@pytest.mark.parametrize("mode", [2, 3])
def test_interact_pairs_many_samples_per_warp(mode):
    """B far above the resident warp count: every warp of the warp-per-sample kernels reuses its shared-memory slice."""
    ops = _ops()
    name = [k for k, v in ops.MODES.items() if v == mode][0]
    g = torch.Generator().manual_seed(mode)
    B, F, D = 50021, 6, 8
    E = torch.randn(B, F, D, generator=g, requires_grad=True)
    ref = oracle.inner_product_interaction(E, name)
    out = ops.interact_fwd(E.detach().to(DEV), mode)
    assert_close(out.reshape(ref.shape), ref, what=name)
    dout = torch.randn(ref.shape, generator=g)
    (ref * dout).sum().backward()
    dE = ops.interact_bwd(E.detach().to(DEV), dout.to(DEV).contiguous(), mode)
    assert_close(dE, E.grad, atol_scale=2e-5, what=name + " bwd")


@pytest.mark.parametrize("order", [1, 2, 5])
@pytest.mark.parametrize("F,D", [(39, 16), (6, 8), (3, 10), (5, 128), (2, 4), (7, 1)])
def test_power_sums(F, D, order):
    """f4: rbx_power_sums_fwd/bwd against the oracle's Q = Q * X; Q.sum(1) chain and its autograd."""
    ops = _ops()
    g = torch.Generator().manual_seed(F * D + order)
    B = 203
    E = (torch.randn(B, F, D, generator=g) * 0.8).requires_grad_(True)
    ref = oracle.power_sums(E, order)
    P = ops.power_sums_fwd(E.detach().to(DEV), order)
    assert_close(P, ref, what="power sums")
    dP = torch.randn(B, order, D, generator=g)
    (ref * dP).sum().backward()
    dE = ops.power_sums_bwd(E.detach().to(DEV), dP.to(DEV))
    assert_close(dE, E.grad, atol_scale=2e-5, what="power sums bwd")
    from recbox_b200 import RbxError
    with pytest.raises(RbxError):
        ops.power_sums_fwd(E.detach().to(DEV), 6)


# ------------------------------------------------------------------------------------ (e) shard
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("N", [1, 255, 2048, 2049, 100003])
def test_shard_route(world, N):
    ops = _ops()
    rng = np.random.default_rng(N + world)
    rows = rng.integers(0, 1_000_000, size=N).astype(np.int32)
    send_r, counts_r, pos_r = oracle.shard_route(rows, world)
    send, pos, counts = ops.shard_route(torch.from_numpy(rows).to(DEV), world)
    assert np.array_equal(counts.cpu().numpy(), counts_r)
    assert np.array_equal(pos.cpu().numpy(), pos_r)
    assert np.array_equal(send.cpu().numpy(), send_r)
    payload = torch.from_numpy(rng.standard_normal((N, 16)).astype(np.float32)).to(DEV)
    sent = ops.shard_permute(payload, pos)
    back = ops.shard_unroute(sent, pos)
    assert torch.equal(back, payload)
    assert np.array_equal(back.cpu().numpy(), oracle.shard_unroute(sent.cpu().numpy(), pos_r))


# ---------------------------------------------------------------------------------------- a1
def test_split_batch_and_pack_columns():
    ops = _ops()
    pb = Problem(999, "ncncccnnc", 16, vocab=1000, seed=11)
    M = pb.batch_matrix(label=True)
    n_cols = M.shape[1]
    kinds, slots, ci, ni = [], [], 0, 0
    for k in pb.kinds:
        kinds.append(1 if k == "c" else 2)
        slots.append(ci if k == "c" else ni)
        ci, ni = ci + (k == "c"), ni + (k != "c")
    kinds.append(3)
    slots.append(0)
    cat_cols = [i for i, k in enumerate(pb.kinds) if k == "c"]
    num_cols = [i for i, k in enumerate(pb.kinds) if k != "c"]
    ids_r, dense_r, label_r = oracle.split_batch_numpy(M.numpy(), cat_cols, num_cols, n_cols - 1)
    rows, dense, label = ops.split_batch(M.to(DEV), kinds, slots, pb.field_off, pb.F, pb.Fn)
    assert np.array_equal(rows.cpu().numpy(), ids_r + np.asarray(pb.field_off)[None, :])
    assert np.array_equal(dense.cpu().numpy(), dense_r)
    assert np.array_equal(label.cpu().numpy(), label_r)
    # dict-of-columns form: strided views of the device matrix and separate tensors of mixed dtype
    Md = M.to(DEV)
    cols = [Md[:, c] for c in cat_cols]
    cols[1] = cols[1].long().contiguous()
    cols[2] = cols[2].int().contiguous()
    packed = ops.pack_columns(cols, add=pb.field_off, as_rows=True)
    assert torch.equal(packed, rows)
    dcols = [Md[:, c] for c in num_cols]
    dcols[0] = dcols[0].float().contiguous()
    assert torch.equal(ops.pack_columns(dcols, as_rows=False), dense)


@pytest.mark.parametrize("B", [1, 31, 33, 70001])
def test_split_batch_ragged_and_strided(B):
    """Row counts around the 32-row block size, an ignored column, and a batch that is a column-slice view (ld > n_cols)."""
    ops = _ops()
    rng = np.random.default_rng(B)
    wide = np.concatenate([rng.integers(0, 50, (B, 3)).astype(np.float64), rng.standard_normal((B, 2)),
                           rng.integers(0, 2, (B, 1)).astype(np.float64), rng.standard_normal((B, 4))], 1)
    Md = torch.from_numpy(wide).to(DEV)[:, :7]                      # stride(0) = 10, 7 visible columns
    kinds, slots = [1, 1, 0, 2, 2, 3, 0], [1, 0, 0, 0, 1, 0, 0]
    rows, dense, label = ops.split_batch(Md, kinds, slots, [100, 200], 2, 2)
    assert np.array_equal(rows.cpu().numpy(), np.stack([wide[:, 1] + 100, wide[:, 0] + 200], 1).astype(np.int32))
    assert np.array_equal(dense.cpu().numpy(), wide[:, 3:5].astype(np.float32))
    assert np.array_equal(label.cpu().numpy(), wide[:, 5].astype(np.float32))


# --------------------------------------------------------------------------------------- a12
def test_clip_and_adam():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    n = 100_003
    w = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * 0.05 for _ in range(3)]
    m, v = torch.zeros(n), torch.zeros(n)
    wd, md, vd = w.to(DEV), m.to(DEV), v.to(DEV)
    coef_d = torch.empty(1, device=DEV)
    norm_d = torch.empty(1, device=DEV)
    for step, gr in enumerate(grads, 1):
        total, coef = oracle.clip_grad_norm([gr], 10.0)
        oracle.adam_step(w, gr * coef, m, v, step)
        acc = torch.zeros(1, dtype=torch.float64, device=DEV)
        gd = gr.to(DEV)
        ops.sqnorm_(gd, acc)
        ops.clip_coef(acc, 10.0, coef_d, norm_d)
        assert_close(norm_d, total.reshape(1), what="norm")
        assert_close(coef_d, coef.reshape(1), what="coef")
        ops.adam_dense_(wd, gd, md, vd, step, clip=coef_d)
        assert_close(wd, w, rtol=1e-5, atol_scale=1e-6, what="adam w step %d" % step)
        assert_close(md, m, what="adam m")
        assert_close(vd, v, what="adam v")


def test_errors_are_loud():
    ops = _ops()
    from recbox_b200 import RbxError
    with pytest.raises(RbxError):
        ops.gather_rows(torch.zeros(4, 4), torch.zeros(2, dtype=torch.int32))       # CPU tensors
    with pytest.raises(RbxError):
        ops.interact_fwd(torch.zeros(2, 3, 4, device=DEV), 7)                        # unknown mode


# ------------------------------------------------------------------------------ a1: range guard, compact ids, zero fill
def test_out_of_vocabulary_ids_never_alias_a_neighbour_table():
    """nn.Embedding raises IndexError for an id outside [0, vocab); in the fused table id + field offset would silently read
    (and scatter into) ANOTHER feature's rows.  With the vocabulary sizes given, such ids become row -1 (zero row, no
    gradient) and are counted."""
    ops = _ops()
    B = 257
    rng = np.random.default_rng(3)
    vocab, off = [10, 20, 5], [0, 10, 30]
    M = np.stack([rng.integers(0, v, B) for v in vocab] + [rng.standard_normal(B)], 1).astype(np.float64)
    M[5, 0], M[6, 1], M[7, 2], M[8, 1] = 10, -1, 5, 1e6
    n_bad = torch.zeros(1, dtype=torch.int32, device=DEV)
    rows, dense, _ = ops.split_batch(torch.from_numpy(M).to(DEV), [1, 1, 1, 2], [0, 1, 2, 0], off, 3, 1, want_label=False,
                                     field_rows=vocab, n_bad=n_bad)
    want = M[:, :3].astype(np.int64) + np.asarray(off)[None]
    want[5, 0] = want[6, 1] = want[7, 2] = want[8, 1] = -1
    assert np.array_equal(rows.cpu().numpy(), want) and int(n_bad) == 4
    unchecked, _, _ = ops.split_batch(torch.from_numpy(M).to(DEV), [1, 1, 1, 2], [0, 1, 2, 0], off, 3, 1, want_label=False)
    assert int(unchecked[5, 0]) == 10                               # (the documented unchecked behaviour)
    n_bad.zero_()
    Md = torch.from_numpy(M).to(DEV)
    packed = ops.pack_columns([Md[:, 0], Md[:, 1].long().contiguous(), Md[:, 2].int().contiguous()], add=off, as_rows=True,
                              vocab=vocab, n_bad=n_bad)
    assert np.array_equal(packed.cpu().numpy(), want) and int(n_bad) == 4


@pytest.mark.parametrize("B,F", [(1, 1), (7, 3), (1000, 26), (65536, 26), (333, 5)])
def test_unpack_ids_u16(B, F):
    ops = _ops()
    rng = np.random.default_rng(B + F)
    ids = rng.integers(0, 65536, (B, F)).astype(np.uint16)
    off = (np.arange(F) * 65536).tolist()
    t = torch.from_numpy(ids.view(np.int16)).to(DEV)
    rows = ops.unpack_ids_u16(t, off)
    assert np.array_equal(rows.cpu().numpy(), ids.astype(np.int64) + np.asarray(off)[None])
    if B > 2:                                                        # a view that is not 16-byte aligned
        rows2 = ops.unpack_ids_u16(t[1:].contiguous()[:], off)
        assert np.array_equal(rows2.cpu().numpy(), (ids.astype(np.int64) + np.asarray(off)[None])[1:])


@pytest.mark.parametrize("n,skew", [(1, 0), (5, 1), (1025, 3), (17000003, 2)])
def test_zero_fill(n, skew):
    ops = _ops()
    buf = torch.full((n + skew + 8,), 7.0, device=DEV)
    ops.zero_(buf[skew:skew + n])
    assert float(buf[skew:skew + n].abs().sum()) == 0.0
    assert float(buf[:skew].sum()) == 7.0 * skew and float(buf[skew + n:].sum()) == 7.0 * 8


def test_ops_run_on_the_tensors_device_not_the_current_one():
    """The reference's get_device(gpu) hands out cuda:<gpu> without set_device (torch_utils.py:37-42): tensors off the
    current device are normal.  Needs two GPUs; on one it checks the mixed-device error only."""
    ops = _ops()
    pb = Problem(300, "ccn", 16, vocab=50, seed=2)
    with pytest.raises(RbxError):
        ops.zero_(torch.zeros(4))                                    # CPU tensor: no CPU path
    if torch.cuda.device_count() < 2:
        return
    f1 = pb.fused("cuda:1")
    assert torch.cuda.current_device() == 0
    E1, S1, fm1, lr1 = ops.embed_fm_fwd(f1["table"], f1["table_lr"], f1["rows"], pb.cat_pos, f1["dense_x"], f1["dense_w"],
                                        f1["dense_w_lr"], pb.num_pos, f1["bias"])
    f0 = pb.fused("cuda:0")
    E0, S0, fm0, lr0 = ops.embed_fm_fwd(f0["table"], f0["table_lr"], f0["rows"], pb.cat_pos, f0["dense_x"], f0["dense_w"],
                                        f0["dense_w_lr"], pb.num_pos, f0["bias"])
    assert E1.device.index == 1 and torch.equal(E1.cpu(), E0.cpu()) and torch.equal(fm1.cpu(), fm0.cpu())
    with pytest.raises(RbxError):
        ops.embed_fm_fwd(f1["table"], f1["table_lr"], f0["rows"], pb.cat_pos, f1["dense_x"], f1["dense_w"],
                         f1["dense_w_lr"], pb.num_pos, f1["bias"])


# ----------------------------------------------------------------------------- deterministic scatter
@pytest.mark.parametrize("D", [1, 10, 16, 64, 200])
def test_scatter_add_rows_deterministic_is_bit_exact_and_reproducible(D):
    """rbx_segment_sum_rows: every row's gradient rows are summed in position order, so the result equals numpy's sequential
    np.add.at BIT FOR BIT and does not change from run to run (the atomic kernel agrees within 1e-5 only)."""
    ops = _ops()
    rng = np.random.default_rng(D)
    N, R = 20000, 300                                   # ~67 contributions per row: the order of the adds matters
    ids = rng.integers(0, R, N).astype(np.int32)
    ids[::97] = 7                                       # a hot row
    g = rng.standard_normal((N, D)).astype(np.float32)
    want = np.zeros((R, D), np.float32)
    keep = ids != 5
    np.add.at(want, ids[keep], g[keep])                 # sequential, in position order, fp32
    outs = []
    for _ in range(2):
        gt = torch.zeros(R, D, device=DEV)
        ops.scatter_add_rows_deterministic(torch.from_numpy(g).to(DEV), torch.from_numpy(ids).to(DEV), 5, gt)
        outs.append(gt.cpu())
    assert torch.equal(outs[0], outs[1]), "not reproducible"
    assert torch.equal(outs[0], torch.from_numpy(want)), "not the position-order sum"
    assert float(outs[0][5].abs().sum()) == 0.0
    gt = torch.zeros(R, D, device=DEV)
    ops.scatter_add_rows(torch.from_numpy(g).to(DEV), torch.from_numpy(ids).to(DEV), 5, gt)
    assert_close(gt, outs[0], atol_scale=1e-5, what="atomic vs deterministic")


def test_embed_fm_bwd_deterministic_matches_the_atomic_kernel():
    ops = _ops()
    B, D = 500, 16
    pb = Problem(B, "nncccccn", D, vocab=23, seed=77, zipf=1.3)
    f = pb.fused(DEV)
    g = torch.Generator().manual_seed(9)
    dE = torch.randn(B, pb.F + pb.Fn, D, generator=g).to(DEV)
    d_fm, d_lr = torch.randn(B, generator=g).to(DEV), torch.randn(B, generator=g).to(DEV)
    E, S, fm, lr = ops.embed_fm_fwd(f["table"], f["table_lr"], f["rows"], pb.cat_pos, f["dense_x"], f["dense_w"], f["dense_w_lr"],
                                    pb.num_pos, f["bias"])
    rt, rt1 = torch.zeros_like(f["table"]), torch.zeros_like(f["table_lr"])
    rw, rw1, rb = torch.zeros_like(f["dense_w"]), torch.zeros_like(f["dense_w_lr"]), torch.zeros(1, device=DEV)
    ops.embed_fm_bwd(f["table"], f["rows"], pb.cat_pos, pb.pad_row, f["dense_x"], f["dense_w"], pb.num_pos, E, S, dE, d_fm, d_lr,
                     rt, rt1, rw, rw1, rb, D, pb.R)
    outs = []
    for _ in range(2):
        gt, gt1 = torch.zeros_like(f["table"]), torch.zeros_like(f["table_lr"])
        ops.embed_fm_bwd_deterministic(f["table"], f["rows"], pb.cat_pos, pb.pad_row, E, S, dE, d_fm, d_lr, gt, gt1)
        outs.append((gt.cpu(), gt1.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "not reproducible"
    assert_close(outs[0][0], rt, atol_scale=2e-5, what="g_table")
    assert_close(outs[0][1], rt1, atol_scale=2e-5, what="g_table_lr")
    for p in pb.pad_row:
        assert float(outs[0][0][p].abs().sum()) == 0.0
