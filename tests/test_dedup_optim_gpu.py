"""a14 (sorted unique / inverse / first occurrence) and f1 (touched-rows clip + optimizers) through the
C ABI against the CPU oracle.  Integer outputs bit-exact; optimizer state within 1e-5 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from helpers import assert_close, oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ops():
    from recbox_b200 import ops
    return ops


# --------------------------------------------------------------------------------------- a14
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
@pytest.mark.parametrize("n,vocab", [(1, 1), (7, 5), (90112, 10_000_001), (4096, 31), (65537, 65536), (300_000, 2_000_003),
                                     (1000, 1 << 22), (1_703_936, 1_000_012), (250_003, 1_000_012), (250_002, 1_000_012),
                                     (33, 33), (100_000, 4097)])
def test_unique_ids_matches_oracle(dtype, n, vocab):
    ops = _ops()
    rng = np.random.default_rng(n + vocab)
    ids = rng.integers(0, vocab, size=n)
    ids[: n // 3] = ids[n // 2: n // 2 + n // 3]            # plenty of repeats
    uniq, first, inv = ops.unique_ids(torch.from_numpy(ids).to(dtype).to(DEV), vocab)
    ur, fr, ir = oracle.unique_items(ids)
    assert uniq.dtype == dtype and inv.dtype == dtype and first.dtype == torch.int64
    assert np.array_equal(uniq.cpu().numpy(), ur)
    assert np.array_equal(first.cpu().numpy(), fr)
    assert np.array_equal(inv.cpu().numpy(), ir)


def test_unique_ids_zipf_and_boundaries():
    ops = _ops()
    vocab = 1_000_001
    rng = np.random.default_rng(1)
    ids = np.minimum(rng.zipf(1.05, size=200_000), vocab - 1)
    ids[:4] = [0, vocab - 1, 31, 32]                          # word / vocabulary edges
    uniq, first, inv = ops.unique_ids(torch.from_numpy(ids).to(DEV).view(-1, 8), vocab)   # any shape flattens
    ur, fr, ir = oracle.unique_items(ids)
    assert np.array_equal(uniq.cpu().numpy(), ur) and np.array_equal(first.cpu().numpy(), fr)
    assert np.array_equal(inv.cpu().numpy(), ir)
    # size-independent properties: uniq strictly ascending; uniq[inverse] reproduces the input
    assert bool((uniq[1:] > uniq[:-1]).all())
    assert torch.equal(uniq[inv], torch.from_numpy(ids).to(DEV))


def test_unique_ids_empty_and_out_of_range():
    ops = _ops()
    from recbox_b200 import RbxError
    uniq, first, inv = ops.unique_ids(torch.zeros(0, dtype=torch.int64, device=DEV), 100)
    assert uniq.numel() == 0 and first.numel() == 0 and inv.numel() == 0
    bad = torch.tensor([3, 100, -1, 7], device=DEV)
    with pytest.raises(RbxError):
        ops.unique_ids(bad, 100)
    uniq, first, inv, n_out = ops.unique_ids(bad, 100, sync=False)
    assert n_out.tolist() == [2, 2] and inv.tolist() == [0, -1, -1, 1] and uniq[:2].tolist() == [3, 7]


def test_collate_unique_matches_reference_golden():
    """tests/golden/collate_unique.npz was minted by the UNMODIFIED reference collate_fn_unique
    (oracle/make_golden.py::gen_collate_unique)."""
    from recbox_b200.optim import collate_unique
    g = np.load(os.path.join(GOLD, "collate_unique.npz"))
    for k in range(int(g["n_cases"])):
        item = torch.from_numpy(g["item_indexes_%d" % k]).to(DEV)
        uniq, uidx, inv = collate_unique(item, int(g["vocab_%d" % k]))
        assert np.array_equal(uniq.cpu().numpy(), g["unique_%d" % k])
        assert np.array_equal(uidx.cpu().numpy(), g["unique_indexes_%d" % k])
        assert np.array_equal(inv.cpu().numpy(), g["inverse_indexes_%d" % k])


# ---------------------------------------------------------------------------------------- f1
def _touched_problem(R, D, n_ids, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, R, (n_ids,), generator=g)
    ids[: n_ids // 4] = ids[n_ids // 2: n_ids // 2 + n_ids // 4]
    grad = torch.zeros(R, D) if D > 0 else torch.zeros(R)
    rows = torch.unique(ids)
    grad[rows] = torch.randn((rows.numel(), D) if D > 0 else (rows.numel(),), generator=g) * 0.05
    w = torch.randn(grad.shape, generator=g)
    return ids, rows, grad, w


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "adam_rows", "sparse_adam"])
@pytest.mark.parametrize("R,D", [(5000, 16), (5000, 10), (777, 128), (100_003, 0), (64, 4)])
def test_optim_rows_matches_oracle(kind, R, D):
    ops = _ops()
    from recbox_b200.optim import TouchedRowsOptimizer
    ids, rows, grad, w = _touched_problem(R, D, 3000, R + D)
    state = {"m": torch.zeros_like(w), "v": torch.zeros_like(w)}
    wd, gd = w.to(DEV), grad.to(DEV)
    opt = TouchedRowsOptimizer([(wd, gd)], kind=kind, lr=0.01)
    uniq, n_rows = opt.touched(ids.to(torch.int32).to(DEV))
    assert int(n_rows) == rows.numel() and torch.equal(uniq[: rows.numel()].cpu().long(), rows)
    for step in range(1, 4):
        gstep = grad * (0.5 + step)
        total, coef = oracle.clip_grad_norm([gstep], 0.05)
        w = oracle.touched_rows_step(kind, w, gstep, state, step, lr=0.01, clip=float(coef))
        gd.copy_(gstep)
        clip = opt.step(ids.to(torch.int32).to(DEV), max_norm=0.05)
        assert_close(opt.clip.norm, total.reshape(1), what="norm")
        assert_close(clip, coef.reshape(1), what="coef")
        assert_close(wd, w, rtol=1e-5, atol_scale=1e-6, what="%s w step %d" % (kind, step))
        assert float(gd.abs().sum()) == 0.0, "consumed gradient rows must be cleared"
    _, _, m, v = opt.tables[0]
    if m is not None:
        assert_close(m, state["m"], what="m")
    if v is not None:
        assert_close(v, state["v"], what="v")
    untouched = torch.ones(R, dtype=torch.bool)
    untouched[rows] = False
    w0 = _touched_problem(R, D, 3000, R + D)[3]
    assert torch.equal(wd.cpu()[untouched], w0[untouched]), "untouched rows must not move"


def test_sgd_rows_equals_dense_sgd_on_fused_backward():
    """End to end on the hot path: fused backward into a zero gradient table, touched-rows SGD, gradient table
    left all-zero -- against dense SGD on the oracle's gradient."""
    from helpers import Problem
    ops = _ops()
    from recbox_b200.optim import TouchedRowsOptimizer
    pb = Problem(512, "c" * 6 + "n" * 3, 16, vocab=300, seed=11)
    f = pb.fused(DEV)
    E, S, fm, lr = ops.embed_fm_fwd(f["table"], f["table_lr"], f["rows"], pb.cat_pos, f["dense_x"], f["dense_w"],
                                    f["dense_w_lr"], pb.num_pos, f["bias"])
    g = torch.Generator().manual_seed(3)
    dE, d_fm, d_lr = torch.randn(512, 9, 16, generator=g), torch.randn(512, generator=g), torch.randn(512, generator=g)
    gt, gt1 = torch.zeros_like(f["table"]), torch.zeros_like(f["table_lr"])
    gw, gw1, gb = torch.zeros_like(f["dense_w"]), torch.zeros_like(f["dense_w_lr"]), torch.zeros(1, device=DEV)
    ops.embed_fm_bwd(f["table"], f["rows"], pb.cat_pos, pb.pad_row, f["dense_x"], f["dense_w"], pb.num_pos, E, S,
                     dE.to(DEV), d_fm.to(DEV), d_lr.to(DEV), gt, gt1, gw, gw1, gb, 16, pb.R)
    r = pb.oracle_grads(dE, d_fm, d_lr)
    w_ref = f["table"].cpu() - 0.1 * r[0]
    w1_ref = f["table_lr"].cpu() - 0.1 * r[1]
    opt = TouchedRowsOptimizer([(f["table"], gt), (f["table_lr"], gt1)], kind="sgd", lr=0.1)
    opt.step(f["rows"])
    assert_close(f["table"], w_ref, rtol=1e-5, atol_scale=2e-5, what="table after sgd")
    assert_close(f["table_lr"], w1_ref, rtol=1e-5, atol_scale=2e-5, what="table_lr after sgd")
    assert float(gt.abs().sum()) == 0.0 and float(gt1.abs().sum()) == 0.0
