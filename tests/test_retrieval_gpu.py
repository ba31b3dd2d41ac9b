"""f2 / f3 on cuda:0 through the C ABI: rbx_topk_ip against the oracle's exact search (index lists bit-exact wherever
scores are separated, score values within fp32 summation-order noise), rbx_rank_metrics and retrieval.evaluate_metrics
against the golden minted from the reference's core/metrics.py, rbx_sample_negatives by its distributional contract."""
import numpy as np
import pytest
import torch

from helpers import oracle
from test_retrieval_oracle import load_retrieval

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check_topk(q, c, k, chunk=None):
    from recbox_b200 import ops
    s_ref, i_ref = oracle.flat_ip_search(q, c, k)
    s, i = ops.topk_ip(torch.from_numpy(q).to(DEV), torch.from_numpy(c).to(DEV), k, chunk=chunk)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert s.shape == s_ref.shape and i.shape == i_ref.shape
    fin = np.isfinite(s_ref)
    assert np.array_equal(np.isfinite(s), fin) and np.all(i[~fin] == -1)
    np.testing.assert_allclose(s[fin], s_ref[fin], rtol=1e-5, atol=1e-5 * max(1.0, float(np.abs(s_ref[fin]).max())))
    assert np.all(np.diff(s, axis=1)[fin[:, 1:]] <= 0), "scores must come out descending"
    # the reported score IS the item's score (recomputed in float64), and no better item was left out
    full = q.astype(np.float64) @ c.astype(np.float64).T
    tol = 1e-5 * max(1.0, float(np.abs(full).max()))
    for u in range(len(q)):
        got = i[u][fin[u]]
        assert len(set(got.tolist())) == len(got), "duplicate item in a top-k list"
        np.testing.assert_allclose(full[u, got], s[u][fin[u]], atol=tol, rtol=0)
        if len(got):
            rest = np.delete(full[u], got)
            assert rest.size == 0 or rest.max() <= full[u, got].min() + tol
    # where neighbouring reference scores are separated by more than the noise the order is exactly the oracle's
    sep = np.ones_like(i_ref, dtype=bool)
    gap = np.abs(np.diff(s_ref, axis=1)) > 4 * tol
    sep[:, 1:] &= gap
    sep[:, :-1] &= gap
    sep &= fin
    assert np.array_equal(i[sep], i_ref[sep])


@pytest.mark.parametrize("U,N,D,k", [(1, 1, 4, 1), (3, 100, 8, 10), (130, 5000, 16, 50), (257, 20011, 64, 500),
                                     (64, 300, 128, 100), (5, 40, 12, 64), (1000, 33000, 64, 100)])
def test_topk_ip_matches_exact_search(U, N, D, k):
    rng = np.random.default_rng(U * 7 + N)
    _check_topk(rng.standard_normal((U, D)).astype(np.float32), rng.standard_normal((N, D)).astype(np.float32), k)


def test_topk_ip_many_small_passes_and_sorted_corpus():
    """chunk = 128 items per pass (hundreds of filter/select rounds); a corpus sorted by score ascending is the worst
    case for the threshold filter (every item beats the running k-th best)."""
    rng = np.random.default_rng(5)
    q = np.abs(rng.standard_normal((9, 8))).astype(np.float32)
    c = np.sort(np.abs(rng.standard_normal((3000, 1))), axis=0).astype(np.float32) * np.ones((1, 8), np.float32)
    _check_topk(q, c, 20, chunk=128)
    _check_topk(q, c[::-1].copy(), 20, chunk=256)
    _check_topk(rng.standard_normal((40, 16)).astype(np.float32), rng.standard_normal((10000, 16)).astype(np.float32), 1000, chunk=4096)


def test_topk_ip_ties_go_to_the_smaller_index():
    from recbox_b200 import ops
    c = np.zeros((700, 4), np.float32)
    c[:, 0] = 1.0                                          # every item scores the same
    q = np.ones((3, 4), np.float32)
    s, i = ops.topk_ip(torch.from_numpy(q).to(DEV), torch.from_numpy(c).to(DEV), 7, chunk=128)
    assert i.cpu().tolist() == [list(range(7))] * 3 and torch.all(s == 1.0)


def test_topk_errors_are_loud():
    from recbox_b200 import RbxError, ops
    q, c = torch.zeros(2, 6, device=DEV), torch.zeros(10, 6, device=DEV)
    with pytest.raises(RbxError):
        ops.topk_ip(q, c, 3)                                # D % 4 != 0 at the C ABI (FlatIPIndex pads)
    with pytest.raises(RbxError):
        ops.topk_ip(torch.zeros(2, 8, device=DEV), torch.zeros(10, 8, device=DEV), 2000)
    with pytest.raises(RbxError):
        ops.topk_ip(torch.zeros(2, 8), torch.zeros(10, 8), 3)


def test_rank_metrics_and_evaluate_metrics_match_reference():
    from recbox_b200 import ops, retrieval
    g, query, train, valid, metrics = load_retrieval()
    got = retrieval.evaluate_metrics(g["user"].astype(np.float64), g["item"].astype(np.float64), train, valid, query, metrics)
    np.testing.assert_allclose([got[m] for m in metrics], g["average"], rtol=1e-9, atol=1e-12)
    # per-user values through the two kernels
    kinds, ks = retrieval.parse_metrics(metrics)
    index = retrieval.FlatIPIndex(g["item"], dim=16)
    _, cand = index.search_device(g["user"], topk=500)
    tp, ti = retrieval.build_csr(train, query, DEV)
    vp, vi = retrieval.build_csr(valid, query, DEV)
    ranked, hit, out = ops.rank_metrics(cand, tp, ti, vp, vi, kinds, ks, kmax=50)
    np.testing.assert_allclose(out.cpu().numpy(), g["per_user"], rtol=1e-9, atol=1e-12)
    funcs = [eval(m, {}, {n: getattr(oracle, n) for n in ops.METRIC_KINDS}) for m in metrics]
    topk_items, _ = oracle.evaluate_block(g["user"], g["item"], query, train, valid, funcs, 50)
    assert np.array_equal(ranked.cpu().numpy(), topk_items)
    # FaissIndex drop-in surface
    s_np, i_np = index.search(g["user"][:5], topk=7)
    s_ref, i_ref = oracle.flat_ip_search(g["user"][:5], g["item"], 7)
    assert isinstance(s_np, np.ndarray) and np.array_equal(i_np, i_ref)
    np.testing.assert_allclose(s_np, s_ref, rtol=1e-5, atol=1e-5)
    # D not a multiple of 4 and l2_normalize go through the padded corpus
    idx2 = retrieval.FlatIPIndex(g["item"][:, :10], dim=10, l2_normalize=True)
    _, i2 = idx2.search(g["user"][:9, :10], topk=5)
    cn = g["item"][:, :10] / np.linalg.norm(g["item"][:, :10], axis=1, keepdims=True)
    un = g["user"][:9, :10] / np.linalg.norm(g["user"][:9, :10], axis=1, keepdims=True)
    assert np.array_equal(i2, oracle.flat_ip_search(un, cn, 5)[1])


def test_sample_negatives_contract():
    from recbox_b200 import ops
    n_items, n_q, negs = 1000, 4096, 20
    pos = torch.randint(0, n_items, (n_q,), device=DEV)
    out, gave_up = ops.sample_negatives(n_q, negs, n_items, seed=7, pos=pos)
    assert gave_up is None and out.shape == (n_q, 1 + negs) and out.dtype == torch.int64
    assert torch.equal(out[:, 0], pos)                                       # h5_generator.py:176-177 hstack
    neg = out[:, 1:]
    assert int(neg.min()) >= 0 and int(neg.max()) < n_items
    out2, _ = ops.sample_negatives(n_q, negs, n_items, seed=7, pos=pos)
    out3, _ = ops.sample_negatives(n_q, negs, n_items, seed=8, pos=pos)
    assert torch.equal(out, out2) and not torch.equal(out, out3)             # reproducible from the seed
    # uniformity: chi-square of 81 920 draws over 1000 bins (dof 999; 1200 is far beyond the 99.99th percentile)
    cnt = torch.bincount(neg.reshape(-1), minlength=n_items).double().cpu().numpy()
    exp = neg.numel() / n_items
    assert ((cnt - exp) ** 2 / exp).sum() < 1200
    ref = oracle.sampling_block(n_items, list(range(n_q)), negs, {}, seed=3)
    cnt_ref = np.bincount(ref.reshape(-1), minlength=n_items)
    assert ((cnt_ref - exp) ** 2 / exp).sum() < 1200                          # the reference's draws pass the same test
    # ignore_pos_items: a user's own items are never drawn; the rest stays uniform
    n_users = 50
    u2i = {u: sorted(set(np.random.default_rng(u).integers(0, n_items, 300).tolist())) for u in range(n_users)}
    ptr = np.cumsum([0] + [len(u2i[u]) for u in range(n_users)])
    items = np.concatenate([u2i[u] for u in range(n_users)])
    uq = torch.randint(0, n_users, (n_q,), device=DEV)
    out, gave_up = ops.sample_negatives(n_q, negs, n_items, seed=11, user_of_query=uq,
                                        pos_ptr=torch.from_numpy(ptr).to(DEV), pos_items=torch.from_numpy(items).to(DEV))
    assert out.shape == (n_q, negs) and int(gave_up) == 0
    o, uqc = out.cpu().numpy(), uq.cpu().numpy()
    for r in range(0, n_q, 37):
        assert not set(o[r].tolist()) & set(u2i[int(uqc[r])])
    u0 = o[uqc == 0].reshape(-1)
    allowed = np.setdiff1d(np.arange(n_items), u2i[0])
    c0 = np.bincount(u0, minlength=n_items)[allowed]
    e0 = len(u0) / len(allowed)
    assert ((c0 - e0) ** 2 / e0).sum() < 2.0 * len(allowed)
    # a user who interacted with everything cannot be served: the counter says so instead of looping forever
    full_ptr = torch.tensor([0, 16], device=DEV)
    full_items = torch.arange(16, device=DEV)
    out, gave_up = ops.sample_negatives(8, 4, 16, seed=1, pos_ptr=full_ptr, pos_items=full_items,
                                        user_of_query=torch.zeros(8, dtype=torch.int64, device=DEV))
    assert int(gave_up) == 32


def test_epoch_negative_sampler_mirrors_train_generator():
    from recbox_b200.loader import EpochNegativeSampler
    rng = np.random.default_rng(2)
    n_items, N = 500, 3000
    query = rng.integers(0, 40, N)
    pos = rng.integers(0, n_items, N)
    u2i = {}
    for qi, it in zip(query, pos):
        u2i.setdefault(int(qi), []).append(int(it))            # get_user2items_dict, h5_generator.py:37-42
    smp = EpochNegativeSampler(n_items, query, pos, 6, user2items_dict=u2i, ignore_pos_items=True, seed=5)
    a = smp.sample()
    b = smp.sample()
    assert a.shape == (N, 7) and torch.equal(a[:, 0].cpu(), torch.from_numpy(pos)) and not torch.equal(a, b)
    an = a.cpu().numpy()
    for r in range(0, N, 11):
        assert not set(an[r, 1:].tolist()) & set(u2i[int(query[r])])
    assert int(smp.gave_up) == 0
    plain = EpochNegativeSampler(n_items, query, pos, 6, seed=5).sample()
    assert plain.shape == (N, 7) and int(plain[:, 1:].max()) < n_items
