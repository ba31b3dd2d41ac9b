"""Pin the retrieval / sampling restatements of oracle/recbox_oracle.py (f2, f3) against tests/golden/retrieval.npz,
minted by running the reference's own evaluate_metrics / evaluate_block / metric classes (recbox/core/metrics.py) with
a numpy stand-in for faiss.IndexFlatIP (oracle/make_golden.py golden_retrieval).  CPU only."""
import numpy as np
import pytest

from helpers import oracle
from test_oracle_golden import GOLD

import os


def load_retrieval():
    z = np.load(os.path.join(GOLD, "retrieval.npz"))
    g = {k: z[k] for k in z.files}
    query = [int(q) for q in g["query"]]
    train = {q: g["train_items"][g["train_ptr"][i]:g["train_ptr"][i + 1]].tolist() for i, q in enumerate(query)}
    valid = {q: g["valid_items"][g["valid_ptr"][i]:g["valid_ptr"][i + 1]].tolist() for i, q in enumerate(query)}
    return g, query, train, valid, [str(m) for m in g["metrics"]]


def test_oracle_evaluate_metrics_matches_reference():
    g, query, train, valid, metrics = load_retrieval()
    got = oracle.evaluate_metrics(g["user"].astype(np.float64), g["item"].astype(np.float64), train, valid, query, metrics)
    np.testing.assert_allclose([got[m] for m in metrics], g["average"], rtol=1e-12, atol=0)
    ns = {n: getattr(oracle, n) for n in ("Recall", "nRecall", "Precision", "F1", "DCG", "NDCG", "MRR", "HitRate", "MAP")}
    funcs = [eval(m, {}, ns) for m in metrics]
    _, per_user = oracle.evaluate_block(g["user"], g["item"], query, train, valid, funcs, 50)
    np.testing.assert_allclose(np.array(per_user, dtype=np.float64), g["per_user"], rtol=1e-12, atol=0)


def test_flat_ip_search_contract():
    rng = np.random.default_rng(0)
    q, c = rng.standard_normal((5, 8)).astype(np.float32), rng.standard_normal((40, 8)).astype(np.float32)
    c[7] = c[3]                                            # an exact tie: the smaller index comes first
    s, i = oracle.flat_ip_search(q, c, 10)
    assert s.shape == (5, 10) and i.dtype == np.int64
    assert np.all(np.diff(s, axis=1) <= 0)
    full = q @ c.T
    for u in range(5):
        assert set(i[u]) == set(np.argsort(-full[u], kind="stable")[:10])
        pos3, pos7 = np.where(i[u] == 3)[0], np.where(i[u] == 7)[0]
        if len(pos3) and len(pos7):
            assert pos3[0] < pos7[0]
    s, i = oracle.flat_ip_search(q, c[:4], 6)              # corpus smaller than k
    assert np.all(i[:, 4:] == -1) and np.all(np.isneginf(s[:, 4:]))


def test_unknown_metric_raises_like_reference():
    g, query, train, valid, _ = load_retrieval()
    with pytest.raises(NotImplementedError):
        oracle.evaluate_metrics(g["user"], g["item"], train, valid, query, ["Bogus(k=3)"])
    from recbox_b200.retrieval import parse_metrics
    with pytest.raises(NotImplementedError):
        parse_metrics(["Bogus(k=3)"])
    assert parse_metrics(["Recall(k=20)", "NDCG(k=5)"]) == ([0, 5], [20, 5])


def test_sampling_block_restatement():
    u2i = {0: [1, 2, 3], 1: list(range(0, 50, 2))}
    a = oracle.sampling_block(50, [0, 1, 0, 1], 200, u2i, ignore_pos_items=True, seed=4)
    assert a.shape == (4, 200) and a.min() >= 0 and a.max() < 50
    for row, q in zip(a, [0, 1, 0, 1]):
        assert not set(row.tolist()) & set(u2i[q])
    b = oracle.sampling_block(50, [0, 1], 1000, u2i, seed=4)
    assert b.shape == (2, 1000) and len(np.unique(b)) == 50


def test_retrieval_host_helpers():
    """build_csr keeps duplicates and sorts each row; the index refuses a CPU device (no CPU path)."""
    import torch
    from recbox_b200 import RbxError, retrieval
    u2i = {7: [5, 2, 2, 9], 3: [], 11: [4]}
    ptr, items = retrieval.build_csr(u2i, [11, 7, 3, 99], "cpu")
    assert ptr.tolist() == [0, 1, 5, 5, 5] and items.tolist() == [4, 2, 2, 5, 9]
    assert ptr.dtype == torch.int64 and items.dtype == torch.int64
    with pytest.raises(RbxError):
        retrieval.FlatIPIndex(np.zeros((4, 8), np.float32), dim=8, device="cpu")
    with pytest.raises(NotImplementedError):
        retrieval.FlatIPIndex(np.zeros((4, 8), np.float32), dim=8, index_name="IndexIVFFlat")
