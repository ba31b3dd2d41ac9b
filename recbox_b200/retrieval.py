"""Retrieval evaluation on the GPU: the host-side mirror of
`recbox/core/metrics.py:11-68` (`evaluate_metrics`, `evaluate_block`) and `recbox/utils/ann/faiss.py:3-14`
(`FaissIndex`), which `MatchingModel.evaluate` (`matching/pytorch/models/match_model.py:205-225`) runs every epoch.

The reference adds every item vector to a CPU faiss.IndexFlatIP, searches top-500 per 1000-user chunk, builds a dense
[chunk, num_items] float64 mask in numpy to push the user's train items down, argsorts, and calls Python metric objects
per user.  Here the corpus stays in HBM, `rbx_topk_ip` streams it once per user chunk with the top-k selection fused into
the GEMM epilogue, and `rbx_rank_metrics` does the mask / re-rank / metrics from CSR lists.  Same signatures, same metric
strings ("Recall(k=20)", ...), same averaging (np.average over users)."""
import re

import numpy as np
import torch

from . import ops
from ._lib import RbxError

_METRIC_RE = re.compile(r"^\s*(\w+)\(k=(\d+)\)\s*$")


def parse_metrics(metrics):
    """["Recall(k=20)", "NDCG(k=50)"] -> ([kind codes], [k]); unknown names raise NotImplementedError like
    core/metrics.py:24-28."""
    kinds, ks = [], []
    for m in metrics:
        mt = _METRIC_RE.match(m)
        if mt is None or mt.group(1) not in ops.METRIC_KINDS:
            raise NotImplementedError("metrics={} not implemented.".format(m))
        kinds.append(ops.METRIC_KINDS[mt.group(1)])
        ks.append(int(mt.group(2)))
    return kinds, ks


def _as_device_f32(x, device):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    return t.to(device=device, dtype=torch.float32).contiguous()


def build_csr(user2items, query_indices, device, sort=True):
    """{query_index: [items]} + the query order -> (ptr int64 [U+1], items int64, sorted ascending per row) on device."""
    lens = np.fromiter((len(user2items.get(q, ())) if hasattr(user2items, "get") else len(user2items[q])
                        for q in query_indices), dtype=np.int64, count=len(query_indices))
    ptr = np.zeros(len(query_indices) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    # at least one element: a chunk whose users have no items at all (cold start, empty train_user2items) must still hand
    # rbx_rank_metrics a non-null items pointer (ptr[U] == 0 tells it there is nothing to read)
    items = np.zeros(max(int(ptr[-1]), 1), dtype=np.int64)
    for i, q in enumerate(query_indices):
        row = user2items.get(q, ()) if hasattr(user2items, "get") else user2items[q]
        if len(row):
            a = np.asarray(row, dtype=np.int64)
            items[ptr[i]:ptr[i + 1]] = np.sort(a) if sort else a
    return torch.from_numpy(ptr).to(device), torch.from_numpy(items).to(device)


class FlatIPIndex(object):
    """Drop-in for recbox.utils.ann.FaissIndex (faiss.IndexFlatIP): exact inner-product search, corpus in HBM."""

    def __init__(self, corpus_vecs, dim=None, l2_normalize=False, index_name="IndexFlatIP", device="cuda"):
        if index_name != "IndexFlatIP":
            raise NotImplementedError("index_name={} not implemented.".format(index_name))
        self.l2_normalize = l2_normalize
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RbxError("FlatIPIndex needs a CUDA device (recbox_b200 has no CPU path)")
        v = _as_device_f32(corpus_vecs, self.device)
        if dim is not None and v.shape[-1] != dim:
            raise RbxError("corpus dim %d != %d" % (v.shape[-1], dim))
        if l2_normalize:
            v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)         # faiss.normalize_L2
        self.vecs = self._pad(v)
        self.dim = v.shape[-1]
        self.ntotal = v.shape[0]

    @staticmethod
    def _pad(v):
        """The kernel wants D % 4 == 0: zero columns change no inner product."""
        D = v.shape[1]
        if D % 4 == 0:
            return v
        out = torch.zeros((v.shape[0], (D + 3) // 4 * 4), dtype=v.dtype, device=v.device)
        out[:, :D] = v
        return out

    def search_device(self, query_vecs, topk=50):
        q = _as_device_f32(query_vecs, self.device)
        if self.l2_normalize:
            q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)
        return ops.topk_ip(self._pad(q), self.vecs, int(topk))

    def search(self, query_vecs, topk=50):
        scores, idx = self.search_device(query_vecs, topk)
        return scores.cpu().numpy(), idx.cpu().numpy()


def evaluate_metrics(user_embs, item_embs, train_user2items, valid_user2items, query_indices, metrics, num_workers=1,
                     device="cuda", chunk_users=1024, search_topk=500):
    """core/metrics.py:11-50.  Returns {metric string: mean over the query users}."""
    kinds, ks = parse_metrics(metrics)
    max_topk = max(ks) if ks else 0
    index = FlatIPIndex(item_embs, device=device)
    U = len(user_embs)
    T = min(int(search_topk), 1024)
    per_user = []
    query_indices = list(query_indices)
    for lo in range(0, U, chunk_users):
        hi = min(lo + chunk_users, U)
        _, cand = index.search_device(user_embs[lo:hi], topk=T)
        qi = query_indices[lo:hi]
        tp, ti = build_csr(train_user2items, qi, index.device)
        vp, vi = build_csr(valid_user2items, qi, index.device)
        _, _, out = ops.rank_metrics(cand, tp, ti, vp, vi, kinds, ks, kmax=max_topk)
        per_user.append(out)
    results = torch.cat(per_user, 0).cpu().numpy() if per_user else np.zeros((0, len(metrics)))
    average_result = np.average(results, axis=0).tolist()
    return dict(zip(metrics, average_result))
