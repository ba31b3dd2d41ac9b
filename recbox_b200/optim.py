"""Clip + optimizer step on the fused embedding tables (a12 exact-dense, f1 touched rows).

Reference: RankingModel.train_step (recbox/ranking/pytorch/models/ranking_model.py:191-197) and
MatchingModel.train_on_epoch (recbox/matching/pytorch/models/match_model.py:193-199) run
`clip_grad_norm_(self.parameters(), max_norm)` then `optimizer.step()` densely over every table row,
every step -- O(table) work for O(batch) information.  Two replacements, both device-resident (the clip
coefficient never visits the host):

  DenseTableOptimizer        exactly the reference's arithmetic (torch.optim.Adam, single tensor), one pass.
  TouchedRowsOptimizer       lists the rows the batch touched (bitmap unique, csrc/dedup.cu) and updates only
                             those; the consumed gradient rows are cleared in the same pass, so the dense
                             gradient table needs no per-step memset.  SGD and Adagrad are exact; "adam_rows" /
                             "sparse_adam" are lazy Adam (torch.optim.SparseAdam semantics), NOT the reference's
                             dense Adam -- use them when that trade is wanted (100 M-row tables).

Also `collate_unique`: the GPU form of the reference's collate_fn_unique (a14).
"""
import torch

from . import ops
from ._lib import RbxError


class _ClipState(object):
    def __init__(self, device):
        self.acc = torch.zeros(1, dtype=torch.float64, device=device)
        self.coef = torch.ones(1, dtype=torch.float32, device=device)
        self.norm = torch.zeros(1, dtype=torch.float32, device=device)


class DenseTableOptimizer(object):
    """clip_grad_norm_ + torch.optim.Adam.step over whole (fused) tables, in the reference's operation order."""

    def __init__(self, params_and_grads, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.pairs = [(w, g, torch.zeros_like(w), torch.zeros_like(w)) for w, g in params_and_grads]
        self.lr, self.betas, self.eps, self.t = lr, betas, eps, 0
        self.clip = _ClipState(self.pairs[0][0].device)

    def step(self, max_norm=None, extra_sqnorm=None):
        """extra_sqnorm: float64 [1] device tensor holding sum |g|^2 of the parameters outside this optimizer
        (the dense tail), so the clip coefficient is the GLOBAL one of clip_grad_norm_(all params)."""
        self.t += 1
        clip = None
        if max_norm is not None:
            c = self.clip
            c.acc.zero_()
            if extra_sqnorm is not None:
                c.acc.add_(extra_sqnorm)
            for _, g, _, _ in self.pairs:
                ops.sqnorm_(g, c.acc)
            ops.clip_coef(c.acc, max_norm, c.coef, c.norm)
            clip = c.coef
        for w, g, m, v in self.pairs:
            ops.adam_dense_(w, g, m, v, self.t, self.lr, self.betas[0], self.betas[1], self.eps, clip=clip)
        return clip


class TouchedRowsOptimizer(object):
    """Clip + update of the rows a batch touched.  tables: list of (w, g) with w [R, D] or [R] sharing ONE row
    numbering (e.g. the fused embedding table and its D = 1 first-order twin when lr_delta is all zero)."""

    def __init__(self, tables, kind="adam_rows", lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if kind not in ops.OPTIM_KINDS:
            raise RbxError("TouchedRowsOptimizer: kind must be one of %s" % sorted(ops.OPTIM_KINDS))
        self.kind, self.lr, self.betas, self.eps, self.t = kind, lr, betas, eps, 0
        k = ops.OPTIM_KINDS[kind]
        self.tables = []
        for w, g in tables:
            m = torch.zeros_like(w) if k >= 2 else None
            v = torch.zeros_like(w) if k >= 1 else None
            self.tables.append((w, g, m, v))
        self.R = int(self.tables[0][0].shape[0])
        self.clip = _ClipState(self.tables[0][0].device)

    def touched(self, rows):
        """rows: int32 CUDA tensor of global row ids (any shape) -> (sorted unique rows buffer, device count)."""
        uniq, _, _, n_out = ops.unique_ids(rows, self.R, want_first=False, want_inverse=False, sync=False)
        self.last_counts = n_out              # [n unique, n ids outside [0, R)] on the device; see out_of_range()
        return uniq, n_out[:1]

    def out_of_range(self):
        """Ids of the last step that fell outside [0, R) and were therefore not updated (synchronises)."""
        c = getattr(self, "last_counts", None)
        return int(c[1].item()) if c is not None and c.numel() > 1 else 0

    def step(self, rows, max_norm=None, extra_sqnorm=None, zero_grad=True):
        self.t += 1
        uniq, n_rows = self.touched(rows)
        clip = None
        if max_norm is not None:
            c = self.clip
            c.acc.zero_()
            if extra_sqnorm is not None:
                c.acc.add_(extra_sqnorm)
            for _, g, _, _ in self.tables:
                ops.sqnorm_rows_(g, uniq, n_rows, c.acc)
            ops.clip_coef(c.acc, max_norm, c.coef, c.norm)
            clip = c.coef
        for w, g, m, v in self.tables:
            ops.optim_rows_(w, g, m, v, uniq, n_rows, self.t, self.kind, self.lr, self.betas[0], self.betas[1], self.eps,
                            clip=clip, zero_grad=zero_grad)
        return clip


def collate_unique(item_indexes, vocab_size, reference_inverse=True):
    """GPU form of collate_fn_unique (recbox/matching/pytorch/dataloaders/h5_generator.py:45-58) for the item-id
    block [B, 1 + negs] of a batch: returns (unique, unique_indexes, inverse_indexes) with
      unique          sorted distinct item ids                          (h5_generator.py:49)
      unique_indexes  first flat position of each (what its flip + scatter_ computes, :50-52) -- index the
                      flattened item features with it to keep one row per distinct item (:54-55)
      inverse_indexes the map back to [B * (1 + negs)].  The reference returns it FLIPPED (it flips in place at
                      :51 and returns that tensor at :58, under a "TODO: check correctness"); reference_inverse=True
                      reproduces that bit for bit, False gives the un-flipped map of torch.unique."""
    uniq, first, inverse = ops.unique_ids(item_indexes, vocab_size)
    if reference_inverse:
        inverse = inverse.flip([0])
    return uniq, first, inverse
