"""Shared test scaffolding: a seeded synthetic multi-slot problem in BOTH layouts --
the reference's (per-feature weights + {feature: column} dict, consumed by oracle/) and the
product's (fused tables + [B,F] int32 global rows, consumed by the C ABI)."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import recbox_oracle as oracle  # noqa: E402  (tests are allowed to import the oracle)


class Problem:
    """kinds: string over {'c','n'} giving the feature order, e.g. 'nncccn'."""

    def __init__(self, B, kinds, D, vocab=50, seed=0, pad=True, scale=0.5, zipf=None, pad_frac=0.1):
        g = torch.Generator().manual_seed(seed)
        rng = np.random.default_rng(seed)
        self.B, self.D, self.kinds = B, D, kinds
        self.features = OrderedDict()
        self.W, self.W1 = OrderedDict(), OrderedDict()      # per-feature weights (reference layout)
        self.X = OrderedDict()                              # {feature: float64 [B]} like get_inputs
        self.cat_names, self.num_names, self.cat_pos, self.num_pos = [], [], [], []
        self.field_off, self.pad_row = [], []
        off = 0
        for pos, k in enumerate(kinds):
            name = "%s%d" % ("C" if k == "c" else "I", pos)
            if k == "c":
                V = int(vocab[len(self.cat_names)]) if not isinstance(vocab, int) else vocab
                spec = {"type": "categorical", "source": "", "vocab_size": V}
                if pad:
                    spec["padding_idx"] = 0
                w = torch.randn(V, D, generator=g) * scale
                w1 = torch.randn(V, 1, generator=g) * scale
                if pad:
                    w[0] = 0
                    w1[0] = 0
                lo = 0 if pad else 0
                if zipf:
                    ids = np.minimum(rng.zipf(zipf, size=B), V - 1)
                else:
                    ids = rng.integers(lo, V, size=B)
                if pad and pad_frac:
                    ids[rng.random(B) < pad_frac] = 0
                self.X[name] = torch.from_numpy(ids.astype(np.float64))
                self.cat_names.append(name)
                self.cat_pos.append(pos)
                self.field_off.append(off)
                self.pad_row.append(off if pad else -1)
                off += V
            else:
                spec = {"type": "numeric", "source": ""}
                w = torch.randn(D, 1, generator=g) * scale     # nn.Linear(1, D).weight
                w1 = torch.randn(1, 1, generator=g) * scale
                self.X[name] = torch.rand(B, generator=g, dtype=torch.float64)
                self.num_names.append(name)
                self.num_pos.append(pos)
            self.features[name] = spec
            self.W[name] = w
            self.W1[name] = w1
        self.R = off
        self.bias = torch.randn(1, generator=g) * scale
        self.F, self.Fn = len(self.cat_names), len(self.num_names)

    # ---- product layout -------------------------------------------------------------------
    def fused(self, device="cpu"):
        D = self.D
        table = torch.cat([self.W[n] for n in self.cat_names], 0) if self.F else torch.zeros(0, D)
        table_lr = torch.cat([self.W1[n].reshape(-1) for n in self.cat_names], 0) if self.F else torch.zeros(0)
        rows = (torch.stack([self.X[n].long() + o for n, o in zip(self.cat_names, self.field_off)], 1).int()
                if self.F else None)
        dense_x = torch.stack([self.X[n].float() for n in self.num_names], 1) if self.Fn else None
        dense_w = torch.stack([self.W[n].reshape(-1) for n in self.num_names], 0) if self.Fn else None
        dense_w_lr = torch.cat([self.W1[n].reshape(-1) for n in self.num_names], 0) if self.Fn else None
        out = dict(table=table, table_lr=table_lr, rows=rows, dense_x=dense_x, dense_w=dense_w,
                   dense_w_lr=dense_w_lr, bias=self.bias.clone())
        return {k: (v.to(device).contiguous() if v is not None else None) for k, v in out.items()}

    def batch_matrix(self, label=True):
        cols = [self.X[n].reshape(-1, 1) for n in self.features]
        if label:
            g = torch.Generator().manual_seed(1234)
            cols.append((torch.rand(self.B, 1, generator=g) < 0.5).double())
        return torch.cat(cols, 1)

    # ---- reference layout, through the oracle ---------------------------------------------
    def oracle_forward(self, leaf=False):
        W = OrderedDict((k, v.clone().requires_grad_(leaf)) for k, v in self.W.items())
        W1 = OrderedDict((k, v.clone().requires_grad_(leaf)) for k, v in self.W1.items())
        bias = self.bias.clone().requires_grad_(leaf)
        E = oracle.dict2tensor(oracle.embed_dict(self.X, self.features, W))
        fm = oracle.inner_product_interaction(E, "product_sum")
        lr = oracle.logistic_regression(self.X, self.features, W1, bias)
        return E, fm, lr, W, W1, bias

    def oracle_grads(self, dE, d_fm, d_lr):
        """dense per-feature grads of sum(E*dE) + sum(fm*d_fm) + sum(lr*d_lr), assembled into the
        fused layout (padding rows are zero because nn.Embedding masks them)."""
        E, fm, lr, W, W1, bias = self.oracle_forward(leaf=True)
        loss = (E * dE).sum() + (fm.reshape(-1) * d_fm).sum() + (lr.reshape(-1) * d_lr).sum()
        loss.backward()
        D = self.D
        z = lambda *s: torch.zeros(*s)
        g_table = torch.cat([W[n].grad for n in self.cat_names], 0) if self.F else z(0, D)
        g_table_lr = torch.cat([W1[n].grad.reshape(-1) for n in self.cat_names], 0) if self.F else z(0)
        g_dense_w = torch.stack([W[n].grad.reshape(-1) for n in self.num_names], 0) if self.Fn else None
        g_dense_w_lr = torch.cat([W1[n].grad.reshape(-1) for n in self.num_names], 0) if self.Fn else None
        return g_table, g_table_lr, g_dense_w, g_dense_w_lr, bias.grad


def assert_close(a, b, rtol=1e-5, atol_scale=1e-5, what=""):
    """|a-b| <= rtol*|b| + atol_scale*max|b| : fp32 contract of BASELINE.json (1e-5 relative), with
    the absolute floor tied to the tensor's own magnitude (sums of mixed-sign terms cancel)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, tuple(a.shape), tuple(b.shape))
    if b.numel() == 0:
        return
    atol = atol_scale * float(b.abs().max())
    err = (a - b).abs()
    tol = rtol * b.abs() + atol
    bad = err > tol
    assert not bool(bad.any()), "%s: %d / %d outside tolerance, max err %.3e (max |ref| %.3e)" % (
        what, int(bad.sum()), b.numel(), float(err.max()), float(b.abs().max()))


class CpuKern:
    """Stand-in kernel provider for host-logic tests of recbox_b200.sharded on CPU (gloo): the same
    call surface as recbox_b200.ops, implemented with the oracle / torch CPU ops.  TEST ONLY."""

    @staticmethod
    def shard_route(rows, world):
        send, counts, pos = oracle.shard_route(rows.numpy(), world)
        return (torch.from_numpy(send.astype(np.int32)), torch.from_numpy(pos.astype(np.int32)),
                torch.from_numpy(counts.astype(np.int32)))

    @staticmethod
    def gather_rows(table, ids):
        return torch.from_numpy(oracle.gather_rows_numpy(table.numpy(), ids.numpy()))

    @staticmethod
    def scatter_add_rows(g, ids, pad_row, g_table):
        keep = torch.ones_like(ids, dtype=torch.bool) if pad_row is None else ids != pad_row
        g_table.index_add_(0, ids[keep].long(), g[keep])

    @staticmethod
    def embed_fm_fwd(table, table_lr, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias,
                     want_E=True, want_lr=True, n_slots=None, **_):
        B, F = rows.shape
        Fn = len(num_pos)
        D = table.shape[1]
        E = torch.zeros(B, n_slots or (F + Fn), D)
        E[:, list(cat_pos)] = table[rows.long()]
        if Fn:
            E[:, list(num_pos)] = dense_x[:, :, None] * dense_w[None]
        S = E.sum(1)
        fm = oracle.inner_product_interaction(E, "product_sum").reshape(-1)
        lr = None
        if want_lr:
            lr = table_lr[rows.long()].sum(1)
            if Fn:
                lr = lr + (dense_x * dense_w_lr[None]).sum(1)
            if lr_bias is not None:
                lr = lr + lr_bias
        return (E if want_E else None), S, fm, lr

    @staticmethod
    def embed_fm_bwd(table, rows, cat_pos, pad_row, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                     g_table, g_table_lr, g_dense_w, g_dense_w_lr, g_lr_bias, D, R, n_slots=None, **_):
        if rows is not None and len(cat_pos):
            B, F = rows.shape
            e_cat = table[rows.long()]                                          # [B,F,D]
            g_e = dE[:, list(cat_pos)] + d_fm.view(-1, 1, 1) * (S[:, None, :] - e_cat)
            keep = torch.ones(B, F, dtype=torch.bool)
            if pad_row is not None:
                keep = rows != torch.tensor(pad_row, dtype=rows.dtype)[None]
            g_table.index_add_(0, rows[keep].long(), g_e[keep])
            if g_table_lr is not None and d_lr is not None:
                g_table_lr.index_add_(0, rows[keep].long(), d_lr.view(-1, 1).expand(B, F)[keep])
        if len(num_pos) and g_dense_w is not None:
            e_num = dense_x[:, :, None] * dense_w[None]
            g_n = dE[:, list(num_pos)] + d_fm.view(-1, 1, 1) * (S[:, None, :] - e_num)
            g_dense_w += (dense_x[:, :, None] * g_n).sum(0)
            if g_dense_w_lr is not None and d_lr is not None:
                g_dense_w_lr += (dense_x * d_lr.view(-1, 1)).sum(0)
        if g_lr_bias is not None and d_lr is not None:
            g_lr_bias += d_lr.sum()


    # ---- streamed exchange (csrc/shard_stream.cu), same buffer layouts and slot bookkeeping as the kernels; "peer
    # pointers" are the ranks' shared-memory tensors (recbox_b200.sharded.FileBlock) ------------------------------
    XS_T = 4                    # small tiles: several tiles and a partial last tile in every test batch

    @staticmethod
    def xs_tile_samples(F, D):
        return CpuKern.XS_T

    @staticmethod
    def xs_route(rows, R, D, rank, world, cap, cursor, tile_base, tile_cnt, pair_sorted, overflow, inbox_peers):
        B, F = rows.shape
        T = CpuKern.XS_T
        r = rows.numpy()
        inb = [t.view(torch.int32) for t in inbox_peers]
        for tile in range((B + T - 1) // T):
            blk = r[tile * T:(tile + 1) * T].reshape(-1)
            n = 0
            for o in range(world):
                sel = [k for k, v in enumerate(blk) if 0 <= v < R and v % world == o]
                base = int(cursor[o])
                cursor[o] += len(sel)
                tile_base[tile, o], tile_cnt[tile, o] = base, len(sel)
                if base + len(sel) > cap:
                    overflow[0] = 1
                for j, k in enumerate(sel):
                    if base + j < cap:
                        inb[o][rank * cap + base + j] = int(blk[k]) // world
                    pair_sorted[tile * T * F + n] = k
                    n += 1

    @staticmethod
    def xs_barrier(flags_peers, meta_peers, cursor, rank, world, epoch, device):
        import time
        if cursor is not None:
            for o in range(world):
                meta_peers[o].view(torch.int32)[rank] = int(cursor[o])
                cursor[o] = 0
        for o in range(world):
            flags_peers[o].view(torch.int32)[rank] = epoch
        mine = flags_peers[rank].view(torch.int32)
        t0 = time.time()
        while any(int(mine[q]) < epoch for q in range(world)):
            time.sleep(0.0005)
            assert time.time() - t0 < 60, "stand-in barrier timed out"

    @staticmethod
    def _lr_view(phys, D, lr_vec, lr_in_row):
        return lr_vec if lr_vec is not None else (phys[:, D] if lr_in_row else None)

    @staticmethod
    def xs_serve(phys, D, lr_vec, lr_in_row, inbox_ids, meta, cap, rank, world, rowbuf_peers, rowbuf_lr_peers):
        lr = CpuKern._lr_view(phys, D, lr_vec, lr_in_row)
        for q in range(world):
            n = min(int(meta[q]), cap)
            ids = inbox_ids[q * cap:q * cap + n].long()
            rowbuf_peers[q][rank * cap * D:(rank * cap + n) * D] = phys[ids, :D].reshape(-1)
            if lr is not None and rowbuf_lr_peers is not None:
                rowbuf_lr_peers[q][rank * cap:rank * cap + n] = lr[ids]

    @staticmethod
    def _pairs(tile_base, tile_cnt, pair_sorted, B, F, world):
        """-> list of (b, f, owner, slot) in the kernels' tile / run order."""
        T = CpuKern.XS_T
        out = []
        for tile in range((B + T - 1) // T):
            n = 0
            for o in range(world):
                for j in range(int(tile_cnt[tile, o])):
                    k = int(pair_sorted[tile * T * F + n])
                    out.append((tile * T + k // F, k % F, o, int(tile_base[tile, o]) + j))
                    n += 1
        return out

    @staticmethod
    def xs_consume(rowbuf, rowbuf_lr, tile_base, tile_cnt, pair_sorted, cat_pos, dense_x, dense_w, dense_w_lr, num_pos,
                   lr_bias, B, cap, D, world, want_E=True, want_lr=True, num_widx=None, n_slots=None, out=None):
        F, Fn = len(cat_pos), len(num_pos)
        E = torch.zeros(B, n_slots or (F + Fn), D)
        lr = torch.zeros(B)
        rb = rowbuf.view(-1, D)
        for b, f, o, slot in CpuKern._pairs(tile_base, tile_cnt, pair_sorted, B, F, world):
            if slot < cap:
                E[b, cat_pos[f]] = rb[o * cap + slot]
                if want_lr:
                    lr[b] += rowbuf_lr[o * cap + slot]
        if Fn:
            E[:, list(num_pos)] = dense_x[:, :, None] * dense_w[None]
            if want_lr:
                lr = lr + (dense_x * dense_w_lr[None]).sum(1)
        if want_lr and lr_bias is not None:
            lr = lr + lr_bias
        S = E.sum(1)
        fm = oracle.inner_product_interaction(E, "product_sum").reshape(-1)
        if out is not None:
            for dst, src in zip(out, (E, S, fm, lr)):
                if dst is not None:
                    dst.copy_(src)
            return out
        return (E if want_E else None), S, fm, (lr if want_lr else None)

    @staticmethod
    def xs_grad_push(E, rowbuf, S, dE, d_fm, d_lr, rows, pad_row, tile_base, tile_cnt, pair_sorted, cat_pos, cap, D,
                     n_slots, rank, world, ginbox_peers, ginbox_lr_peers, dense_x=None, dense_w=None, num_pos=(),
                     num_widx=None, g_dense_w=None, g_dense_w_lr=None, g_lr_bias=None):
        B, F = rows.shape
        CpuKern.embed_fm_bwd(None, None, [], None, dense_x, dense_w, num_pos, None, S, dE, d_fm, d_lr, None, None,
                             g_dense_w, g_dense_w_lr, g_lr_bias, D, 0)
        rb = rowbuf.view(-1, D)
        for b, f, o, slot in CpuKern._pairs(tile_base, tile_cnt, pair_sorted, B, F, world):
            if slot >= cap:
                continue
            g = torch.zeros(D)
            is_pad = pad_row is not None and int(rows[b, f]) == pad_row[f]
            if not is_pad:
                if dE is not None:
                    g = dE[b, cat_pos[f]].clone()
                if d_fm is not None:
                    e = E[b, cat_pos[f]] if E is not None else rb[o * cap + slot]
                    g = g + d_fm[b] * (S[b] - e)
            ginbox_peers[o][(rank * cap + slot) * D:(rank * cap + slot + 1) * D] = g
            if d_lr is not None and ginbox_lr_peers is not None:
                ginbox_lr_peers[o][rank * cap + slot] = 0.0 if is_pad else float(d_lr[b])

    @staticmethod
    def xs_apply(ginbox, ginbox_lr, inbox_ids, meta, cap, world, g_phys, D, g_lr_vec, lr_in_row):
        g_lr = CpuKern._lr_view(g_phys, D, g_lr_vec, lr_in_row)
        gi = ginbox.view(-1, D)
        for q in range(world):
            n = min(int(meta[q]), cap)
            ids = inbox_ids[q * cap:q * cap + n].long()
            g_phys[:, :D].index_add_(0, ids, gi[q * cap:q * cap + n])
            if g_lr is not None and ginbox_lr is not None:
                g_lr.index_add_(0, ids, ginbox_lr[q * cap:q * cap + n])
