"""Thin tensor-level wrappers over the C ABI (include/recbox_b200.h).

torch is plumbing here: it owns device memory and the current stream; every op below is ONE call
into librecbox_b200.so with raw pointers.  There is no fallback of any kind: a non-CUDA tensor or a
missing library raises.
"""
import ctypes
import threading

import torch

from . import _lib
from ._lib import RbxError

MODES = {"product_sum": 0, "bi_interaction": 1, "inner_product": 2, "elementwise_product": 3}
_DTYPE_CODE = {torch.float64: 0, torch.float32: 1, torch.int64: 2, torch.int32: 3}


# Device of the call being assembled.  Arguments are evaluated left to right, so every `_p()` of a call
# runs before its trailing `_stream()`, which in turn runs before `_call()`: the first tensor pins the
# device, a tensor on another device is an error, the stream is that device's current stream, and the C
# call runs inside that device's context (the reference's `get_device(gpu)` hands out `cuda:<gpu>`
# without ever calling `set_device`, torch_utils.py:37-42, so tensors off the current device are normal).
_ctx = threading.local()


def _note_device(dev, name="tensor"):
    cur = getattr(_ctx, "dev", None)
    if cur is None:
        _ctx.dev = dev
    elif cur != dev:
        _ctx.dev = None
        raise RbxError("%s is on %s but an earlier argument of the same call is on %s" % (name, dev, cur))


def _p(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not t.is_cuda:
        _ctx.dev = None
        raise RbxError("%s must be a CUDA tensor (recbox_b200 has no CPU path)" % name)
    if not t.is_contiguous():
        _ctx.dev = None
        raise RbxError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        _ctx.dev = None
        raise RbxError("%s must be %s, got %s" % (name, dtype, t.dtype))
    _note_device(t.device, name)
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    dev = getattr(_ctx, "dev", None)
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _i32(xs):
    xs = list(xs)
    return (ctypes.c_int32 * max(len(xs), 1))(*xs)


def _call(name, *args):
    lib = _lib.load()
    dev, _ctx.dev = getattr(_ctx, "dev", None), None
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RbxError("%s failed (%d): %s" % (name, rc, lib.rbx_last_error().decode()))


F32, I32 = torch.float32, torch.int32


def l2_set_persisting_bytes(nbytes):
    """Reserve part of the current device's L2 for the kernels' evict_last lines; returns the size in effect."""
    got = _lib.load().rbx_l2_set_persisting_bytes(int(nbytes))
    if got < 0:
        raise RbxError("rbx_l2_set_persisting_bytes failed: %s" % _lib.load().rbx_last_error().decode())
    return int(got)


# ------------------------------------------------------------------------------------------- a1
def split_batch(batch, col_kind, col_slot, field_off, F, Fn, want_label=True, field_rows=None, n_bad=None):
    """[B, n_cols] float64 device matrix -> (rows int32 [B,F], dense_x fp32 [B,Fn], label fp32 [B]).
    field_rows: vocabulary size per categorical slot -- ids outside it become row -1 (zero row, no gradient) and are
    counted in the int32 device counter `n_bad` instead of aliasing another feature's rows."""
    if batch.dim() != 2 or batch.dtype != torch.float64 or not batch.is_cuda or batch.stride(1) != 1:
        raise RbxError("split_batch: batch must be a CUDA float64 matrix with unit column stride")
    B, n_cols = batch.shape
    dev = batch.device
    rows = torch.empty((B, F), dtype=I32, device=dev) if F else None
    dense = torch.empty((B, Fn), dtype=F32, device=dev) if Fn else None
    label = torch.empty((B,), dtype=F32, device=dev) if want_label and 3 in col_kind else None
    kinds = (ctypes.c_int8 * max(n_cols, 1))(*col_kind)
    slots = (ctypes.c_int16 * max(n_cols, 1))(*col_slot)
    offs = (ctypes.c_int64 * max(F, 1))(*field_off)
    nrows = (ctypes.c_int64 * max(F, 1))(*field_rows) if field_rows is not None else None
    _note_device(dev, "batch")
    _call("rbx_split_batch_f64", ctypes.c_void_p(batch.data_ptr()), B, n_cols, batch.stride(0), kinds, slots, offs, nrows,
          F, Fn, _p(rows), _p(dense), _p(label), _p(n_bad, I32, "n_bad"), _stream())
    return rows, dense, label


def pack_columns(cols, add=None, as_rows=True, vocab=None, n_bad=None):
    """list of [B] device tensors (any of f64/f32/i64/i32, any stride) -> [B, n] int32 (+add) or fp32.
    vocab / n_bad: per-column vocabulary guard, see split_batch."""
    n = len(cols)
    B = cols[0].shape[0]
    dev = cols[0].device
    for c in cols:
        if not c.is_cuda or c.dim() != 1 or c.shape[0] != B or c.dtype not in _DTYPE_CODE:
            raise RbxError("pack_columns: every column must be a 1-D CUDA tensor of f64/f32/i64/i32 and equal length")
    out = torch.empty((B, n), dtype=I32 if as_rows else F32, device=dev)
    ptrs = (ctypes.c_void_p * n)(*[c.data_ptr() for c in cols])
    strides = (ctypes.c_int64 * n)(*[c.stride(0) if B > 1 else 1 for c in cols])
    dts = (ctypes.c_int8 * n)(*[_DTYPE_CODE[c.dtype] for c in cols])
    adds = (ctypes.c_int64 * n)(*(add if add is not None else [0] * n))
    voc = (ctypes.c_int64 * n)(*vocab) if (vocab is not None and as_rows) else None
    _call("rbx_pack_columns", ptrs, strides, dts, adds, voc, n, B, 1 if as_rows else 0, _p(out), _p(n_bad, I32, "n_bad"),
          _stream())
    return out


def unpack_ids_u16(ids16, field_off):
    """[B, F] int16 tensor holding uint16 local ids (loader.PackedDataset compact form) -> int32 fused-table rows."""
    if ids16.dtype != torch.int16 or ids16.dim() != 2:
        raise RbxError("unpack_ids_u16: ids must be a [B, F] int16 tensor (uint16 bit patterns)")
    B, F = ids16.shape
    rows = torch.empty((B, F), dtype=I32, device=ids16.device)
    offs = (ctypes.c_int64 * max(F, 1))(*field_off) if field_off is not None else None
    _call("rbx_unpack_ids_u16", _p(ids16, torch.int16, "ids"), B, F, offs, _p(rows), _stream())
    return rows


def zero_(t):
    """Streaming zero-fill of a contiguous fp32 tensor on the current stream (the fused gradient buffer)."""
    _call("rbx_zero_f32", _p(t, F32, "t"), t.numel(), _stream())
    return t


# ---------------------------------------------------------------------------------- K1 / K2 / K3
def embed_fm_fwd(table, table_lr, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias,
                 want_E=True, want_S=True, want_fm=True, want_lr=True, B=None, lr_delta=None, num_widx=None, D=None,
                 n_slots=None, device=None):
    """Fused gather + FM + LR forward.  Returns (E, S, fm_out, lr_out); unrequested ones are None."""
    F = len(cat_pos)
    Fn = len(num_pos)
    ref = table if table is not None else (dense_w if dense_w is not None else (rows if rows is not None else dense_x))
    if D is None:
        D = ref.shape[-1] if (table is not None or dense_w is not None) else 1
    dev = device if device is not None else ref.device
    if B is None:
        B = rows.shape[0] if rows is not None else dense_x.shape[0]
    R = table.shape[0] if table is not None else (table_lr.numel() if table_lr is not None else 0)
    Ft = n_slots or (F + Fn)
    E = torch.empty((B, Ft, D), dtype=F32, device=dev) if want_E else None
    S = torch.empty((B, D), dtype=F32, device=dev) if want_S else None
    fm = torch.empty((B,), dtype=F32, device=dev) if want_fm else None
    lr = torch.empty((B,), dtype=F32, device=dev) if want_lr else None
    _call("rbx_embed_fm_fwd", _p(table, F32, "table"), _p(table_lr, F32, "table_lr"), _p(rows, I32, "rows"),
          _i32(cat_pos), _i32(lr_delta) if lr_delta is not None else None, _p(dense_x, F32, "dense_x"),
          _p(dense_w, F32, "dense_w"), _p(dense_w_lr, F32, "dense_w_lr"), _i32(num_pos),
          _i32(num_widx) if num_widx is not None else None, _p(lr_bias, F32, "lr_bias"),
          _p(E), _p(S), _p(fm), _p(lr), B, R, F, Fn, D, Ft, _stream())
    return E, S, fm, lr


def embed_fm_bwd(table, rows, cat_pos, pad_row, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                 g_table, g_table_lr, g_dense_w, g_dense_w_lr, g_lr_bias, D, R, B=None, lr_delta=None, num_widx=None,
                 n_slots=None):
    """Gradient scatter-add (accumulates into the g_* tensors, which the caller zero-fills)."""
    F = len(cat_pos)
    Fn = len(num_pos)
    if B is None:
        B = rows.shape[0] if rows is not None else dense_x.shape[0]
    _call("rbx_embed_fm_bwd", _p(table, F32, "table"), _p(rows, I32, "rows"), _i32(cat_pos),
          _i32(pad_row if pad_row is not None else [-1] * F), _i32(lr_delta) if lr_delta is not None else None,
          _p(dense_x, F32, "dense_x"), _p(dense_w, F32, "dense_w"), _i32(num_pos),
          _i32(num_widx) if num_widx is not None else None, _p(E, F32, "E"), _p(S, F32, "S"), _p(dE, F32, "dE"),
          _p(d_fm, F32, "d_fm"), _p(d_lr, F32, "d_lr"), _p(g_table, F32, "g_table"),
          _p(g_table_lr, F32, "g_table_lr"), _p(g_dense_w, F32, "g_dense_w"),
          _p(g_dense_w_lr, F32, "g_dense_w_lr"), _p(g_lr_bias, F32, "g_lr_bias"), B, R, F, Fn, D, n_slots or (F + Fn),
          _stream())


def embed_fm_fwd_sharded(shard_tables, shard_tables_lr, world, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos,
                         lr_bias, R, D, want_E=True, want_S=True, want_fm=True, want_lr=True, num_widx=None, n_slots=None):
    """Fused forward over a row-sharded table reached through NVLink peer mappings.  shard_tables /
    shard_tables_lr: ctypes arrays of `world` device pointers valid on the current device."""
    F, Fn = len(cat_pos), len(num_pos)
    B = rows.shape[0] if rows is not None else dense_x.shape[0]
    dev = (rows if rows is not None else dense_x).device
    Ft = n_slots or (F + Fn)
    E = torch.empty((B, Ft, D), dtype=F32, device=dev) if want_E else None
    S = torch.empty((B, D), dtype=F32, device=dev) if want_S else None
    fm = torch.empty((B,), dtype=F32, device=dev) if want_fm else None
    lr = torch.empty((B,), dtype=F32, device=dev) if want_lr else None
    _call("rbx_embed_fm_fwd_sharded", shard_tables, shard_tables_lr, world, _p(rows, I32, "rows"), _i32(cat_pos),
          _p(dense_x, F32, "dense_x"), _p(dense_w, F32, "dense_w"), _p(dense_w_lr, F32, "dense_w_lr"), _i32(num_pos),
          _i32(num_widx) if num_widx is not None else None, _p(lr_bias, F32, "lr_bias"), _p(E), _p(S), _p(fm), _p(lr),
          B, R, F, Fn, D, Ft, _stream())
    return E, S, fm, lr


def embed_fm_bwd_sharded(shard_tables, shard_g_tables, shard_g_tables_lr, world, rows, cat_pos, pad_row, dense_x,
                         dense_w, num_pos, E, S, dE, d_fm, d_lr, g_dense_w, g_dense_w_lr, g_lr_bias, R, D,
                         num_widx=None, n_slots=None):
    F, Fn = len(cat_pos), len(num_pos)
    B = rows.shape[0] if rows is not None else dense_x.shape[0]
    _call("rbx_embed_fm_bwd_sharded", shard_tables, shard_g_tables, shard_g_tables_lr, world, _p(rows, I32, "rows"),
          _i32(cat_pos), _i32(pad_row if pad_row is not None else [-1] * F), _p(dense_x, F32, "dense_x"),
          _p(dense_w, F32, "dense_w"), _i32(num_pos), _i32(num_widx) if num_widx is not None else None,
          _p(E, F32, "E"), _p(S, F32, "S"), _p(dE, F32, "dE"), _p(d_fm, F32, "d_fm"), _p(d_lr, F32, "d_lr"),
          _p(g_dense_w, F32, "g_dense_w"), _p(g_dense_w_lr, F32, "g_dense_w_lr"), _p(g_lr_bias, F32, "g_lr_bias"),
          B, R, F, Fn, D, n_slots or (F + Fn), _stream())


def embed_fm_fwd_sharded_rowlr(shard_tables, world, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, R, D,
                               want_E=True, want_S=True, want_fm=True, want_lr=True, num_widx=None, n_slots=None):
    """As embed_fm_fwd_sharded, for shards in the ROW+LR layout ([cap, 2 D] floats, first-order weight at column D)."""
    F, Fn = len(cat_pos), len(num_pos)
    B = rows.shape[0] if rows is not None else dense_x.shape[0]
    dev = (rows if rows is not None else dense_x).device
    Ft = n_slots or (F + Fn)
    E = torch.empty((B, Ft, D), dtype=F32, device=dev) if want_E else None
    S = torch.empty((B, D), dtype=F32, device=dev) if want_S else None
    fm = torch.empty((B,), dtype=F32, device=dev) if want_fm else None
    lr = torch.empty((B,), dtype=F32, device=dev) if want_lr else None
    _call("rbx_embed_fm_fwd_sharded_rowlr", shard_tables, world, _p(rows, I32, "rows"), _i32(cat_pos),
          _p(dense_x, F32, "dense_x"), _p(dense_w, F32, "dense_w"), _p(dense_w_lr, F32, "dense_w_lr"), _i32(num_pos),
          _i32(num_widx) if num_widx is not None else None, _p(lr_bias, F32, "lr_bias"), _p(E), _p(S), _p(fm), _p(lr),
          B, R, F, Fn, D, Ft, _stream())
    return E, S, fm, lr


def embed_fm_bwd_sharded_rowlr(shard_tables, shard_g_tables, world, rows, cat_pos, pad_row, dense_x, dense_w, num_pos,
                               E, S, dE, d_fm, d_lr, g_dense_w, g_dense_w_lr, g_lr_bias, R, D, num_widx=None, n_slots=None):
    F, Fn = len(cat_pos), len(num_pos)
    B = rows.shape[0] if rows is not None else dense_x.shape[0]
    _call("rbx_embed_fm_bwd_sharded_rowlr", shard_tables, shard_g_tables, world, _p(rows, I32, "rows"),
          _i32(cat_pos), _i32(pad_row if pad_row is not None else [-1] * F), _p(dense_x, F32, "dense_x"),
          _p(dense_w, F32, "dense_w"), _i32(num_pos), _i32(num_widx) if num_widx is not None else None,
          _p(E, F32, "E"), _p(S, F32, "S"), _p(dE, F32, "dE"), _p(d_fm, F32, "d_fm"), _p(d_lr, F32, "d_lr"),
          _p(g_dense_w, F32, "g_dense_w"), _p(g_dense_w_lr, F32, "g_dense_w_lr"), _p(g_lr_bias, F32, "g_lr_bias"),
          B, R, F, Fn, D, n_slots or (F + Fn), _stream())


# ------------------------------------------------------------------------------------- a5 / a9
def gather_rows(table, ids):
    N = ids.numel()
    D = table.shape[1]
    out = torch.empty(tuple(ids.shape) + (D,), dtype=F32, device=table.device)
    _call("rbx_gather_rows", _p(table, F32, "table"), _p(ids, I32, "ids"), _p(out), N, D, _stream())
    return out


def scatter_add_rows(g, ids, pad_row, g_table):
    N = ids.numel()
    D = g_table.shape[1]
    _call("rbx_scatter_add_rows", _p(g, F32, "g"), _p(ids, I32, "ids"), -1 if pad_row is None else int(pad_row),
          _p(g_table, F32, "g_table"), N, D, _stream())


def scatter_add_rows_deterministic(g, ids, pad_row, g_table):
    """g_table[ids[i]] += g[i] with a reproducible summation order: stable sort of the ids (torch.sort: plumbing),
    then one warp per distinct row sums its run in position order and adds it with a plain read-modify-write
    (rbx_segment_sum_rows).  Bit-identical from run to run; slower than the atomic kernel, opt-in."""
    flat = ids.reshape(-1)
    sorted_ids, order = torch.sort(flat, stable=True)
    N, D = flat.numel(), g_table.shape[1]
    _call("rbx_segment_sum_rows", _p(g.reshape(N, D) if g.is_contiguous() else g.contiguous().view(N, D), F32, "g"), _p(order, torch.int64, "order"),
          _p(sorted_ids, I32, "sorted_ids"), -1 if pad_row is None else int(pad_row), _p(g_table, F32, "g_table"), N, D, _stream())


def embed_fm_bwd_deterministic(table, rows, cat_pos, pad_row, E, S, dE, d_fm, d_lr, g_table, g_table_lr):
    """Reproducible table gradients of the fused FM backward (the atomic kernel's sums depend on the order the red.adds land):
    the per-slot gradient rows g = dE + d_fm * (S - e) are materialised (element-wise torch ops: plumbing) and reduced per
    table row by the deterministic scatter above; first-order rows likewise.  Numeric-slot / bias gradients are plain batch
    sums (torch.sum is deterministic).  Opt-in, ~3x the traffic of rbx_embed_fm_bwd."""
    B, F = rows.shape
    D = table.shape[1]
    pos = torch.as_tensor(list(cat_pos), device=rows.device)
    e = E.index_select(1, pos) if E is not None else gather_rows(table, rows.clamp_min(0))
    G = dE.index_select(1, pos) if dE is not None else torch.zeros_like(e)
    if d_fm is not None and S is not None:
        G = G + d_fm.view(-1, 1, 1) * (S.view(B, 1, D) - e)
    flat_rows = rows.reshape(-1)
    pads = torch.as_tensor([p if p is not None else -1 for p in pad_row], device=rows.device, dtype=rows.dtype).repeat(B)
    keep_rows = torch.where(flat_rows == pads, torch.full_like(flat_rows, -1), flat_rows)       # padding rows get no gradient
    scatter_add_rows_deterministic(G.reshape(B * F, D).contiguous(), keep_rows, -1, g_table)
    if d_lr is not None and g_table_lr is not None:
        g1 = d_lr.view(B, 1).expand(B, F).reshape(B * F, 1).contiguous()
        scatter_add_rows_deterministic(g1, keep_rows, -1, g_table_lr.view(-1, 1))


def pooled_gather_fwd(table, ids, mode, out=None, want_cnt=None):
    """ids int32 [B, L] (row stride may exceed L) -> (out [B,D], cnt [B] | None)."""
    if ids.dim() != 2 or ids.dtype != I32 or not ids.is_cuda or ids.stride(1) != 1:
        raise RbxError("pooled_gather_fwd: ids must be a CUDA int32 matrix with unit column stride")
    B, L = ids.shape
    D = table.shape[1]
    if out is None:
        out = torch.empty((B, D), dtype=F32, device=table.device)
    if out.stride(-1) != 1:
        raise RbxError("pooled_gather_fwd: out must have unit inner stride")
    want_cnt = (mode == 1) if want_cnt is None else want_cnt
    cnt = torch.empty((B,), dtype=F32, device=table.device) if want_cnt else None
    _call("rbx_pooled_gather_fwd", _p(table, F32, "table"), ctypes.c_void_p(ids.data_ptr()), ids.stride(0),
          ctypes.c_void_p(out.data_ptr()), out.stride(0), _p(cnt), B, L, D, mode, _stream())
    return out, cnt


def pooled_gather_bwd(g, ids, cnt, pad_row, g_table, mode):
    B, L = ids.shape
    D = g_table.shape[1]
    if g.stride(-1) != 1:
        raise RbxError("pooled_gather_bwd: g must have unit inner stride")
    _call("rbx_pooled_gather_bwd", ctypes.c_void_p(g.data_ptr()), g.stride(0), ctypes.c_void_p(ids.data_ptr()),
          ids.stride(0), _p(cnt, F32, "cnt"), -1 if pad_row is None else int(pad_row), _p(g_table, F32, "g_table"),
          B, L, D, mode, _stream())


def pool_fwd(emb, mask, mode):
    """materialised emb [B,L,D] (+ optional bool/uint8 mask [B,L]) -> (out [B,D], cnt [B] | None)."""
    B, L, D = emb.shape
    out = torch.empty((B, D), dtype=F32, device=emb.device)
    cnt = torch.empty((B,), dtype=F32, device=emb.device) if mode == 1 else None
    if mask is not None:
        mask = mask.to(torch.uint8).contiguous()
        if tuple(mask.shape) != (B, L):
            raise RbxError("pool_fwd: mask must be [B, L]")
    _call("rbx_pool_fwd", _p(emb, F32, "emb"), _p(mask, torch.uint8, "mask"), _p(out), _p(cnt), B, L, D, mode, _stream())
    return out, cnt


def pool_bwd(g, cnt, shape, mode):
    B, L, D = shape
    d_emb = torch.empty(shape, dtype=F32, device=g.device)
    _call("rbx_pool_bwd", _p(g, F32, "g"), _p(cnt, F32, "cnt"), _p(d_emb), B, L, D, mode, _stream())
    return d_emb


# ------------------------------------------------------------------------------------------ a10
def rowdot_fwd(u, v):
    B, D = u.shape
    K = v.numel() // (B * D) if B * D else 0
    y = torch.empty((B, K), dtype=F32, device=u.device)
    _call("rbx_rowdot_fwd", _p(u, F32, "u"), _p(v, F32, "v"), _p(y), B, K, D, _stream())
    return y


def rowdot_bwd(u, v, dy, need_du=True, need_dv=True):
    B, D = u.shape
    K = dy.shape[1]
    du = torch.empty_like(u) if need_du else None
    dv = torch.empty_like(v) if need_dv else None
    _call("rbx_rowdot_bwd", _p(u, F32, "u"), _p(v, F32, "v"), _p(dy, F32, "dy"), _p(du), _p(dv), B, K, D, _stream())
    return du, dv


# ------------------------------------------------------------------------------------------- a6
def interact_out_shape(B, F, D, mode):
    if mode not in (0, 1, 2, 3):
        raise RbxError("InnerProductInteraction mode %r is not supported" % (mode,))
    P = F * (F - 1) // 2
    return {0: (B, 1), 1: (B, D), 2: (B, P), 3: (B, P, D)}[mode]


def interact_fwd(E, mode):
    B, F, D = E.shape
    out = torch.empty(interact_out_shape(B, F, D, mode), dtype=F32, device=E.device)
    _call("rbx_interact_fwd", _p(E, F32, "E"), _p(out), B, F, D, mode, _stream())
    return out


def interact_bwd(E, dout, mode):
    B, F, D = E.shape
    dE = torch.empty_like(E)
    _call("rbx_interact_bwd", _p(E, F32, "E"), _p(dout, F32, "dout"), _p(dE), B, F, D, mode, _stream())
    return dE


# ------------------------------------------------------------------------------------------- (e)
def shard_set_rank(rank):
    """Tell the library which shard this process owns (hint only; see rbx_shard_set_rank)."""
    _call("rbx_shard_set_rank", int(rank))


def shard_route(rows, world):
    """flat int32 global rows -> (send local rows grouped by owner, pos, counts[world] int32 DEVICE)."""
    flat = rows.reshape(-1)
    N = flat.numel()
    lib = _lib.load()
    ws_bytes = lib.rbx_shard_ws_bytes(N, world)
    dev = rows.device
    ws = torch.empty((max(ws_bytes // 4, 1),), dtype=I32, device=dev)
    send = torch.empty((N,), dtype=I32, device=dev)
    pos = torch.empty((N,), dtype=I32, device=dev)
    counts = torch.empty((world,), dtype=I32, device=dev)
    _call("rbx_shard_route", _p(flat, I32, "rows"), N, world, _p(ws), ws_bytes, _p(send), _p(pos), _p(counts), _stream())
    return send, pos, counts


def shard_permute(x, pos):
    N = pos.numel()
    D = x.numel() // N if N else 1
    out = torch.empty((N, D), dtype=F32, device=x.device)
    _call("rbx_shard_permute", _p(x, F32, "x"), _p(pos, I32, "pos"), _p(out), N, D, _stream())
    return out


def shard_unroute(recv, pos):
    N = pos.numel()
    D = recv.numel() // N if N else 1
    out = torch.empty((N, D), dtype=F32, device=recv.device)
    _call("rbx_shard_unroute", _p(recv, F32, "recv"), _p(pos, I32, "pos"), _p(out), N, D, _stream())
    return out


def shard_push_ids(send, counts, inbox_ids_ptrs, inbox_meta_ptrs, rank, world, cap):
    _call("rbx_shard_push_ids", _p(send, I32, "send"), _p(counts, I32, "counts"), inbox_ids_ptrs, inbox_meta_ptrs,
          rank, world, cap, send.numel(), _stream())


def shard_serve_rows(table, table_lr, inbox_ids, inbox_meta, out_rows_ptrs, out_lr_ptrs, world, cap):
    _call("rbx_shard_serve_rows", _p(table, F32, "table"), _p(table_lr, F32, "table_lr"), table.shape[1],
          _p(inbox_ids, I32, "inbox_ids"), _p(inbox_meta, I32, "inbox_meta"), out_rows_ptrs, out_lr_ptrs, world, cap,
          _stream())


def shard_push_grads(gsend, gsend_lr, counts, ginbox_ptrs, ginbox_lr_ptrs, rank, world, cap):
    N, D = gsend.shape
    _call("rbx_shard_push_grads", _p(gsend, F32, "gsend"), _p(gsend_lr, F32, "gsend_lr"), _p(counts, I32, "counts"),
          ginbox_ptrs, ginbox_lr_ptrs, rank, world, cap, D, N, _stream())


def shard_apply_grads(ginbox, ginbox_lr, inbox_ids, inbox_meta, g_table, g_table_lr, world, cap, pad_local=None):
    D = g_table.shape[1]
    _call("rbx_shard_apply_grads", _p(ginbox, F32, "ginbox"), _p(ginbox_lr, F32, "ginbox_lr"), _p(inbox_ids, I32, "inbox_ids"),
          _p(inbox_meta, I32, "inbox_meta"), _p(g_table, F32, "g_table"), _p(g_table_lr, F32, "g_table_lr"), world, cap, D,
          _p(pad_local, I32, "pad_local"), 0 if pad_local is None else pad_local.numel(), _stream())


# ---- streamed exchange (csrc/shard_stream.cu): only contiguous runs cross NVLink ---------------------------------
def xs_tile_samples(F, D):
    return int(_lib.load().rbx_xs_tile_samples(int(F), int(D)))


def _lr_of(phys, D, lr_vec, lr_in_row):
    """(pointer, element stride) of the first-order weight of local row r: a separate vector, or column D of the row."""
    if lr_vec is not None:
        return _p(lr_vec, F32, "lr"), 1
    if lr_in_row:
        return ctypes.c_void_p(phys.data_ptr() + 4 * D), phys.shape[1]
    return None, 0


def xs_route(rows, R, D, rank, world, cap, cursor, tile_base, tile_cnt, pair_sorted, overflow, inbox_ids_ptrs):
    B, F = rows.shape
    _call("rbx_xs_route", _p(rows, I32, "rows"), B, F, int(R), int(D), rank, world, int(cap), _p(cursor, I32, "cursor"),
          _p(tile_base, I32, "tile_base"), _p(tile_cnt, I32, "tile_cnt"), _p(pair_sorted, torch.int16, "pair_sorted"),
          _p(overflow, I32, "overflow"), inbox_ids_ptrs, _stream())


def xs_barrier(flags_ptrs, meta_ptrs, cursor, rank, world, epoch, device):
    _note_device(torch.device(device), "device")
    _call("rbx_xs_barrier", flags_ptrs, meta_ptrs, _p(cursor, I32, "cursor"), rank, world, int(epoch) & 0xFFFFFFFF, _stream())


def xs_serve(phys, D, lr_vec, lr_in_row, inbox_ids, meta, cap, rank, world, rowbuf_ptrs, rowbuf_lr_ptrs):
    lr_ptr, lr_stride = _lr_of(phys, D, lr_vec, lr_in_row)
    _call("rbx_xs_serve", _p(phys, F32, "table"), phys.shape[1], lr_ptr, lr_stride, int(D), _p(inbox_ids, I32, "inbox_ids"),
          _p(meta, I32, "meta"), int(cap), rank, world, rowbuf_ptrs, rowbuf_lr_ptrs if lr_ptr is not None else None, _stream())


def xs_consume(rowbuf, rowbuf_lr, tile_base, tile_cnt, pair_sorted, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias,
               B, cap, D, world, want_E=True, want_lr=True, num_widx=None, n_slots=None, out=None):
    """out: (E | None, S, fm, lr | None) preallocated (contiguous slices of the full batch's outputs), else allocated."""
    F, Fn = len(cat_pos), len(num_pos)
    Ft = n_slots or (F + Fn)
    dev = rowbuf.device
    if out is not None:
        E, S, fm, lr = out
    else:
        E = torch.empty((B, Ft, D), dtype=F32, device=dev) if want_E else None
        S = torch.empty((B, D), dtype=F32, device=dev)
        fm = torch.empty((B,), dtype=F32, device=dev)
        lr = torch.empty((B,), dtype=F32, device=dev) if want_lr else None
    _call("rbx_xs_consume", _p(rowbuf, F32, "rowbuf"), _p(rowbuf_lr, F32, "rowbuf_lr") if want_lr else None,
          _p(tile_base, I32, "tile_base"), _p(tile_cnt, I32, "tile_cnt"), _p(pair_sorted, torch.int16, "pair_sorted"),
          _i32(cat_pos), _p(dense_x, F32, "dense_x"), _p(dense_w, F32, "dense_w"), _p(dense_w_lr, F32, "dense_w_lr"),
          _i32(num_pos), _i32(num_widx) if num_widx is not None else None, _p(lr_bias, F32, "lr_bias"),
          _p(E, F32, "E"), _p(S, F32, "S"), _p(fm, F32, "fm"), _p(lr, F32, "lr"), B, int(cap), F, Fn, int(D), Ft, world, _stream())
    return E, S, fm, lr


def xs_grad_push(E, rowbuf, S, dE, d_fm, d_lr, rows, pad_row, tile_base, tile_cnt, pair_sorted, cat_pos, cap, D, n_slots,
                 rank, world, ginbox_ptrs, ginbox_lr_ptrs, dense_x=None, dense_w=None, num_pos=(), num_widx=None,
                 g_dense_w=None, g_dense_w_lr=None, g_lr_bias=None):
    B, F = rows.shape
    Fn = len(num_pos)
    _call("rbx_xs_grad_push", _p(E, F32, "E"), _p(rowbuf, F32, "rowbuf"), _p(S, F32, "S"), _p(dE, F32, "dE"),
          _p(d_fm, F32, "d_fm"), _p(d_lr, F32, "d_lr"), _p(rows, I32, "rows"),
          _i32(pad_row if pad_row is not None else [-1] * F), _p(tile_base, I32, "tile_base"),
          _p(tile_cnt, I32, "tile_cnt"), _p(pair_sorted, torch.int16, "pair_sorted"), _i32(cat_pos),
          _p(dense_x, F32, "dense_x"), _p(dense_w, F32, "dense_w"), _i32(num_pos),
          _i32(num_widx) if num_widx is not None else None, Fn, _p(g_dense_w, F32, "g_dense_w"),
          _p(g_dense_w_lr, F32, "g_dense_w_lr"), _p(g_lr_bias, F32, "g_lr_bias"), B, int(cap), F, int(D),
          int(n_slots), rank, world, ginbox_ptrs, ginbox_lr_ptrs, _stream())


def xs_apply(ginbox, ginbox_lr, inbox_ids, meta, cap, world, g_phys, D, g_lr_vec, lr_in_row):
    lr_ptr, lr_stride = _lr_of(g_phys, D, g_lr_vec, lr_in_row)
    _call("rbx_xs_apply", _p(ginbox, F32, "ginbox"), _p(ginbox_lr, F32, "ginbox_lr") if lr_ptr is not None else None,
          _p(inbox_ids, I32, "inbox_ids"), _p(meta, I32, "meta"), int(cap), world, _p(g_phys, F32, "g_table"),
          g_phys.shape[1], lr_ptr, lr_stride, int(D), _stream())


# ------------------------------------------------------------------------------------------ a12
def sqnorm_(g, acc):
    """acc (float64 [1], device) += sum(g^2)."""
    _call("rbx_sqnorm", _p(g, F32, "g"), g.numel(), _p(acc, torch.float64, "acc"), _stream())


def clip_coef(acc, max_norm, coef, norm_out=None):
    _call("rbx_clip_coef", _p(acc, torch.float64, "acc"), float(max_norm), _p(coef, F32, "coef"), _p(norm_out), _stream())


def adam_dense_(w, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip=None):
    _call("rbx_adam_dense", _p(w, F32, "w"), _p(g, F32, "g"), _p(m, F32, "m"), _p(v, F32, "v"), w.numel(),
          _p(clip, F32, "clip"), float(lr), float(beta1), float(beta2), float(eps), int(step), _stream())


# ------------------------------------------------------------------------------------------- f1
OPTIM_KINDS = {"sgd": 0, "adagrad": 1, "adam_rows": 2, "sparse_adam": 3}


def sqnorm_rows_(g, rows, n_rows, acc):
    """acc (float64 [1]) += sum over the touched `rows` (int32, first n_rows[0] valid) of |g[row,:]|^2."""
    D = 1 if g.dim() == 1 else g.shape[1]
    _call("rbx_sqnorm_rows", _p(g, F32, "g"), _p(rows, I32, "rows"), _p(n_rows, torch.int64, "n_rows"), rows.numel(), D,
          _p(acc, torch.float64, "acc"), _stream())


def optim_rows_(w, g, m, v, rows, n_rows, step, kind="adam_rows", lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, clip=None,
                zero_grad=True):
    """Touched-rows optimizer update of table w [R,D] (or [R]) in place; see rbx_optim_rows for `kind`."""
    if kind not in OPTIM_KINDS:
        raise RbxError("optim_rows_: kind must be one of %s" % sorted(OPTIM_KINDS))
    D = 1 if w.dim() == 1 else w.shape[1]
    _call("rbx_optim_rows", _p(w, F32, "w"), _p(g, F32, "g"), _p(m, F32, "m"), _p(v, F32, "v"), _p(rows, I32, "rows"),
          _p(n_rows, torch.int64, "n_rows"), rows.numel(), D, _p(clip, F32, "clip"), OPTIM_KINDS[kind], float(lr),
          float(beta1), float(beta2), float(eps), int(step), int(bool(zero_grad)), _stream())


# ------------------------------------------------------------------------------------------ a14
_unique_ws = {}


def unique_ids(ids, vocab, want_first=True, want_inverse=True, sync=True):
    """Sorted unique ids, first flat position of each, inverse map -- ids int64 or int32 CUDA tensor of
    any shape, every id in [0, vocab).  sync=True trims the outputs to the U uniques found (one 16-byte
    D2H read, as torch.unique does); sync=False returns full-capacity buffers plus the device counter
    n_out (int64 [2]: U, #out-of-range) for consumers that read the count on the device."""
    if not ids.is_cuda:
        raise RbxError("ids must be a CUDA tensor (recbox_b200 has no CPU path)")
    if ids.dtype not in (torch.int64, torch.int32):
        raise RbxError("ids must be int64 or int32, got %s" % ids.dtype)
    flat = ids.contiguous().view(-1)
    n, vocab = flat.numel(), int(vocab)
    lib = _lib.load()
    need = int(lib.rbx_unique_ws_bytes(vocab))
    key = (flat.device, vocab)
    ws = _unique_ws.get(key)
    if ws is None or ws.numel() * 8 < need:
        ws = _unique_ws[key] = torch.empty((need + 7) // 8, dtype=torch.int64, device=flat.device)
    cap = max(min(n, vocab), 1)
    uniq = torch.empty(cap, dtype=flat.dtype, device=flat.device)
    first = torch.empty(cap, dtype=torch.int64, device=flat.device) if want_first else None
    inverse = torch.empty(max(n, 1), dtype=flat.dtype, device=flat.device)[:n] if want_inverse else None
    n_out = torch.empty(2, dtype=torch.int64, device=flat.device)
    name = "rbx_unique_ids_i64" if flat.dtype == torch.int64 else "rbx_unique_ids_i32"
    _call(name, _p(flat), n, vocab, _p(ws), ws.numel() * 8, _p(uniq), _p(first), _p(inverse), _p(n_out), _stream())
    if not sync:
        return uniq, first, inverse, n_out
    U, bad = (int(x) for x in n_out.tolist())
    if bad:
        raise RbxError("unique_ids: %d id(s) outside [0, %d)" % (bad, vocab))
    return uniq[:U], (first[:U] if want_first else None), inverse


# ------------------------------------------------------------------------------------------- f4
def power_sums_fwd(E, order):
    """E [B,F,D] -> P [B,order,D], P[:,k-1] = (E ** k).sum(1)  (InteractionMachine's p_1..p_order in one pass)."""
    B, F, D = E.shape
    P = torch.empty((B, order, D), dtype=F32, device=E.device)
    _call("rbx_power_sums_fwd", _p(E, F32, "E"), _p(P), B, F, D, int(order), _stream())
    return P


def power_sums_bwd(E, dP):
    B, F, D = E.shape
    dE = torch.empty_like(E)
    _call("rbx_power_sums_bwd", _p(E, F32, "E"), _p(dP, F32, "dP"), _p(dE), B, F, D, dP.shape[1], _stream())
    return dE


# ------------------------------------------------------------------------------------------- f2
def sample_negatives(n_queries, num_negs, num_items, seed, pos=None, user_of_query=None, pos_ptr=None, pos_items=None,
                     device=None):
    """[n_queries, (pos is not None) + num_negs] int64: the positive column (when given) then uniform negatives; with
    the CSR (pos_ptr, pos_items: each user's interacted items, sorted) draws hitting the query user's items are redrawn
    (h5_generator.py:72-95 ignore_pos_items).  Returns (out, gave_up int32[1] | None)."""
    I64 = torch.int64
    dev = device if device is not None else (pos.device if pos is not None else pos_ptr.device)
    lead = 0 if pos is None else 1
    out = torch.empty((n_queries, lead + num_negs), dtype=I64, device=dev)
    if not out.is_cuda:
        raise RbxError("sample_negatives needs a CUDA device (recbox_b200 has no CPU path)")
    gave_up = torch.zeros(1, dtype=I32, device=dev) if pos_ptr is not None else None
    _call("rbx_sample_negatives", int(n_queries), int(num_negs), int(num_items), ctypes.c_uint64(int(seed) & (2 ** 64 - 1)),
          _p(pos, I64, "pos"), _p(user_of_query, I64, "user_of_query"), _p(pos_ptr, I64, "pos_ptr"),
          _p(pos_items, I64, "pos_items"), _p(out), _p(gave_up), _stream())
    return out, gave_up


# ------------------------------------------------------------------------------------------- f3
_topk_ws = {}


def topk_ip(q, items, k, chunk=None, want_scores=True):
    """Exact inner-product top-k of q [U,D] against items [N,D] (faiss.IndexFlatIP.search): (scores [U,k] desc | None,
    idx [U,k] int64)."""
    if q.dim() != 2 or items.dim() != 2 or q.shape[1] != items.shape[1]:
        raise RbxError("topk_ip: q [U,D] and items [N,D] must share D")
    U, D = q.shape
    N = items.shape[0]
    lib = _lib.load()
    if chunk is None:
        chunk = max(4096, min(131072, (1 << 27) // max(U, 1) // 128 * 128))      # <= 1 GiB of candidate queue
    need = int(lib.rbx_topk_ws_bytes(U, int(k), int(chunk))) if U else 0
    if U and need == 0:
        raise RbxError("topk_ip: k=%d outside [1, 1024]" % k)
    ws = _topk_ws.get(q.device)
    if ws is None or ws.numel() < need:
        _topk_ws[q.device] = None
        ws = _topk_ws[q.device] = torch.empty(max(need, 256), dtype=torch.uint8, device=q.device)
    scores = torch.empty((U, k), dtype=F32, device=q.device) if want_scores else None
    idx = torch.empty((U, k), dtype=torch.int64, device=q.device)
    _call("rbx_topk_ip", _p(q, F32, "q"), _p(items, F32, "items"), U, N, D, int(k), int(chunk), _p(scores), _p(idx), _p(ws),
          ws.numel(), _stream())
    return scores, idx


METRIC_KINDS = {"Recall": 0, "nRecall": 1, "Precision": 2, "F1": 3, "DCG": 4, "NDCG": 5, "MRR": 6, "HitRate": 7, "MAP": 8}


def rank_metrics(cand, train_ptr, train_items, valid_ptr, valid_items, kinds, ks, kmax=None):
    """cand [U,T] int64 top-T ids (descending) -> (ranked [U,kmax], hit [U,kmax] uint8, per-user metrics [U,M] float64)
    after pushing the user's train items behind the others (core/metrics.py:52-68) -- see rbx_rank_metrics."""
    I64 = torch.int64
    U, T = cand.shape
    M = len(kinds)
    kmax = int(kmax or (max(ks) if M else T))
    dev = cand.device
    ranked = torch.empty((U, kmax), dtype=I64, device=dev)
    hit = torch.empty((U, kmax), dtype=torch.uint8, device=dev)
    out = torch.empty((U, M), dtype=torch.float64, device=dev) if M else None
    kd = torch.tensor(list(kinds), dtype=I32, device=dev) if M else None
    kk = torch.tensor(list(ks), dtype=I32, device=dev) if M else None
    _call("rbx_rank_metrics", _p(cand, I64, "cand"), T, U, _p(train_ptr, I64, "train_ptr"), _p(train_items, I64, "train_items"),
          _p(valid_ptr, I64, "valid_ptr"), _p(valid_items, I64, "valid_items"), kmax, _p(kd), _p(kk), M, _p(ranked), _p(hit),
          _p(out), _stream())
    return ranked, hit, out


# ------------------------------------------------------------------------------------------- a13
GEMM_PRECISION = 3       # 3 = 3xTF32 (fp32-level, the 1e-5 contract); 1 = plain TF32 (stated option, ~1e-3 rel)


def _mat(t, name):
    """2-D fp32 CUDA tensor whose rows are contiguous -> (pointer, pitch in floats)."""
    if t.dim() != 2 or t.dtype != F32 or not t.is_cuda or (t.shape[1] > 1 and t.stride(1) != 1):
        raise RbxError("%s must be a 2-D fp32 CUDA tensor with contiguous rows" % name)
    _note_device(t.device, name)
    return ctypes.c_void_p(t.data_ptr()), (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1))


def gemm(a, b, a_mn=False, b_mn=False, bias=None, relu=False, mask=None, out=None, accumulate=False, precision=None):
    """out[M,N] (+)= relu?(A B^T + bias) * (mask > 0)?   (rbx_gemm_f32, tcgen05 tensor cores).
    a: [M,K] (a_mn: stored [K,M]); b: [N,K] as nn.Linear.weight (b_mn: stored [K,N])."""
    (M, K) = (a.shape[1], a.shape[0]) if a_mn else a.shape
    (N, Kb) = (b.shape[1], b.shape[0]) if b_mn else b.shape
    if K != Kb:
        raise RbxError("gemm: inner dimensions differ (%d vs %d)" % (K, Kb))
    if out is None:
        out = torch.empty((M, N), dtype=F32, device=a.device)
        accumulate = False
    elif tuple(out.shape) != (M, N):
        raise RbxError("gemm: out must be [%d, %d]" % (M, N))
    if mask is not None and tuple(mask.shape) != (M, N):
        raise RbxError("gemm: mask must be [%d, %d]" % (M, N))
    pa, lda = _mat(a, "a")
    pb, ldb = _mat(b, "b")
    pc, ldc = _mat(out, "out")
    pm, ldm = _mat(mask, "mask") if mask is not None else (None, 0)
    _call("rbx_gemm_f32", pa, lda, 1 if a_mn else 0, pb, ldb, 1 if b_mn else 0, pc, ldc, M, N, K, _p(bias, F32, "bias"), 1 if relu else 0,
          pm, ldm, int(precision or GEMM_PRECISION), 1 if accumulate else 0, _stream())
    return out


def colsum(x, out=None, accumulate=False):
    """out[n] (+)= sum_m x[m, n]  (bias gradient)."""
    M, N = x.shape
    if out is None:
        out = torch.empty((N,), dtype=F32, device=x.device)
        accumulate = False
    px, ldx = _mat(x, "x")
    _call("rbx_colsum_f32", px, ldx, _p(out, F32, "out"), M, N, 1 if accumulate else 0, _stream())
    return out
