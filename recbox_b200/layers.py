"""The reference's layer / operator API on the fused B200 path.

Same class names, constructor arguments, state_dict keys and output shapes as

  recbox/ranking/pytorch/layers : FeatureEmbedding, FeatureEmbeddingDict        (embeddings/feature_embedding.py:28-214)
                                  InnerProductInteraction                       (interactions/inner_product.py:22-56)
                                  LogisticRegression, FactorizationMachine      (blocks/logistic_regression.py:23-35,
                                                                                 blocks/factorization_machine.py:24-34)
                                  MaskedAveragePooling, MaskedSumPooling        (pooling.py:22-40)
  recbox/core/pytorch/layers    : EmbeddingLayer, EmbeddingDictLayer            (embedding.py:10-138)
                                  MaskedAveragePooling, MaskedSumPooling        (sequence.py:4-20)   [Core* classes]

so that existing model code and configs load unchanged (`recbox_b200.install()` rebinds the
reference's symbols).  What differs is underneath:

  * every per-feature nn.Embedding / nn.Linear(1, D) is still there under
    `embedding_layers.<feature>` (same parameter names, same init order under a given seed, same
    `type(v) == nn.Embedding` the reference's init / regulariser code scans for), but all of them
    are VIEWS into one fused [sum V_f, D] table (+ one [Fn, D] block for the numeric slots);
  * a forward is one launch of the fused gather (+ FM product_sum + LR) kernel writing [B, F, D]
    in place -- no per-feature lookups, no torch.stack; the backward is one scatter-add launch
    whose dense gradients are handed to autograd as views of one fused gradient buffer;
  * FactorizationMachine / InnerProductInteraction("product_sum") / LogisticRegression called on
    the tensors produced here return the values the same launch already computed.

There is no CPU path: the modules raise if their parameters or inputs are not on a CUDA device.
"""
import weakref
from collections import OrderedDict
from functools import partial  # noqa: F401  (used by eval'd initializer strings, as in the reference)

import numpy as np
import torch
from torch import nn

from . import ops
from ._lib import RbxError
from .loader import PackedBatch, PackedColumns

I32, F32 = torch.int32, torch.float32


# =================================================================================================
# pooling layers (feature_encoder / embedding_callback targets)
# =================================================================================================
class _PoolFn(torch.autograd.Function):
    """[B,L,D] -> [B,D] pooling of an already materialised tensor (someone calls the pooling module
    directly on an embedding tensor); the fused id -> pooled path lives in _PooledGatherFn."""

    @staticmethod
    def forward(ctx, emb, mask, average):
        if not emb.is_cuda:
            raise RbxError("recbox_b200 layers have no CPU path")
        mode = 1 if average else 0
        out, cnt = ops.pool_fwd(emb.contiguous(), mask, mode)
        ctx.save_for_backward(cnt)
        ctx.meta = (tuple(emb.shape), mode)
        return out

    @staticmethod
    def backward(ctx, g):
        (cnt,) = ctx.saved_tensors
        shape, mode = ctx.meta
        return ops.pool_bwd(g.contiguous(), cnt, shape, mode), None, None


class MaskedAveragePooling(nn.Module):
    """ranking/pytorch/layers/pooling.py:22-31: sum over L / (#rows with non-zero element sum + 1e-12)."""

    def __init__(self):
        super(MaskedAveragePooling, self).__init__()

    def forward(self, embedding_matrix, mask=None):
        return _PoolFn.apply(embedding_matrix, mask, True)


class MaskedSumPooling(nn.Module):
    """ranking/pytorch/layers/pooling.py:34-40."""

    def __init__(self):
        super(MaskedSumPooling, self).__init__()

    def forward(self, embedding_matrix):
        return _PoolFn.apply(embedding_matrix, None, False)


class CoreMaskedAveragePooling(MaskedAveragePooling):
    """core/pytorch/layers/sequence.py:4-12 (no mask argument)."""

    def forward(self, embedding_matrix):
        return _PoolFn.apply(embedding_matrix, None, True)


CoreMaskedSumPooling = MaskedSumPooling


def _pool_mode(module):
    """0 = sum, 1 = masked average, None = not one of ours (generic encoder)."""
    if type(module) in (MaskedAveragePooling, CoreMaskedAveragePooling):
        return 1
    if type(module) is MaskedSumPooling:
        return 0
    return None


# =================================================================================================
# fused parameter storage
# =================================================================================================
_FUSE_GEN = [0]          # bumped by every _FusedStore.fuse(); launch-plan caches are valid for one generation


class _Group(object):
    """All unique embedding modules of one embedding dim, stored back to back."""

    def __init__(self, D):
        self.D = D
        self.emb, self.emb_off, self.R = [], {}, 0        # nn.Embedding modules, id(module) -> row offset
        self.lin, self.lin_idx = [], {}                   # nn.Linear(1, D) modules, id(module) -> row of dense_w
        self.table = None                                  # [R, D]
        self.dense_w = None                                # [Fn, D]
        self.ptrs = None


class _FusedStore(object):
    """Keeps the Parameters of a ModuleDict of nn.Embedding / nn.Linear(1, D) pointing into one
    allocation per embedding dim.  Parameter identity never changes (optimizers, state_dict,
    named_parameters, DDP hooks all see the reference's per-feature parameters)."""

    def __init__(self, modules):
        self.groups = OrderedDict()
        seen = set()
        for m in modules:
            if id(m) in seen:
                continue
            seen.add(id(m))
            if isinstance(m, nn.Embedding):
                g = self.groups.setdefault(m.embedding_dim, _Group(m.embedding_dim))
                g.emb_off[id(m)] = g.R
                g.emb.append(m)
                g.R += m.num_embeddings
            elif isinstance(m, nn.Linear):
                g = self.groups.setdefault(m.out_features, _Group(m.out_features))
                g.lin_idx[id(m)] = len(g.lin)
                g.lin.append(m)
        self.fuse()

    def fuse(self):
        _FUSE_GEN[0] += 1              # every cached launch plan holds Parameter objects / offsets of the old fusion
        for g in self.groups.values():
            ref = (g.emb[0] if g.emb else g.lin[0]).weight
            dev, dt = ref.device, ref.dtype
            if g.R > 2 ** 31 - 1:
                raise RbxError("fused table of %d rows exceeds int32 row ids" % g.R)
            if g.emb:
                table = torch.empty((g.R, g.D), dtype=dt, device=dev)
                for m in g.emb:
                    o = g.emb_off[id(m)]
                    view = table[o:o + m.num_embeddings]
                    view.copy_(m.weight.data)
                    m.weight.data = view
                g.table = table
            if g.lin:
                dw = torch.empty((len(g.lin), g.D), dtype=dt, device=dev)
                for m in g.lin:
                    i = g.lin_idx[id(m)]
                    dw[i].copy_(m.weight.data.reshape(-1))
                    m.weight.data = dw[i].view(g.D, 1)
                g.dense_w = dw
            g.ptrs = self._expected(g)

    @staticmethod
    def _expected(g):
        out = []
        if g.emb:
            base, es = g.table.data_ptr(), g.table.element_size()
            out += [base + g.emb_off[id(m)] * g.D * es for m in g.emb]
        if g.lin:
            base, es = g.dense_w.data_ptr(), g.dense_w.element_size()
            out += [base + g.lin_idx[id(m)] * g.D * es for m in g.lin]
        return out

    def ensure(self):
        """Re-fuse if someone re-pointed a parameter (module.to(), weight = Parameter(...), ...)."""
        for g in self.groups.values():
            if [m.weight.data_ptr() for m in g.emb] + [m.weight.data_ptr() for m in g.lin] != g.ptrs:
                self.fuse()
                return


# =================================================================================================
# row-sharded storage behind the same modules (BASELINE configs[3])
# =================================================================================================
_SHARD_CFG = [None]
_SHARD_DIMS = (4, 8, 16, 32, 64, 128)         # physical row widths the sharded kernels cover (csrc/embed_fm.cu)


class sharded_tables(object):
    """Switch (context manager) under which FeatureEmbeddingDict / EmbeddingDictLayer keep their embedding tables
    ROW-SHARDED across the ranks of `group` (recbox_b200.sharded.ShardedEmbeddingFM: global fused row r lives on rank
    r % world) instead of replicated:

        with recbox_b200.layers.sharded_tables():          # or RECBOX_B200_SHARD=peer in the environment
            model = DeepFM(feature_map, ...)                # unmodified model code, reference ctor arguments
        model.load_state_dict(reference_checkpoint)         # full tables, reference key names -- still on the host
        model.to(device)                                    # <- the tables are cut here (collective over `group`)

    After the cut `embedding_layers.<feature>.weight` is this rank's [rows owned, D] shard (same parameter names;
    optimizers and state_dict see the shard; gather_state_dict() rebuilds the global one).  Every rank feeds its own batch
    shard; a training forward is: zero of the local gradient shard -> cross-rank fence -> fused gather over NVLink; the
    backward reduces straight into the owners' gradient shards, fences, and hands autograd views of the local shard.
    Replicated parameters (numeric-slot weights, LR bias, the dense tail) are the caller's to all-reduce
    (sync_replica_gradients skips the sharded ones).  Supported: one training forward / backward per embedding
    dictionary and step, plain categorical + numeric slots of one embedding dim (no pooled sequence slots)."""

    def __init__(self, mode="peer", group=None, alloc="symm", max_ids=None, slack=1.5, kern=None):
        self.mode, self.group, self.alloc, self.max_ids, self.slack, self.kern = mode, group, alloc, max_ids, slack, kern
        self._prev = None

    def __enter__(self):
        self._prev, _SHARD_CFG[0] = _SHARD_CFG[0], self
        return self

    def __exit__(self, *exc):
        _SHARD_CFG[0] = self._prev
        return False


def _shard_config():
    cfg = _SHARD_CFG[0]
    if cfg is None:
        import os
        mode = os.environ.get("RECBOX_B200_SHARD", "")
        if mode and mode != "0":
            cfg = sharded_tables(mode="peer" if mode == "1" else mode)
    return cfg


class _ShardedStore(_FusedStore):
    """_FusedStore whose embedding groups turn into row shards once the parameters reach a CUDA device (or on
    shard_now()).  Until then it IS a _FusedStore: construction, seeded init and load_state_dict of a full reference
    checkpoint run on the host exactly as in the replicated case, so the cut table equals the reference's row for row."""

    def __init__(self, modules, cfg):
        self.cfg = cfg
        self.sharded = False
        self.serial = 0                    # bumped by every training forward (= every zero-fill of the gradient shards)
        _FusedStore.__init__(self, modules)

    def fuse(self):
        if self.sharded:
            self.ensure()
            return
        ref = next((m.weight for g in self.groups.values() for m in (g.emb + g.lin)), None)
        if ref is not None and ref.device.type == "cuda":
            self.shard()
        else:
            _FusedStore.fuse(self)

    def ensure(self):
        if not self.sharded:
            return _FusedStore.ensure(self)
        for g in self.groups.values():
            if [m.weight.data_ptr() for m in g.emb] + [m.weight.data_ptr() for m in g.lin] != g.ptrs:
                raise RbxError("a row-sharded table stays where it was cut: parameters cannot be re-pointed or moved "
                               "afterwards (gather_state_dict() rebuilds the global tables)")

    def shard(self):
        """Cut every embedding group: allocate this rank's table / gradient shards (peer-visible), keep the rows
        r % world == rank of the fused table, re-point the per-feature Parameters at their slice.  Collective."""
        from . import sharded
        _FUSE_GEN[0] += 1
        cfg = self.cfg
        for g in self.groups.values():
            ref = (g.emb[0] if g.emb else g.lin[0]).weight
            dev = ref.device
            if ref.dtype != F32:
                raise RbxError("sharded tables are fp32")
            if dev.type != "cuda" and cfg.kern is None:
                raise RbxError("recbox_b200 layers have no CPU path: move the model to a CUDA device (that is where the tables are cut)")
            D = g.D
            Dp = D if D in _SHARD_DIMS else (4 if D < 4 else None)
            if Dp is None:
                raise RbxError("sharded tables cover embedding dims 1..4 and %s (got %d)" % (_SHARD_DIMS[1:], D))
            g.Dp, g.sem, g.local, g.store = Dp, None, [], self
            if g.emb:
                sem = sharded.ShardedEmbeddingFM(g.R, Dp, mode=cfg.mode, group=cfg.group, device=dev, with_lr=False,
                                                 kern=cfg.kern, max_ids=cfg.max_ids, slack=cfg.slack, alloc=cfg.alloc)
                W, rank = sem.world, sem.rank
                for m in g.emb:
                    o, V = g.emb_off[id(m)], m.num_embeddings
                    lo, hi = sharded.local_rows(o, W, rank), sharded.local_rows(o + V, W, rank)
                    view = sem.table[lo:hi, :D]
                    view.copy_(m.weight.data[(rank - o) % W::W])
                    m.weight.data = view
                    m.weight._rbx_shard = (self, g, m)
                    g.local.append((lo, hi))
                g.sem, g.table = sem, None
                sem.barrier()
            if g.lin:
                dw = torch.zeros((len(g.lin), Dp), dtype=F32, device=dev)
                for m in g.lin:
                    i = g.lin_idx[id(m)]
                    dw[i, :D].copy_(m.weight.data.reshape(-1))
                    m.weight.data = dw[i, :D].view(D, 1)
                g.dense_w = dw
            g.ptrs = [m.weight.data_ptr() for m in g.emb] + [m.weight.data_ptr() for m in g.lin]
        self.sharded = True

    def gather_param(self, g, m, grad=False):
        """-> the full [V, D] table of one feature (or its gradient), rebuilt from every rank's rows (collective)."""
        import torch.distributed as dist
        sem = g.sem
        W, local = sem.world, m.weight.data
        if grad:
            local = m.weight.grad if m.weight.grad is not None else torch.zeros_like(local)
        if W == 1:
            return local.clone()
        o, V, D = g.emb_off[id(m)], m.num_embeddings, g.D
        n_max = (V + W - 1) // W + 1
        mine = torch.zeros((n_max, D), dtype=F32, device=local.device)
        mine[:local.shape[0]] = local
        parts = [torch.empty_like(mine) for _ in range(W)]
        dist.all_gather(parts, mine, group=sem.group)
        full = torch.empty((V, D), dtype=F32, device=local.device)
        for w in range(W):
            first = (w - o) % W
            full[first::W] = parts[w][:len(range(first, V, W))]
        return full


def shard_now(module):
    """Cut the tables of every not-yet-sharded dictionary under `module` on the device they are on (Module.to(cuda)
    does this by itself; this is the explicit form).  Collective over the configured group."""
    for sub in module.modules():
        st = getattr(sub, "_store", None)
        if isinstance(st, _ShardedStore) and not st.sharded:
            st.shard()
            sub._calls, sub._calls_gen = {}, _FUSE_GEN[0]
    return module


def is_sharded(p):
    return getattr(p, "_rbx_shard", None) is not None


def gather_state_dict(module, grads=False):
    """state_dict() with every row-sharded parameter rebuilt to its global [V, D] shape under the reference's key names
    (checkpoints stay interchangeable with the replicated / reference model).  grads=True: the parameters' gradients
    instead (sharded ones rebuilt, the others as they are).  Collective."""
    named = OrderedDict(module.named_parameters(remove_duplicate=False))
    out = OrderedDict()
    for k, v in (named.items() if grads else module.state_dict().items()):
        p = named.get(k)
        if p is not None and is_sharded(p):
            store, g, m = p._rbx_shard
            out[k] = store.gather_param(g, m, grad=grads)
        elif grads:
            out[k] = None if v.grad is None else v.grad.detach().clone()
        else:
            out[k] = v.detach().clone()
    return out


def clip_grad_norm_(parameters, max_norm, group=None):
    """torch.nn.utils.clip_grad_norm_ (2-norm) for a model whose tables are row-sharded: the squared norms of the sharded
    gradients are summed over the ranks, those of the replicated ones (already all-reduced) counted once
    (RankingModel.train_step, ranking_model.py:191-197).  Returns the global norm."""
    import torch.distributed as dist
    ps = [p for p in parameters if p.grad is not None]
    if not ps:
        return torch.zeros(())
    dev = ps[0].grad.device
    sq_sh, sq_rep = torch.zeros((), dtype=F32, device=dev), torch.zeros((), dtype=F32, device=dev)
    for p in ps:
        s = p.grad.detach().float().pow(2).sum()
        if is_sharded(p):
            sq_sh = sq_sh + s
        else:
            sq_rep = sq_rep + s
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sq_sh, group=group)
    total = (sq_sh + sq_rep).sqrt()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for p in ps:
        p.grad.detach().mul_(coef)
    return total


def train_step(model, batch_data, group=None):
    """RankingModel.train_step (ranking_model.py:191-197: zero_grad, total loss, backward, clip_grad_norm_, optimizer step) for
    ONE model trained by several ranks -- row-sharded tables (sharded_tables) and / or replicas: every rank passes its equal-sized
    slice of the global batch; the slice's mean loss is scaled by 1 / world so that the table gradients (reduced into their
    owners' shards by the backward) and the all-reduced replicated gradients are those of the GLOBAL batch mean, the clip uses
    the global norm, and every rank takes the same optimizer step.  Returns the global mean loss.  Bind it over the reference's
    method (`model.train_step = functools.partial(layers.train_step, model)`) to keep `fit()` / `train_epoch()` unchanged.
    With one rank it is the reference's step.  Regularisers are not covered (their sharded / replicated parts scale differently)."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1 and (getattr(model, "_embedding_regularizer", None) or getattr(model, "_net_regularizer", None)):
        raise NotImplementedError("layers.train_step: embedding / net regularizers are not covered at world > 1")
    model.optimizer.zero_grad()
    loss = model.get_total_loss(batch_data)
    (loss / world if world > 1 else loss).backward()
    params = list(model.parameters())
    if world > 1:
        sync_replica_gradients(params, group=group)
    if any(is_sharded(p) for p in params):
        clip_grad_norm_(params, model._max_gradient_norm, group=group)
    else:
        nn.utils.clip_grad_norm_(params, model._max_gradient_norm)
    model.optimizer.step()
    if world > 1:
        loss = loss.detach().clone()
        dist.all_reduce(loss, group=group)
        loss = loss / world
    return loss


# =================================================================================================
# the autograd node around the two fused kernels
# =================================================================================================
class _Call(object):
    """Static description of one fused launch (built once per feature selection, cached)."""
    __slots__ = ("D", "F", "Fn", "Ft", "cat_pos", "num_pos", "num_widx", "pad_row", "lr_delta", "group",
                 "lr_group", "lr_bias", "params", "kinds", "emb_sizes", "lr_emb_sizes", "with_main", "with_lr", "seq")


ASYNC_ZERO = True        # zero-fill of the dense gradient buffer on a side stream, under the forward
_side_streams = {}


def _grad_layout(call, with_main, with_lr):
    """Sizes and offsets (floats, 16-byte aligned) of [table | numeric weights | lr table | lr numeric | lr bias] in the one
    gradient allocation of a fused launch."""
    g, lg, D = call.group, call.lr_group, call.D
    sizes = (g.R * D if (with_main and g is not None and g.emb) else 0,
             len(g.lin) * D if (with_main and g is not None and g.lin) else 0,
             lg.R if (with_lr and lg is not None and lg.emb) else 0,
             len(lg.lin) if (with_lr and lg is not None and lg.lin) else 0,
             1 if (with_lr and call.lr_bias is not None) else 0)
    offs, tot = [], 0
    for n in sizes:
        offs.append(tot)
        tot += (n + 3) // 4 * 4
    return sizes, offs, tot


def _prezero(layout, dev):
    """-> (buffer, event): buffer allocated on the current stream, zero-filled (rbx_zero_f32) on the device's side stream."""
    tot = max(layout[2], 1)
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(device=dev)
    buf = torch.empty(tot, dtype=F32, device=dev)
    here = torch.cuda.Event()
    here.record(main)                  # the block is free in `main`'s order from here on
    side.wait_event(here)
    with torch.cuda.stream(side):
        ops.zero_(buf)
    done = torch.cuda.Event()
    done.record(side)
    return buf, done


class _FusedEmbedFn(torch.autograd.Function):
    """E, fm, lr = fused(rows, dense_x; per-feature parameters).  The parameters are passed so that
    autograd tracks them; the kernels read the fused buffers they alias."""

    @staticmethod
    def forward(ctx, call, rows, dense_x, seq_ids, *params):
        g, lg = call.group, call.lr_group
        seq = call.seq or ()
        want_E = call.with_main
        want_fm = call.with_main and not seq          # pooled sequence slots are not in the kernel's FM sums
        want_lr = call.with_lr
        B = (rows if rows is not None else (dense_x if dense_x is not None else seq_ids[0])).shape[0]
        E, S, fm, lr = ops.embed_fm_fwd(
            g.table if (g is not None and call.F) else None,
            lg.table.view(-1) if (lg is not None and call.F) else None,
            rows, call.cat_pos, dense_x,
            g.dense_w if (g is not None and call.Fn) else None,
            lg.dense_w.view(-1) if (lg is not None and call.Fn) else None,
            call.num_pos, call.lr_bias.data if call.lr_bias is not None else None,
            want_E=want_E, want_S=want_fm, want_fm=want_fm, want_lr=want_lr, B=B,
            lr_delta=call.lr_delta, num_widx=call.num_widx, D=call.D, n_slots=call.Ft,
            device=(g.table if g.table is not None else g.dense_w).device if g is not None else None)
        cnts = []
        for sq, ids in zip(seq, seq_ids):              # a9: pooled sequence slots, straight into their slot of E
            _, cnt = ops.pooled_gather_fwd(g.table, ids, sq["mode"], out=E[:, sq["pos"], :])
            cnts.append(cnt)
        ctx.call = call
        ctx.n_seq = len(seq)
        ctx.has_cnt = [c is not None for c in cnts]
        # optimizer.zero_grad() of the dense gradient buffer, off the backward's critical path: the buffer the backward
        # will scatter into is allocated NOW and zero-filled on a side stream while the forward kernels (and the rest
        # of the model's forward) run; the backward only waits for the fill's event
        ctx.gbuf = None
        if ASYNC_ZERO and E is not None and E.is_cuda and any(ctx.needs_input_grad[4:]):
            ctx.gbuf = _prezero(_grad_layout(call, call.with_main, call.with_lr), E.device)
        # the backward needs e only for the FM term d_fm * (S - e); it re-reads e from the table (rows mostly L2-resident)
        # rather than from the [B, F, D] stream the forward wrote (measured: 355 vs 351 M samples/s on the bench step), so
        # E is saved only when there is no table to re-read from (numeric-only selections)
        keep_E = want_fm and not (g is not None and g.emb and call.F)
        ctx.save_for_backward(rows, dense_x, E if keep_E else None, S, *seq_ids, *[c for c in cnts if c is not None])
        return (E, fm.view(-1, 1) if fm is not None else None, lr.view(-1, 1) if lr is not None else None)

    @staticmethod
    def backward(ctx, dE, d_fm, d_lr):
        call = ctx.call
        saved = ctx.saved_tensors
        rows, dense_x, E, S = saved[:4]
        seq_ids = saved[4:4 + ctx.n_seq]
        cnt_it = iter(saved[4 + ctx.n_seq:])
        cnts = [next(cnt_it) if h else None for h in ctx.has_cnt]
        g, lg = call.group, call.lr_group
        dev = (rows if rows is not None else (dense_x if dense_x is not None else seq_ids[0])).device
        B = (rows if rows is not None else (dense_x if dense_x is not None else seq_ids[0])).shape[0]
        D = call.D
        have_main = call.with_main and (dE is not None or d_fm is not None)
        have_lr = call.with_lr and d_lr is not None
        # one allocation, one zero-fill, for every dense gradient this launch produces
        pre, ctx.gbuf = ctx.gbuf, None                 # (a second backward through a retained graph gets a fresh buffer)
        if pre is not None:
            (n_t, n_w, n_t1, n_w1, n_b), offs, tot = _grad_layout(call, call.with_main, call.with_lr)
            buf, ev = pre
            torch.cuda.current_stream(dev).wait_event(ev)
            if not have_main:
                n_t = n_w = 0
            if not have_lr:
                n_t1 = n_w1 = n_b = 0
        else:
            (n_t, n_w, n_t1, n_w1, n_b), offs, tot = _grad_layout(call, have_main, have_lr)
            buf = torch.zeros(max(tot, 1), dtype=F32, device=dev)
        g_table = buf[offs[0]:offs[0] + n_t].view(-1, D) if n_t else None
        g_dw = buf[offs[1]:offs[1] + n_w].view(-1, D) if n_w else None
        g_t1 = buf[offs[2]:offs[2] + n_t1] if n_t1 else None
        g_dw1 = buf[offs[3]:offs[3] + n_w1] if n_w1 else None
        g_b = buf[offs[4]:offs[4] + 1] if n_b else None
        dE = dE.contiguous() if (dE is not None and have_main) else None
        if (have_main or have_lr) and (call.F or call.Fn):
            ops.embed_fm_bwd(
                g.table if (g is not None and g.emb) else None, rows, call.cat_pos, call.pad_row, dense_x,
                g.dense_w if (g is not None and g.lin) else None, call.num_pos,
                E, S, dE,
                d_fm.contiguous().view(-1) if (d_fm is not None and have_main and S is not None) else None,
                d_lr.contiguous().view(-1) if have_lr else None,
                g_table, g_t1, g_dw, g_dw1, g_b, D, g.R if (g is not None and g.emb) else (lg.R if lg is not None else 0),
                B=B, lr_delta=call.lr_delta, num_widx=call.num_widx, n_slots=call.Ft)
        if dE is not None and g_table is not None:
            for sq, ids, cnt in zip(call.seq or (), seq_ids, cnts):
                ops.pooled_gather_bwd(dE[:, sq["pos"], :], ids, cnt, sq["pad_row"], g_table, sq["mode"])
        # hand the per-parameter views back to autograd (they alias `buf`; AccumulateGrad adopts them)
        t_views = g_table.split(call.emb_sizes, 0) if g_table is not None else None
        t1_views = g_t1.split(call.lr_emb_sizes, 0) if g_t1 is not None else None
        grads = []
        for kind, idx in call.kinds:
            if kind == "emb":
                grads.append(t_views[idx] if t_views is not None else None)
            elif kind == "lin":
                grads.append(g_dw[idx].view(D, 1) if g_dw is not None else None)
            elif kind == "lr_emb":
                grads.append(t1_views[idx].view(-1, 1) if t1_views is not None else None)
            elif kind == "lr_lin":
                grads.append(g_dw1[idx].view(1, 1) if g_dw1 is not None else None)
            else:
                grads.append(g_b if g_b is not None else None)
        for i, p in enumerate(call.params):
            if not ctx.needs_input_grad[4 + i]:
                grads[i] = None
        return (None, None, None, None) + tuple(grads)


class _ShardedEmbedFn(torch.autograd.Function):
    """E, fm = fused(rows, dense_x; this rank's parameter shards) over the row-sharded table of a _ShardedStore group.
    forward : [zero of the local gradient shard -> cross-rank fence] -> ShardedEmbeddingFM.forward (remote rows over NVLink)
    backward: ShardedEmbeddingFM.backward (reductions land in the owners' gradient shards) -> fence -> views of the local
              gradient shard for the per-feature parameters; numeric-slot weights get this rank's partial sums."""

    @staticmethod
    def forward(ctx, call, rows, dense_x, *params):
        g = call.group
        sem, store = g.sem, g.store
        train = any(ctx.needs_input_grad[3:])
        dw = None
        if call.Fn:
            dw = g.dense_w if list(call.num_widx) == list(range(g.dense_w.shape[0])) else g.dense_w[list(call.num_widx)]
        if call.F:
            if train:
                # the zero-fill precedes the fence: no peer starts this step's backward (which reduces into this shard)
                # before every rank has passed the fence, i.e. before every shard is clean; and every rank's optimizer
                # step on its table shard has landed before any peer gathers from it
                sem.zero_grad()
                store.serial += 1
                sem.device_barrier()
            E, S, fm, _ = sem.forward(rows, call.cat_pos, dense_x, dw, None, call.num_pos, None, want_E=True, n_slots=call.Ft)
        else:
            E, S, fm, _ = ops.embed_fm_fwd(None, None, None, [], dense_x, dw, None, call.num_pos, None, want_lr=False,
                                           n_slots=call.Ft)
        ctx.call, ctx.dw, ctx.serial, ctx.done = call, dw, store.serial, False
        ctx.save_for_backward(rows, dense_x, E, S)
        return E, fm.view(-1, 1)

    @staticmethod
    def backward(ctx, dE, d_fm):
        call = ctx.call
        g = call.group
        sem, store, D, dw = g.sem, g.store, call.D, ctx.dw
        rows, dense_x, E, S = ctx.saved_tensors
        if ctx.done or (call.F and ctx.serial != store.serial):
            raise RbxError("sharded tables take ONE training forward / backward per embedding dictionary and step: the "
                           "gradient shard this backward reduces into was re-zeroed by a later forward (or used twice)")
        ctx.done = True
        dE = torch.zeros_like(E) if dE is None else dE.contiguous()
        d_fm = torch.zeros(E.shape[0], dtype=F32, device=E.device) if d_fm is None else d_fm.contiguous().view(-1)
        gw = torch.zeros_like(dw) if dw is not None else None
        if call.F:
            sem.backward(rows, call.cat_pos, call.pad_row, dense_x, dw, call.num_pos, E, S, dE, d_fm, None, gw, None, None,
                         n_slots=call.Ft)
            sem.device_barrier()               # every peer's reductions into this rank's shard have landed
        else:
            ops.embed_fm_bwd(None, None, [], None, dense_x, dw, call.num_pos, None, S, dE, d_fm, None, None, None, gw, None,
                             None, E.shape[-1], 0, B=E.shape[0], n_slots=call.Ft)
        widx = list(call.num_widx)
        grads = []
        for (kind, idx), p, need in zip(call.kinds, call.params, ctx.needs_input_grad[3:]):
            v = None
            if need and kind == "emb" and call.F:
                lo, hi = g.local[idx]
                v = sem.g_table[lo:hi, :D]
                if p.grad is not None and p.grad.data_ptr() == v.data_ptr() and p.grad.stride() == v.stride():
                    v = None                   # .grad still aliases the shard (zero_grad(set_to_none=False)): already in place
            elif need and kind == "lin" and gw is not None and idx in widx:
                v = gw[widx.index(idx), :D].reshape(D, 1)
            grads.append(v)
        return (None, None, None) + tuple(grads)


def fused_grad_buffers(params):
    """-> (shared, single): `shared` = one flat fp32 view per gradient STORAGE that several parameters' .grad alias (the
    per-feature .grad tensors a fused backward hands to autograd are views of one buffer per launch, see
    _FusedEmbedFn.backward) -- a replica's gradient exchange is then one collective over that buffer instead of one per
    feature; `single` = the gradients that own their storage."""
    groups = OrderedDict()
    for p in params:
        g = p.grad
        if g is None:
            continue
        groups.setdefault(g.untyped_storage().data_ptr(), []).append(g)
    shared, single = [], []
    for gs in groups.values():
        g0 = gs[0]
        if len(gs) == 1 and g0.is_contiguous() and g0.untyped_storage().nbytes() == g0.numel() * g0.element_size():
            single.append(g0)
        elif all(g.dtype == g0.dtype for g in gs):
            st = g0.untyped_storage()
            shared.append(torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, 0, (st.nbytes() // g0.element_size(),)))
        else:
            single.extend(gs)
    return shared, single


def sync_replica_gradients(params, group=None, average=False, reducer=None):
    """Data-parallel replicas of a table that fits every GPU (SURVEY 8e "replicas only"): dense all-reduce of the gradients
    over NCCL / NVLink, in place -- what DistributedDataParallel does for the reference's RecBole / rechub trainers
    (third_party/recbole/trainer/trainer.py:60-64) -- in as few collectives as the storage allows: one per fused gradient
    buffer, and ONE flattened bucket for all gradients that own their storage (copied in and out, like a DDP bucket).
    Returns the number of collectives issued."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    shared, single = fused_grad_buffers([p for p in params if not is_sharded(p)])   # sharded rows have one owner: nothing to reduce
    n = 0
    for b in shared:
        if reducer is not None and b.numel() <= reducer.numel:      # recbox_b200.replica.ReplicaReducer: in-switch all-reduce
            reducer.reduce_tensor(b, average)
        else:
            dist.all_reduce(b, group=group)
            if average:
                b.div_(dist.get_world_size(group))
        n += 1
    by_kind = OrderedDict()
    for g in single:
        by_kind.setdefault((g.dtype, g.device), []).append(g)
    for gs in by_kind.values():
        if len(gs) == 1:
            flat = gs[0]
            dist.all_reduce(flat, group=group)
            if average:
                flat.div_(dist.get_world_size(group))
        else:
            flat = torch.cat([g.reshape(-1) for g in gs])
            dist.all_reduce(flat, group=group)
            if average:
                flat.div_(dist.get_world_size(group))
            parts = flat.split([g.numel() for g in gs])
            if all(g.is_contiguous() for g in gs):
                torch._foreach_copy_(gs, [c.view_as(g) for c, g in zip(parts, gs)])      # one multi-tensor launch
            else:
                for g, c in zip(gs, parts):
                    g.copy_(c.reshape(g.shape))
        n += 1
    return n


# =================================================================================================
# a13: dense MLP tail on the tcgen05 GEMM (csrc/gemm.cu)
# =================================================================================================
class _MLPChainFn(torch.autograd.Function):
    """A run of Linear (+ ReLU) layers as ONE autograd node: forward = one fused GEMM per layer (bias + ReLU in the
    epilogue); backward = per layer dW = dZ^T H (split-K GEMM reading both operands as stored), db = column sums, and
    dH = (dZ W) * (H_prev > 0) with the ReLU mask of the previous layer fused into the epilogue, so no elementwise pass
    runs between the layers.  x: [M, K0]; args: relu flags, has-bias flags, then weight / bias tensors."""

    @staticmethod
    def forward(ctx, x, relus, *wb):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        hs, ws, bs = [x2], [], []
        it = iter(wb)
        for relu in relus:
            w, b = next(it), next(it)
            ws.append(w)
            bs.append(b)
            hs.append(ops.gemm(hs[-1], w.detach(), bias=b.detach() if b is not None else None, relu=relu))
        ctx.relus, ctx.n_out = tuple(relus), len(hs)
        ctx.shape = x.shape
        ctx.save_for_backward(*hs, *ws)
        ctx.has_bias = [b is not None for b in bs]
        return hs[-1].view(*x.shape[:-1], hs[-1].shape[-1])

    @staticmethod
    def backward(ctx, dy):
        L = len(ctx.relus)
        saved = ctx.saved_tensors
        hs, ws = saved[:L + 1], saved[L + 1:]
        dz = dy.reshape(-1, dy.shape[-1])
        if not dz.is_contiguous():
            dz = dz.contiguous()
        if ctx.relus[-1]:
            dz = dz * (hs[-1] > 0)                     # only when the chain ENDS in a ReLU (not the DeepFM / DNN shape)
        grads = [None] * (2 * L)
        dx = None
        for l in range(L - 1, -1, -1):
            need_w = ctx.needs_input_grad[2 + 2 * l]
            need_b = ctx.has_bias[l] and ctx.needs_input_grad[3 + 2 * l]
            if need_w:
                grads[2 * l] = ops.gemm(dz, hs[l], a_mn=True, b_mn=True)           # dW [N, K] = dZ^T H
            if need_b:
                grads[2 * l + 1] = ops.colsum(dz)
            if l > 0 or ctx.needs_input_grad[0]:
                prev_relu = l > 0 and ctx.relus[l - 1]
                dz = ops.gemm(dz, ws[l].detach(), b_mn=True, mask=hs[l] if prev_relu else None)   # dH = (dZ W) * (H > 0)
                if l == 0:
                    dx = dz.view(ctx.shape)
        return (dx, None) + tuple(grads)


def _mlp_forward(seq, x):
    """nn.Sequential of MLP_Block / MLP_Layer: maximal runs of Linear (+ nn.ReLU) go through _MLPChainFn, every other
    module (BatchNorm1d, Dropout, other activations) is called as it is."""
    mods = list(seq)
    i, n = 0, len(mods)
    while i < n:
        if type(mods[i]) is nn.Linear and x.is_cuda and x.dtype == F32:
            relus, wb = [], []
            while i < n and type(mods[i]) is nn.Linear:
                lin = mods[i]
                relu = i + 1 < n and type(mods[i + 1]) is nn.ReLU
                relus.append(relu)
                wb += [lin.weight, lin.bias]
                i += 2 if relu else 1
            x = _MLPChainFn.apply(x, tuple(relus), *wb)
        elif type(mods[i]) is nn.Linear:
            raise RbxError("MLP_Block: input must be an fp32 CUDA tensor (recbox_b200 has no CPU path)")
        else:
            x = mods[i](x)
            i += 1
    return x


def _activation(act, units=None):
    """fuxictr.pytorch.torch_utils.get_activation (ranking/pytorch/torch_utils.py:85-110) for the names the configs use."""
    if isinstance(act, str):
        low = act.lower()
        if low == "relu":
            return nn.ReLU()
        if low == "sigmoid":
            return nn.Sigmoid()
        if low == "tanh":
            return nn.Tanh()
        if low == "softmax":
            return nn.Softmax(dim=-1)
        if low == "prelu":
            return nn.PReLU(units, init=0.1)
        return getattr(nn, act)()
    return act


class MLP_Block(nn.Module):
    """Drop-in for ranking/pytorch/layers/blocks/mlp_block.py:23-61: same constructor, same `mlp` nn.Sequential (state_dict
    keys mlp.<i>.weight / bias, default nn.Linear init order), forward through the fused GEMM chain."""

    def __init__(self, input_dim, hidden_units=[], hidden_activations="ReLU", output_dim=None, output_activation=None,
                 dropout_rates=0.0, batch_norm=False, norm_before_activation=True, use_bias=True):
        super(MLP_Block, self).__init__()
        dense_layers = []
        if not isinstance(dropout_rates, list):
            dropout_rates = [dropout_rates] * len(hidden_units)
        if not isinstance(hidden_activations, list):
            hidden_activations = [hidden_activations] * len(hidden_units)
        hidden_activations = [_activation(a, u) for a, u in zip(hidden_activations, hidden_units)]
        hidden_units = [input_dim] + list(hidden_units)
        for idx in range(len(hidden_units) - 1):
            dense_layers.append(nn.Linear(hidden_units[idx], hidden_units[idx + 1], bias=use_bias))
            if norm_before_activation and batch_norm:
                dense_layers.append(nn.BatchNorm1d(hidden_units[idx + 1]))
            if hidden_activations[idx]:
                dense_layers.append(hidden_activations[idx])
            if not norm_before_activation and batch_norm:
                dense_layers.append(nn.BatchNorm1d(hidden_units[idx + 1]))
            if dropout_rates[idx] > 0:
                dense_layers.append(nn.Dropout(p=dropout_rates[idx]))
        if output_dim is not None:
            dense_layers.append(nn.Linear(hidden_units[-1], output_dim, bias=use_bias))
        if output_activation is not None:
            dense_layers.append(_activation(output_activation))
        self.mlp = nn.Sequential(*dense_layers)

    def forward(self, inputs):
        return _mlp_forward(self.mlp, inputs)


class MLP_Layer(nn.Module):
    """Drop-in for core/pytorch/layers/mlp.py:9-41 (the matching-side name and argument order of the same block)."""

    def __init__(self, input_dim, output_dim=None, hidden_units=[], hidden_activations="ReLU", final_activation=None,
                 dropout_rates=[], batch_norm=False, use_bias=True):
        super(MLP_Layer, self).__init__()
        dense_layers = []
        if not isinstance(dropout_rates, list):
            dropout_rates = [dropout_rates] * len(hidden_units)
        if not isinstance(hidden_activations, list):
            hidden_activations = [hidden_activations] * len(hidden_units)
        hidden_activations = [_activation(a) for a in hidden_activations]
        hidden_units = [input_dim] + list(hidden_units)
        for idx in range(len(hidden_units) - 1):
            dense_layers.append(nn.Linear(hidden_units[idx], hidden_units[idx + 1], bias=use_bias))
            if batch_norm:
                dense_layers.append(nn.BatchNorm1d(hidden_units[idx + 1]))
            if hidden_activations[idx]:
                dense_layers.append(hidden_activations[idx])
            if dropout_rates[idx] > 0:
                dense_layers.append(nn.Dropout(p=dropout_rates[idx]))
        if output_dim is not None:
            dense_layers.append(nn.Linear(hidden_units[-1], output_dim, bias=use_bias))
        if final_activation is not None:
            dense_layers.append(_activation(final_activation))
        self.mlp = nn.Sequential(*dense_layers)

    def forward(self, inputs):
        return _mlp_forward(self.mlp, inputs)


DETERMINISTIC = False    # True: the plain lookups' backward uses the sorted, atomics-free scatter (rbx_segment_sum_rows); the fused FM
                         # backward has the same mode at the ops level (ops.embed_fm_bwd_deterministic)


class _GatherFn(torch.autograd.Function):
    """Plain lookup [.., ] ids -> [.., D] (un-pooled sequence features, generic encoders)."""

    @staticmethod
    def forward(ctx, weight, ids, pad):
        ctx.save_for_backward(ids)
        ctx.meta = (weight.shape, pad)
        return ops.gather_rows(weight.data if isinstance(weight, nn.Parameter) else weight, ids)

    @staticmethod
    def backward(ctx, g):
        (ids,) = ctx.saved_tensors
        shape, pad = ctx.meta
        gt = torch.zeros(shape, dtype=F32, device=g.device)
        if DETERMINISTIC:          # reproducible summation order: sort + segmented sum, no atomics
            ops.scatter_add_rows_deterministic(g.contiguous(), ids, pad, gt)
        else:
            ops.scatter_add_rows(g.contiguous(), ids, pad, gt)
        return gt, None, None


class _PooledGatherFn(torch.autograd.Function):
    """ids [B,L] -> pooled [B,D] without the [B,L,D] intermediate (a9)."""

    @staticmethod
    def forward(ctx, weight, ids, pad, mode):
        out, cnt = ops.pooled_gather_fwd(weight, ids, mode)
        ctx.save_for_backward(ids, cnt)
        ctx.meta = (weight.shape, pad, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        ids, cnt = ctx.saved_tensors
        shape, pad, mode = ctx.meta
        gt = torch.zeros(shape, dtype=F32, device=g.device)
        ops.pooled_gather_bwd(g.contiguous(), ids, cnt, pad, gt, mode)
        return gt, None, None, None


class _InteractFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, mode):
        E = E.contiguous()
        ctx.save_for_backward(E)
        ctx.mode = mode
        return ops.interact_fwd(E, mode)

    @staticmethod
    def backward(ctx, g):
        (E,) = ctx.saved_tensors
        return ops.interact_bwd(E, g.contiguous(), ctx.mode), None


class _PowerSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, order):
        E = E.contiguous()
        ctx.save_for_backward(E)
        return ops.power_sums_fwd(E, order)

    @staticmethod
    def backward(ctx, g):
        (E,) = ctx.saved_tensors
        return ops.power_sums_bwd(E, g.contiguous()), None


class _RowDotFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, v):
        u, v = u.contiguous(), v.contiguous()
        ctx.save_for_backward(u, v)
        return ops.rowdot_fwd(u, v)

    @staticmethod
    def backward(ctx, dy):
        u, v = ctx.saved_tensors
        du, dv = ops.rowdot_bwd(u, v, dy.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return du, (dv.view_as(v) if dv is not None else None)


def two_tower_score(u, v):
    """y[b,k] = <u[b], v[b,k]> for u [B,D], v [B*K, D] or [B,K,D]  (match_model.py:71-75 layout;
    rechub dssm.py:48 is K = 1).  Differentiable; one kernel each way."""
    return _RowDotFn.apply(u, v)


# =================================================================================================
# input packing (a1)
# =================================================================================================
class PackedInputs(dict):
    """{feature: column} dict as RankingModel.get_inputs returns it, that also carries the device
    batch matrix it was sliced from, so the embedding layer converts all slots in ONE launch
    (rbx_split_batch_f64) instead of gathering 39 strided columns."""

    def __init__(self, feature_map, batch):
        super(PackedInputs, self).__init__()
        self.batch = batch
        self.feature_map = feature_map
        for feature, spec in feature_map.features.items():
            if spec["type"] == "meta":
                continue
            self[feature] = batch[:, feature_map.get_column_index(feature)]


def get_inputs(model, inputs, feature_source=None):
    """Drop-in for RankingModel.get_inputs (ranking_model.py:106-116): ONE host->device copy of the
    batch matrix, then per-feature views of it."""
    if isinstance(inputs, PackedBatch):          # loader.PackedDataLoader: 160 B / Criteo sample, no casts
        probe = next((t for t in (inputs.ids, inputs.ids16, inputs.dense, inputs.labels) if t is not None), None)
        pb = inputs if probe is None or probe.is_cuda else inputs.to(model.device, non_blocking=True)
        X = PackedColumns(pb)
        if feature_source:
            if type(feature_source) == str:
                feature_source = [feature_source]
            for feature, spec in model.feature_map.features.items():
                if spec["type"] != "meta" and spec["source"] not in feature_source:
                    X.pop(feature, None)
        return X
    batch = inputs.to(model.device, non_blocking=True)
    X = PackedInputs(model.feature_map, batch)
    if feature_source:
        if type(feature_source) == str:
            feature_source = [feature_source]
        for feature, spec in model.feature_map.features.items():
            if spec["type"] != "meta" and spec["source"] not in feature_source:
                X.pop(feature, None)
    return X


def get_labels(model, inputs):
    """Drop-in for RankingModel.get_labels (ranking_model.py:118-122): [B,1] float labels on the device."""
    if isinstance(inputs, PackedBatch):
        return inputs.labels.to(model.device, non_blocking=True).float().view(-1, 1)
    labels = model.feature_map.labels
    assert len(labels) == 1, "Please override get_labels(), add_loss(), evaluate() when using multiple labels!"
    return inputs[:, model.feature_map.get_column_index(labels[0])].to(model.device).float().view(-1, 1)


class _EmbDict(OrderedDict):
    """OrderedDict of per-feature embeddings that remembers the stacked tensor they are views of.

    The reference's dict2tensor always stacks the CURRENT dict values (feature_embedding.py:169-186), and model code is
    free to replace an entry first (DIN-style `feature_emb_dict[seq_field] = pooled`).  So any mutation after the producer
    has sealed the dict drops the cached stack, and so does an in-place write into it (checked through `_version`)."""
    _stacked = None
    _stacked_version = -1
    _sealed = False
    names = ()

    @property
    def stacked(self):
        E = self._stacked
        if E is not None and E._version != self._stacked_version:
            self._stacked = E = None          # somebody wrote into E (or into one of its views) in place
        return E

    @stacked.setter
    def stacked(self, E):
        self._stacked = E
        self._stacked_version = E._version if E is not None else -1
        self._sealed = E is not None

    def _touch(self):
        if self._sealed:
            self._stacked = None

    def __setitem__(self, key, value):
        self._touch()
        OrderedDict.__setitem__(self, key, value)

    def __delitem__(self, key):
        self._touch()
        OrderedDict.__delitem__(self, key)

    def pop(self, *a, **k):
        self._touch()
        return OrderedDict.pop(self, *a, **k)

    def popitem(self, *a, **k):
        self._touch()
        return OrderedDict.popitem(self, *a, **k)

    def update(self, *a, **k):
        self._touch()
        return OrderedDict.update(self, *a, **k)

    def setdefault(self, key, default=None):
        if key not in self:
            self._touch()
        return OrderedDict.setdefault(self, key, default)

    def clear(self):
        self._touch()
        OrderedDict.clear(self)

    def move_to_end(self, *a, **k):
        self._touch()
        return OrderedDict.move_to_end(self, *a, **k)


def _stash_of(feature_emb):
    """The fused launch's by-products riding on its E tensor -- ignored once E has been modified in place."""
    st = getattr(feature_emb, "_rbx_stash", None)
    if st is not None and st.version != feature_emb._version:
        return None
    return st


class _Stash(object):
    __slots__ = ("X", "fm", "lr", "lr_owner", "producer", "full", "version")


# =================================================================================================
# shared machinery of FeatureEmbeddingDict (ranking) and EmbeddingDictLayer (core / matching)
# =================================================================================================
class _FusedDictBase(nn.Module):
    _specs_attr = "features"          # FeatureMap attribute holding the OrderedDict of feature specs

    def _specs(self):
        return getattr(self._feature_map, self._specs_attr)

    def _encoders(self):
        raise NotImplementedError

    # -- storage ---------------------------------------------------------------------------------
    def _build_store(self):
        cfg = _shard_config()
        mods = list(self.embedding_layers.values())
        self._store = _FusedStore(mods) if cfg is None else _ShardedStore(mods, cfg)
        self._calls = {}
        self._calls_gen = _FUSE_GEN[0]
        self._lr_partner = None

    def _sync_store(self):
        """Re-fuse if a parameter was re-pointed (`emb.weight = nn.Parameter(...)`, pretrained weights loaded after
        construction), and drop every cached launch plan made before the last fusion of ANY store: plans hold the
        Parameter objects autograd routes gradients to, and the partner LR module's as well."""
        self._store.ensure()
        if self._calls_gen != _FUSE_GEN[0]:
            self._calls = {}
            self._calls_gen = _FUSE_GEN[0]

    def __getstate__(self):
        # the store keys its offsets by id(module) and aliases one allocation: rebuilt on unpickle
        if getattr(self._store, "sharded", False):
            raise RbxError("a module with row-sharded tables is not picklable: save gather_state_dict(module)")
        state = dict(self.__dict__)
        for k in ("_store", "_calls", "_calls_gen", "_lr_partner"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super(_FusedDictBase, self).__setstate__(state)
        self._build_store()

    def _apply(self, fn, *args, **kwargs):
        out = super(_FusedDictBase, self)._apply(fn, *args, **kwargs)
        if getattr(self, "_store", None) is not None:
            self._store.fuse()          # Module.to()/cuda()/float() re-pointed every parameter
            self._calls = {}
            self._calls_gen = _FUSE_GEN[0]
        return out

    def __deepcopy__(self, memo):
        import copy
        if getattr(self._store, "sharded", False):
            raise RbxError("a module with row-sharded tables cannot be deep-copied: rebuild it from gather_state_dict(module)")
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_store", "_calls", "_calls_gen", "_lr_partner"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._build_store()
        return new

    def row_offsets(self, cat_names):
        """First fused-table row of each named categorical feature (loader.PackedDataLoader.bind)."""
        self._store.ensure()
        offs, dims = [], set()
        for n in cat_names:
            m = self.embedding_layers[n]
            if not isinstance(m, nn.Embedding):
                raise RbxError("feature %r has no embedding table" % n)
            dims.add(m.embedding_dim)
            offs.append(self._store.groups[m.embedding_dim].emb_off[id(m)])
        if len(dims) > 1:
            raise RbxError("row_offsets: features of different embedding dims live in different fused tables")
        return offs

    # -- call plans ------------------------------------------------------------------------------
    def _select(self, feature_source, feature_type):
        raise NotImplementedError

    def _plan(self, key, names):
        """Classify the selected features: which go through the fused launch, which need the
        pooled / plain gather kernels.  Cached per selection."""
        plan = self._calls.get(key)
        if plan is not None:
            return plan
        specs, enc = self._specs(), self._encoders()
        dims = set()
        fused, seq_pool, generic = [], [], []
        for name in names:
            mod = self.embedding_layers[name]
            spec = specs[name]
            d = mod.embedding_dim if isinstance(mod, nn.Embedding) else mod.out_features
            dims.add(d)
            if spec["type"] in ("numeric", "categorical") and name not in enc:
                fused.append(name)
            elif spec["type"] == "sequence" and name in enc and _pool_mode(enc[name]) is not None:
                seq_pool.append(name)
            else:
                generic.append(name)
        plan = {"names": list(names), "fused": fused, "seq_pool": seq_pool, "generic": generic,
                "uniform": len(dims) == 1 and not generic, "call": None}
        if plan["uniform"] and (fused or seq_pool):
            D = next(iter(dims))
            plan["D"] = D
            plan["call"] = self._make_call(names, fused, D, None, seq_pool)
        self._calls[key] = plan
        return plan

    def _make_call(self, names, fused, D, lr_module, seq_pool=()):
        """Static launch description for the `fused` features placed at their positions among
        `names`; with lr_module (a LogisticRegression over the same features) the first-order term
        is computed in the same launch."""
        specs = self._specs()
        g = self._store.groups[D]
        c = _Call()
        c.D, c.group = D, g
        pos = {n: i for i, n in enumerate(names)}
        cats = [n for n in fused if specs[n]["type"] == "categorical"]
        nums = [n for n in fused if specs[n]["type"] == "numeric"]
        c.F, c.Fn, c.Ft = len(cats), len(nums), len(names)
        c.cat_pos = [pos[n] for n in cats]
        c.num_pos = [pos[n] for n in nums]
        c.num_widx = [g.lin_idx[id(self.embedding_layers[n])] for n in nums]
        offs = [g.emb_off[id(self.embedding_layers[n])] for n in cats]
        c.pad_row = []
        for n, o in zip(cats, offs):
            pi = self.embedding_layers[n].padding_idx
            c.pad_row.append(-1 if pi is None else o + pi)
        c.with_main, c.with_lr = True, False
        c.lr_group, c.lr_bias, c.lr_delta = None, None, None
        enc = self._encoders()
        c.seq = []
        for n in seq_pool:
            m = self.embedding_layers[n]
            o = g.emb_off[id(m)]
            c.seq.append({"name": n, "pos": pos[n], "off": o, "mode": _pool_mode(enc[n]), "vocab": m.num_embeddings,
                          "pad_row": -1 if m.padding_idx is None else o + m.padding_idx})
        params, kinds = [], []
        for i, m in enumerate(g.emb):
            params.append(m.weight)
            kinds.append(("emb", i))
        for i, m in enumerate(g.lin):
            params.append(m.weight)
            kinds.append(("lin", i))
        c.emb_sizes = [m.num_embeddings for m in g.emb]
        c.lr_emb_sizes = []
        if lr_module is not None:
            ld = lr_module.embedding_layer.embedding_layer
            lg = ld._store.groups[1]
            c.lr_group, c.lr_bias, c.with_lr = lg, lr_module.bias, True
            c.lr_delta = [lg.emb_off[id(ld.embedding_layers[n])] - o for n, o in zip(cats, offs)]
            for i, m in enumerate(lg.emb):
                params.append(m.weight)
                kinds.append(("lr_emb", i))
            for i, m in enumerate(lg.lin):
                params.append(m.weight)
                kinds.append(("lr_lin", i))
            if lr_module.bias is not None:
                params.append(lr_module.bias)
                kinds.append(("bias", 0))
            c.lr_emb_sizes = [m.num_embeddings for m in lg.emb]
        c.params, c.kinds = params, kinds
        c_names = {"cats": cats, "nums": nums, "offs": offs,
                   "vocab": [self.embedding_layers[n].num_embeddings for n in cats]}
        return c, c_names

    # -- packing ---------------------------------------------------------------------------------
    def _bad_counter(self, dev):
        """int32 device counter of ids that fell outside their feature's vocabulary (where nn.Embedding raises
        IndexError); such ids read a zero row and get no gradient.  `out_of_range_ids()` reads it (synchronises)."""
        c = self.__dict__.get("_n_bad")
        if c is None or c.device != dev:
            c = torch.zeros(1, dtype=I32, device=dev)
            self.__dict__["_n_bad"] = c
        return c

    def out_of_range_ids(self):
        c = self.__dict__.get("_n_bad")
        return int(c.item()) if c is not None else 0

    def _pack(self, inputs, cats, nums, offs, vocab=None):
        fm = self._feature_map
        if isinstance(inputs, PackedInputs) and inputs.feature_map is fm and inputs.batch.is_cuda \
                and inputs.batch.dtype == torch.float64 and inputs.batch.stride(1) == 1:
            n_cols = inputs.batch.shape[1]
            kind, slot = [0] * n_cols, [0] * n_cols
            for i, n in enumerate(cats):
                ci = fm.get_column_index(n)
                kind[ci], slot[ci] = 1, i
            for i, n in enumerate(nums):
                ci = fm.get_column_index(n)
                kind[ci], slot[ci] = 2, i
            rows, dense_x, _ = ops.split_batch(inputs.batch, kind, slot, offs, len(cats), len(nums), want_label=False,
                                               field_rows=vocab, n_bad=self._bad_counter(inputs.batch.device) if vocab else None)
            return rows, dense_x
        rows = dense_x = None
        if isinstance(inputs, PackedColumns):
            pb = inputs.packed
            if pb.offsets is not None:                  # the loader pre-added ITS layer's row offsets
                pre = dict(zip(pb.cat_names, pb.offsets))
                offs = [o - pre.get(n, 0) for n, o in zip(cats, offs)]
            if cats and pb.ids is not None and pb.ids.is_cuda and list(cats) == list(pb.cat_names) \
                    and pb.offsets is not None and not any(offs):
                rows = pb.ids                           # the block IS the kernels' `rows` argument
            elif cats and pb.ids16 is not None and pb.ids16.is_cuda and list(cats) == list(pb.cat_names):
                rows = ops.unpack_ids_u16(pb.ids16, offs)   # compact block: uint16 local ids + row offsets, one launch
            if nums and pb.dense is not None and pb.dense.is_cuda and list(nums) == list(pb.num_names):
                dense_x = pb.dense
        if cats and rows is None:
            cols = [self._col(inputs[n]) for n in cats]
            rows = ops.pack_columns(cols, add=offs, as_rows=True, vocab=vocab,
                                    n_bad=self._bad_counter(cols[0].device) if vocab else None)
        if nums and dense_x is None:
            dense_x = ops.pack_columns([self._col(inputs[n]) for n in nums], as_rows=False)
        return rows, dense_x

    @staticmethod
    def _col(t):
        if not t.is_cuda:
            raise RbxError("recbox_b200 layers need CUDA inputs (no CPU path); got a %s tensor" % t.device)
        return t.reshape(-1) if t.dim() != 1 else t

    # -- the forward all variants share ------------------------------------------------------------
    def _embed(self, inputs, key, names):
        """-> _EmbDict of per-feature embeddings for `names` (in order)."""
        lr_partner = self._lr_partner() if self._lr_partner is not None else None
        if lr_partner is not None:      # its parameters ride in this launch: a re-fusion there invalidates our plans too
            lr_partner.embedding_layer.embedding_layer._sync_store()
        self._sync_store()
        if getattr(self._store, "sharded", False):
            return self._embed_sharded(inputs, key, names)
        for g in self._store.groups.values():
            ref = g.table if g.table is not None else g.dense_w
            if not ref.is_cuda:
                raise RbxError("recbox_b200 layers have no CPU path: move the model to a CUDA device")
        plan = self._plan(key, names)
        specs, enc = self._specs(), self._encoders()
        out = _EmbDict()
        out.names = tuple(names)
        if plan["call"] is not None:
            call, cn = plan["call"]
            partner = self._lr_partner() if self._lr_partner is not None else None
            use_lr = None
            if partner is not None and not plan["seq_pool"] and key == self._full_key():
                pc = plan.get("lr_call")
                if pc is None or pc[2] is not partner:
                    pc = self._make_call(names, plan["fused"], plan["D"], partner) + (partner,)
                    plan["lr_call"] = pc
                call, cn, use_lr = pc[0], pc[1], partner
            rows, dense_x = self._pack(inputs, cn["cats"], cn["nums"], cn["offs"], cn.get("vocab"))
            seq_ids = tuple(self._seq_rows(inputs[sq["name"]], sq["off"], sq["vocab"],
                                          self._bad_counter(inputs[sq["name"]].device)) for sq in call.seq)
            E, fm, lr = _FusedEmbedFn.apply(call, rows, dense_x, seq_ids, *call.params)
            if not plan["seq_pool"]:
                st = _Stash()
                st.X, st.fm, st.lr, st.lr_owner, st.producer = inputs, fm, lr, use_lr, weakref.ref(self)
                st.full = key == self._full_key()
                st.version = E._version
                E._rbx_stash = st
            for i, name in enumerate(names):
                out[name] = E[:, i, :]
            out.stacked = E                  # seals the dict: a later mutation drops the cached stack
            return out
        # generic path: per-feature kernels, torch.stack later (mixed dims, custom encoders, ...)
        for name in names:
            mod, spec = self.embedding_layers[name], specs[name]
            if spec["type"] == "numeric":
                x = self._col(inputs[name]).float().view(-1, 1)
                e = x * mod.weight.view(1, -1)
            elif spec["type"] in ("categorical", "sequence"):
                ids = inputs[name]
                if not ids.is_cuda:
                    raise RbxError("recbox_b200 layers need CUDA inputs (no CPU path)")
                ids = ids.to(I32).contiguous()
                pad = mod.padding_idx
                mode = _pool_mode(enc[name]) if name in enc else None
                if spec["type"] == "sequence" and mode is not None:
                    e = _PooledGatherFn.apply(mod.weight, ids, pad, mode)
                    out[name] = e
                    continue
                e = _GatherFn.apply(mod.weight, ids, pad)
            else:
                raise NotImplementedError
            if name in enc:
                e = enc[name](e)
            out[name] = e
        return out

    def _embed_sharded(self, inputs, key, names):
        """_embed over row-sharded tables (_ShardedStore): the same [B, F, D] tensor and FM by-product, produced by the
        sharded fused kernels; the first-order term is NOT paired into the launch (LogisticRegression is its own
        dictionary, sharded the same way, and runs its own lookup)."""
        plan = self._plan(key, names)
        if plan["call"] is None or plan["seq_pool"]:
            raise RbxError("sharded tables cover plain categorical / numeric slots of ONE embedding dim (no pooled sequence "
                           "slots, custom encoders or mixed dims): selection %r is outside that" % (list(names),))
        call, cn = plan["call"]
        rows, dense_x = self._pack(inputs, cn["cats"], cn["nums"], cn["offs"], cn.get("vocab"))
        E, fm = _ShardedEmbedFn.apply(call, rows, dense_x, *call.params)
        if E.shape[-1] != call.D:                  # D < 4 tables (the D = 1 LogisticRegression trick) live in 4-float rows
            E, fm = E[..., :call.D], None
        out = _EmbDict()
        out.names = tuple(names)
        st = _Stash()
        st.X, st.fm, st.lr, st.lr_owner, st.producer = inputs, fm, None, None, weakref.ref(self)
        st.full = key == self._full_key()
        st.version = E._version
        E._rbx_stash = st
        for i, name in enumerate(names):
            out[name] = E[:, i, :]
        out.stacked = E
        return out

    @staticmethod
    def _seq_rows(ids, off, vocab=0, n_bad=None):
        """[B, L] ids (float64 / int64 / int32 as the loader delivers them) -> int32 global rows."""
        if not ids.is_cuda:
            raise RbxError("recbox_b200 layers need CUDA inputs (no CPU path)")
        if ids.dim() != 2:
            ids = ids.reshape(ids.shape[0], -1)
        L = ids.shape[1]
        if L == 0 or L > 192:
            out = ids.to(I32)
            if vocab:
                out = torch.where((out >= 0) & (out < vocab), out + off, torch.full_like(out, -1))
                return out.contiguous()
            return (out + off if off else out).contiguous()
        return ops.pack_columns([ids[:, l] for l in range(L)], add=[off] * L, as_rows=True,
                                vocab=[vocab] * L if vocab else None, n_bad=n_bad)

    def _full_key(self):
        return ((), ())


# =================================================================================================
# ranking side
# =================================================================================================
class FeatureEmbeddingDict(_FusedDictBase):
    """ranking/pytorch/layers/embeddings/feature_embedding.py:52-214 (same ctor, attributes,
    parameter names, init order)."""
    _specs_attr = "features"

    def __init__(self, feature_map, embedding_dim, embedding_initializer="partial(nn.init.normal_, std=1e-4)",
                 required_feature_columns=None, not_required_feature_columns=None, use_pretrain=True,
                 use_sharing=True):
        super(FeatureEmbeddingDict, self).__init__()
        self._feature_map = feature_map
        self.required_feature_columns = required_feature_columns
        self.not_required_feature_columns = not_required_feature_columns
        self.use_pretrain = use_pretrain
        self.embedding_initializer = embedding_initializer
        self.embedding_layers = nn.ModuleDict()
        self.feature_encoders = nn.ModuleDict()
        for feature, feature_spec in self._feature_map.features.items():
            if self.is_required(feature):
                if not (use_pretrain and use_sharing) and embedding_dim == 1:
                    feat_emb_dim = 1  # in case for LR
                    if feature_spec["type"] == "sequence":
                        self.feature_encoders[feature] = MaskedSumPooling()
                else:
                    feat_emb_dim = feature_spec.get("embedding_dim", embedding_dim)
                    if feature_spec.get("feature_encoder", None):
                        self.feature_encoders[feature] = self.get_feature_encoder(feature_spec["feature_encoder"])
                if use_sharing and feature_spec.get("share_embedding") in self.embedding_layers:
                    self.embedding_layers[feature] = self.embedding_layers[feature_spec["share_embedding"]]
                    continue
                if feature_spec["type"] == "numeric":
                    self.embedding_layers[feature] = nn.Linear(1, feat_emb_dim, bias=False)
                elif feature_spec["type"] in ("categorical", "sequence"):
                    padding_idx = feature_spec.get("padding_idx", None)
                    embedding_matrix = nn.Embedding(feature_spec["vocab_size"], feat_emb_dim, padding_idx=padding_idx)
                    if use_pretrain and "pretrained_emb" in feature_spec:
                        embedding_matrix = self.load_pretrained_embedding(
                            embedding_matrix, feature_map, feature, freeze=feature_spec["freeze_emb"],
                            padding_idx=padding_idx)
                    self.embedding_layers[feature] = embedding_matrix
        self.reset_parameters()
        self._build_store()

    def _encoders(self):
        return self.feature_encoders

    def get_feature_encoder(self, encoder):
        import sys
        layers = sys.modules[__name__]  # noqa: F841  ("layers.MaskedAveragePooling()" strings)
        try:
            if type(encoder) == list:
                return nn.Sequential(*[eval(enc) for enc in encoder])
            return eval(encoder)
        except Exception:
            raise ValueError("feature_encoder={} is not supported.".format(encoder))

    def reset_parameters(self):
        init = self.embedding_initializer
        if isinstance(init, str):
            try:
                init = eval(init)
            except Exception:
                raise ValueError("initializer={} is not supported.".format(init))
        self.embedding_initializer = init
        for k, v in self.embedding_layers.items():
            if self.use_pretrain and "pretrained_emb" in self._feature_map.features[k]:
                continue
            if "share_embedding" in self._feature_map.features[k] and v.weight.requires_grad == False:  # noqa: E712
                continue
            if type(v) == nn.Embedding:
                if v.padding_idx is not None:  # using 0 index as padding_idx
                    self.embedding_initializer(v.weight[1:, :])
                else:
                    self.embedding_initializer(v.weight)

    def is_required(self, feature):
        feature_spec = self._feature_map.features[feature]
        if feature_spec["type"] == "meta":
            return False
        elif self.required_feature_columns and (feature not in self.required_feature_columns):
            return False
        elif self.not_required_feature_columns and (feature in self.not_required_feature_columns):
            return False
        return True

    def get_pretrained_embedding(self, pretrained_path, feature_name):
        import h5py
        with h5py.File(pretrained_path, 'r') as hf:
            embeddings = hf[feature_name][:]
        return embeddings

    def load_pretrained_embedding(self, embedding_matrix, feature_map, feature_name, freeze=False, padding_idx=None):
        import os
        pretrained_path = os.path.join(feature_map.data_dir, feature_map.features[feature_name]["pretrained_emb"])
        embeddings = self.get_pretrained_embedding(pretrained_path, feature_name)
        if padding_idx is not None:
            embeddings[padding_idx] = np.zeros(embeddings.shape[-1])
        assert embeddings.shape[-1] == embedding_matrix.embedding_dim, \
            "{}\'s embedding_dim is not correctly set to match its pretrained_emb shape".format(feature_name)
        embedding_matrix.weight = torch.nn.Parameter(torch.from_numpy(embeddings).float())
        if freeze:
            embedding_matrix.weight.requires_grad = False
        return embedding_matrix

    def _select(self, feature_source, feature_type):
        if type(feature_source) != list:
            feature_source = [feature_source]
        if type(feature_type) != list:
            feature_type = [feature_type]
        key = (tuple(feature_source), tuple(feature_type))
        names = []
        for feature, spec in self._feature_map.features.items():
            if feature_source and spec["source"] not in feature_source:
                continue
            if feature_type and spec["type"] not in feature_type:
                continue
            if feature in self.embedding_layers:
                if spec["type"] not in ("numeric", "categorical", "sequence"):
                    raise NotImplementedError
                names.append(feature)
        return key, names

    def dict2tensor(self, embedding_dict, feature_source=[], feature_type=[], dynamic_emb_dim=False):
        if type(feature_source) != list:
            feature_source = [feature_source]
        if type(feature_type) != list:
            feature_type = [feature_type]
        names = []
        for feature, spec in self._feature_map.features.items():
            if feature_source and spec["source"] not in feature_source:
                continue
            if feature_type and spec["type"] not in feature_type:
                continue
            if feature in embedding_dict:
                names.append(feature)
        if (not dynamic_emb_dim and isinstance(embedding_dict, _EmbDict) and embedding_dict.stacked is not None
                and tuple(names) == embedding_dict.names):
            return embedding_dict.stacked            # already [B, F, D] in place: no stack copy
        vals = [embedding_dict[n] for n in names]
        return torch.cat(vals, dim=-1) if dynamic_emb_dim else torch.stack(vals, dim=1)

    def forward(self, inputs, feature_source=[], feature_type=[]):
        key, names = self._select(feature_source, feature_type)
        return self._embed(inputs, key, names)


class FeatureEmbedding(nn.Module):
    """ranking/pytorch/layers/embeddings/feature_embedding.py:28-49."""

    def __init__(self, feature_map, embedding_dim, embedding_initializer="partial(nn.init.normal_, std=1e-4)",
                 required_feature_columns=None, not_required_feature_columns=None, use_pretrain=True,
                 use_sharing=True):
        super(FeatureEmbedding, self).__init__()
        self.embedding_layer = FeatureEmbeddingDict(feature_map, embedding_dim,
                                                    embedding_initializer=embedding_initializer,
                                                    required_feature_columns=required_feature_columns,
                                                    not_required_feature_columns=not_required_feature_columns,
                                                    use_pretrain=use_pretrain, use_sharing=use_sharing)

    def forward(self, X, feature_source=[], feature_type=[], dynamic_emb_dim=False):
        feature_emb_dict = self.embedding_layer(X, feature_source=feature_source, feature_type=feature_type)
        return self.embedding_layer.dict2tensor(feature_emb_dict, dynamic_emb_dim=dynamic_emb_dim)


class InnerProductInteraction(nn.Module):
    """ranking/pytorch/layers/interactions/inner_product.py:22-56.
    output: product_sum (bs x 1), bi_interaction (bs x dim), inner_product (bs x f(f-1)/2),
            elementwise_product (bs x f(f-1)/2 x dim)"""

    def __init__(self, num_fields, output="product_sum"):
        super(InnerProductInteraction, self).__init__()
        self._output_type = output
        if output not in ["product_sum", "bi_interaction", "inner_product", "elementwise_product"]:
            raise ValueError("InnerProductInteraction output={} is not supported.".format(output))
        if output == "inner_product":       # kept so state_dict keys equal the reference's
            self.interaction_units = int(num_fields * (num_fields - 1) / 2)
            self.triu_mask = nn.Parameter(torch.triu(torch.ones(num_fields, num_fields), 1).bool(),
                                          requires_grad=False)
        elif output == "elementwise_product":
            self.triu_index = nn.Parameter(torch.triu_indices(num_fields, num_fields, offset=1), requires_grad=False)

    def forward(self, feature_emb):
        if self._output_type == "product_sum":
            st = _stash_of(feature_emb)
            if st is not None and st.fm is not None:
                return st.fm                 # computed by the launch that produced feature_emb
        if not feature_emb.is_cuda:
            raise RbxError("recbox_b200 layers have no CPU path")
        return _InteractFn.apply(feature_emb, ops.MODES[self._output_type])


class InteractionMachine(nn.Module):
    """ranking/pytorch/layers/interactions/interaction_machine.py:20-70 (same ctor, `bn` / `fc` parameter names).  The
    power sums p_k = sum_f e_f^k come from ONE pass over [B,F,D] (rbx_power_sums_fwd) instead of `order` passes with
    `order` [B,F,D] temporaries; the Newton-identity polynomials of :29-42 act on the [B,D] sums."""

    def __init__(self, embedding_dim, order=2, batch_norm=False):
        super(InteractionMachine, self).__init__()
        assert order < 6, "order={} is not supported.".format(order)
        self.order = order
        self.bn = nn.BatchNorm1d(embedding_dim * order) if batch_norm else None
        self.fc = nn.Linear(order * embedding_dim, 1)

    def second_order(self, p1, p2):
        return (p1.pow(2) - p2) / 2

    def third_order(self, p1, p2, p3):
        return (p1.pow(3) - 3 * p1 * p2 + 2 * p3) / 6

    def fourth_order(self, p1, p2, p3, p4):
        return (p1.pow(4) - 6 * p1.pow(2) * p2 + 3 * p2.pow(2)
                + 8 * p1 * p3 - 6 * p4) / 24

    def fifth_order(self, p1, p2, p3, p4, p5):
        return (p1.pow(5) - 10 * p1.pow(3) * p2 + 20 * p1.pow(2) * p3 - 30 * p1 * p4
                - 20 * p2 * p3 + 15 * p1 * p2.pow(2) + 24 * p5) / 120

    def forward(self, X):
        if not X.is_cuda:
            raise RbxError("recbox_b200 layers have no CPU path")
        if self.order < 1:
            out = X.new_zeros((X.shape[0], 0))
        else:
            P = _PowerSumFn.apply(X, self.order)
            p = [P[:, k, :] for k in range(self.order)]
            out = [p[0]]
            if self.order >= 2:
                out.append(self.second_order(p[0], p[1]))
            if self.order >= 3:
                out.append(self.third_order(p[0], p[1], p[2]))
            if self.order >= 4:
                out.append(self.fourth_order(p[0], p[1], p[2], p[3]))
            if self.order == 5:
                out.append(self.fifth_order(p[0], p[1], p[2], p[3], p[4]))
            out = torch.cat(out, dim=-1)
        if self.bn is not None:
            out = self.bn(out)
        return self.fc(out)


class LogisticRegression(nn.Module):
    """ranking/pytorch/layers/blocks/logistic_regression.py:23-35."""

    def __init__(self, feature_map, use_bias=True):
        super(LogisticRegression, self).__init__()
        self.bias = nn.Parameter(torch.zeros(1), requires_grad=True) if use_bias else None
        # A trick for quick one-hot encoding in LR
        self.embedding_layer = FeatureEmbedding(feature_map, 1, use_pretrain=False, use_sharing=False)

    def forward(self, X):
        d = self.embedding_layer.embedding_layer
        d._sync_store()
        key, names = d._select([], [])
        plan = d._plan(key, names)
        if plan["call"] is None or plan["seq_pool"] or getattr(d._store, "sharded", False):
            embed_weights = self.embedding_layer(X)        # generic route (sequence slots, sharded tables): sum over fields
            output = embed_weights.sum(dim=1)
            if self.bias is not None:
                output = output + self.bias
            return output
        d._store.ensure()
        c = plan.get("lr_only")
        if c is None:
            call, cn = plan["call"]
            lo = _Call()
            for k in _Call.__slots__:
                if hasattr(call, k):
                    setattr(lo, k, getattr(call, k))
            g = d._store.groups[1]
            lo.group, lo.lr_group, lo.with_main, lo.with_lr = None, g, False, True
            lo.lr_bias, lo.lr_delta = self.bias, None
            lo.params = [m.weight for m in g.emb] + [m.weight for m in g.lin] + ([self.bias] if self.bias is not None else [])
            lo.kinds = [("lr_emb", i) for i in range(len(g.emb))] + [("lr_lin", i) for i in range(len(g.lin))] + \
                       ([("bias", 0)] if self.bias is not None else [])
            lo.lr_emb_sizes, lo.emb_sizes, lo.seq = [m.num_embeddings for m in g.emb], [], []
            c = (lo, cn)
            plan["lr_only"] = c
        call, cn = c
        rows, dense_x = d._pack(X, cn["cats"], cn["nums"], cn["offs"], cn.get("vocab"))
        _, _, lr = _FusedEmbedFn.apply(call, rows, dense_x, (), *call.params)
        return lr


class FactorizationMachine(nn.Module):
    """ranking/pytorch/layers/blocks/factorization_machine.py:24-34."""

    def __init__(self, feature_map):
        super(FactorizationMachine, self).__init__()
        self.fm_layer = InnerProductInteraction(feature_map.num_fields, output="product_sum")
        self.lr_layer = LogisticRegression(feature_map, use_bias=True)

    def forward(self, X, feature_emb):
        st = _stash_of(feature_emb)
        if st is not None and st.X is X and st.lr is not None and st.lr_owner is self.lr_layer:
            return st.fm + st.lr             # both came out of the launch that produced feature_emb
        lr_out = self.lr_layer(X)
        fm_out = self.fm_layer(feature_emb)
        if st is not None and st.X is X and st.full:
            _try_pair(st.producer(), self.lr_layer)
        return fm_out + lr_out


def _try_pair(producer, lr_module):
    """Let `producer` (the FeatureEmbeddingDict that made feature_emb) compute this FM block's
    first-order term in its own launch from now on.  Only when both see the same features as
    plain categorical / numeric slots."""
    if producer is None or not FUSE_FM or getattr(producer._store, "sharded", False):
        return
    ld = lr_module.embedding_layer.embedding_layer
    if ld._feature_map is not producer._feature_map or getattr(ld._store, "sharded", False):
        return
    kp, np_ = producer._select([], [])
    kl, nl = ld._select([], [])
    if np_ != nl:
        return
    pp, pl = producer._plan(kp, np_), ld._plan(kl, nl)
    if pp["call"] is None or pl["call"] is None or pp["seq_pool"] or pl["seq_pool"]:
        return
    if pp["call"][0].num_widx != pl["call"][0].num_widx:
        return
    producer._lr_partner = weakref.ref(lr_module)


FUSE_FM = True   # set False to keep FactorizationMachine's first-order term in its own launch


# =================================================================================================
# core / matching side
# =================================================================================================
class EmbeddingDictLayer(_FusedDictBase):
    """core/pytorch/layers/embedding.py:30-138 (same ctor, attributes and parameter names)."""
    _specs_attr = "feature_specs"

    def __init__(self, feature_map, embedding_dim, disable_sharing_pretrain=False, required_feature_columns=None,
                 not_required_feature_columns=None):
        super(EmbeddingDictLayer, self).__init__()
        import sys
        layers = _CoreNamespace(sys.modules[__name__])  # noqa: F841  ("layers.MaskedAveragePooling()" strings)
        self._feature_map = feature_map
        self.required_feature_columns = required_feature_columns
        self.not_required_feature_columns = not_required_feature_columns
        self.embedding_layers = nn.ModuleDict()
        self.embedding_callbacks = nn.ModuleDict()
        for feature, feature_spec in self._feature_map.feature_specs.items():
            if self.is_required(feature):
                if disable_sharing_pretrain:  # in case for LR
                    assert embedding_dim == 1
                    feat_emb_dim = embedding_dim
                else:
                    feat_emb_dim = feature_spec.get("embedding_dim", embedding_dim)
                if (not disable_sharing_pretrain) and "embedding_callback" in feature_spec:
                    self.embedding_callbacks[feature] = eval(feature_spec["embedding_callback"])
                if (not disable_sharing_pretrain) and "share_embedding" in feature_spec:
                    self.embedding_layers[feature] = self.embedding_layers[feature_spec["share_embedding"]]
                    continue
                if feature_spec["type"] == "numeric":
                    self.embedding_layers[feature] = nn.Linear(1, feat_emb_dim, bias=False)
                elif feature_spec["type"] in ("categorical", "sequence"):
                    padding_idx = feature_spec.get("padding_idx", None)
                    embedding_matrix = nn.Embedding(feature_spec["vocab_size"], feat_emb_dim, padding_idx=padding_idx)
                    if (not disable_sharing_pretrain) and "pretrained_emb" in feature_spec:
                        embedding_matrix = self.load_pretrained_embedding(
                            embedding_matrix, feature_map, feature, freeze=feature_spec["freeze_emb"],
                            padding_idx=padding_idx)
                    self.embedding_layers[feature] = embedding_matrix
        self._build_store()

    def _encoders(self):
        return self.embedding_callbacks

    def is_required(self, feature):
        if self.required_feature_columns and (feature not in self.required_feature_columns):
            return False
        if self.not_required_feature_columns and (feature in self.not_required_feature_columns):
            return False
        return True

    def get_pretrained_embedding(self, pretrained_path, feature_name):
        import h5py
        with h5py.File(pretrained_path, 'r') as hf:
            embeddings = hf[feature_name][:]
        return embeddings

    def load_pretrained_embedding(self, embedding_matrix, feature_map, feature_name, freeze=False, padding_idx=None):
        import os
        pretrained_path = os.path.join(feature_map.data_dir, feature_map.feature_specs[feature_name]["pretrained_emb"])
        embeddings = self.get_pretrained_embedding(pretrained_path, feature_name)
        if padding_idx is not None:
            embeddings[padding_idx] = np.zeros(embeddings.shape[-1])
        embedding_matrix.weight = torch.nn.Parameter(torch.from_numpy(embeddings).float())
        if freeze:
            embedding_matrix.weight.requires_grad = False
        return embedding_matrix

    def dict2tensor(self, embedding_dict):
        if len(embedding_dict) == 1:
            return list(embedding_dict.values())[0]
        if isinstance(embedding_dict, _EmbDict) and embedding_dict.stacked is not None \
                and tuple(embedding_dict.keys()) == embedding_dict.names:
            return embedding_dict.stacked
        return torch.stack(list(embedding_dict.values()), dim=1)

    def _select(self, feature_source, feature_type):
        key = (feature_source, feature_type)
        names = []
        for feature, spec in self._feature_map.feature_specs.items():
            if feature_source and spec["source"] != feature_source:
                continue
            if feature_type and spec["type"] != feature_type:
                continue
            if feature in self.embedding_layers:
                if spec["type"] not in ("numeric", "categorical", "sequence"):
                    raise NotImplementedError
                names.append(feature)
        return key, names

    def _full_key(self):
        return (None, None)

    def forward(self, inputs, feature_source=None, feature_type=None):
        key, names = self._select(feature_source, feature_type)
        return self._embed(inputs, key, names)


class _CoreNamespace(object):
    """What `layers.X` resolves to inside core feature specs' embedding_callback strings."""

    def __init__(self, mod):
        self._mod = mod

    def __getattr__(self, name):
        if name == "MaskedAveragePooling":
            return CoreMaskedAveragePooling
        return getattr(self._mod, name)


class EmbeddingLayer(nn.Module):
    """core/pytorch/layers/embedding.py:10-27."""

    def __init__(self, feature_map, embedding_dim, disable_sharing_pretrain=False, required_feature_columns=[],
                 not_required_feature_columns=[]):
        super(EmbeddingLayer, self).__init__()
        self.embedding_layer = EmbeddingDictLayer(feature_map, embedding_dim,
                                                  disable_sharing_pretrain=disable_sharing_pretrain,
                                                  required_feature_columns=required_feature_columns,
                                                  not_required_feature_columns=not_required_feature_columns)

    def forward(self, X, feature_source=None):
        feature_emb_dict = self.embedding_layer(X, feature_source=feature_source)
        return self.embedding_layer.dict2tensor(feature_emb_dict)
