"""Row-sharded embedding table across the GPUs of one box (BASELINE configs[3]).

The reference has no sharded embedding (its only multi-device modes are nn.DataParallel / DDP
replicas), so this module has no reference counterpart; its contract is: the sharded forward /
backward produce exactly what the single-table fused kernels (ops.embed_fm_fwd / embed_fm_bwd, the
reference-parity path) produce on the concatenated table.

Partition: global row r (= id + field offset in the fused table) lives on rank r % world at local
row r // world.  One process per GPU, torch.distributed for the plumbing.  Data paths:

  mode="stream" (product path at world > 1; csrc/shard_stream.cu) the all-to-all done by the kernels themselves with
               every random access local to the row's owner and only CONTIGUOUS runs crossing NVLink: ids are
               bucketed by owner tile by tile (one run per tile and owner), owners gather and store rows back in
               slot order, the requester scatters the runs through a shared-memory E tile into the FM sums; the
               backward writes row gradients in slot order into the owners' inboxes, owners reduce locally.
               Cross-rank barriers are one tiny kernel over peer-mapped flags (no NCCL, no host sync).

  mode="peer"  (product path)  every rank maps every other rank's table / gradient shard through
               CUDA IPC (rbx_peer_*), and ONE fused kernel per direction does compute + exchange:
               the forward gathers remote rows with plain loads over NVLink / NVSwitch, the backward
               reduces (red.global.add) straight into the owner's gradient shard.  No routing, no
               staging buffers, no collective on the data path; a barrier separates steps.
  mode="push"  (product path for tables too large for remote gathers -- random 64-byte reads into a
               multi-GB peer mapping collapse to ~10 GB/s on NVSwitch, measured) the same exchange as
               "a2a" but done by the kernels themselves: ids, rows and row gradients are streamed
               into the peers' inboxes with plain stores over NVLink (csrc/shard_push.cu); no NCCL on
               the data path, no host-side split sizes, no host synchronisation.
  mode="a2a"   (the NCCL baseline the north star names) bucket ids by owner (rbx_shard_route),
               all_to_all_single ids -> owners gather (rbx_gather_rows) -> all_to_all_single rows
               back; the un-permute is folded into the fused FM kernel by handing it the received
               buffer as its "table" and the send positions as its "rows".  Backward mirrors it.

`kern` is the kernel provider (recbox_b200.ops by default).  Host-side orchestration (split sizes,
buffer bookkeeping, pad rows) is independent of it, which is what the world_size-2 gloo tests on
CPU exercise with a stand-in provider.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import RbxError

F32, I32 = torch.float32, torch.int32


# -------------------------------------------------------------------------------------------------
# peer-visible device memory
# -------------------------------------------------------------------------------------------------
class _RawCuda(object):
    """Minimal __cuda_array_interface__ carrier so torch can view memory this library allocated."""

    def __init__(self, ptr, numel, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                         "strides": None}


class PeerBlock(object):
    """One cudaMalloc'd fp32 block (IPC-exportable), viewed as a torch tensor."""

    def __init__(self, numel, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.numel = int(numel)
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rbx_peer_alloc(ctypes.c_size_t(max(self.numel, 4) * 4), ctypes.byref(out)), self.lib)
        self.ptr = out.value
        self.tensor = torch.as_tensor(_RawCuda(self.ptr, max(self.numel, 4), self), device=self.device)[:self.numel]
        self.tensor.zero_()

    def handle(self):
        buf = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.rbx_peer_export(ctypes.c_void_p(self.ptr), buf), self.lib)
        return bytes(buf)

    def free(self):
        if self.ptr:
            self.tensor = None
            with torch.cuda.device(self.device):
                self.lib.rbx_peer_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


class SymmBlock(object):
    """The same block allocated through torch's symmetric memory (CUDA VMM: cuMemCreate / cuMemMap,
    2 MB pages on the importing side too) instead of legacy CUDA IPC.  Collective over `group`."""

    def __init__(self, numel, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.device = torch.device(device)
        self.numel = int(numel)
        self.tensor = symm_mem.empty(max(self.numel, 4), dtype=F32, device=self.device)
        self.tensor.zero_()
        self.hdl = symm_mem.rendezvous(self.tensor, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.ptr = self.tensor.data_ptr()
        self.tensor = self.tensor[:self.numel]

    def free(self):
        self.tensor = None
        self.hdl = None


class FileBlock(object):
    """Host stand-in used by the world_size-2 gloo tests (no GPU): the same peer-visible block as a file-backed
    shared mapping that every rank opens, so the stream-mode orchestration (parity buffers, slot bookkeeping,
    barrier epochs) runs unchanged with a CPU kernel provider.  Never used on a CUDA device."""
    _count = 0

    def __init__(self, numel, group=None):
        import os
        import tempfile
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [tempfile.mkdtemp(prefix="rbx_fileblock_") if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        FileBlock._count += 1
        self.numel = int(numel)
        n = max(self.numel, 4)
        paths = [os.path.join(box[0], "blk%d_r%d" % (FileBlock._count, w)) for w in range(world)]
        self.tensor = torch.from_file(paths[rank], shared=True, size=n, dtype=F32)
        self.tensor.zero_()
        dist.barrier(group=group)
        self.peer_tensors = [self.tensor if w == rank else torch.from_file(paths[w], shared=True, size=n, dtype=F32)
                             for w in range(world)]
        self.ptr = None

    def free(self):
        self.tensor = None
        self.peer_tensors = None


def open_peer(handle, device):
    lib = _lib.load()
    out = ctypes.c_void_p()
    buf = (ctypes.c_ubyte * 64)(*handle)
    with torch.cuda.device(device):
        _lib.check(lib.rbx_peer_open(buf, ctypes.byref(out)), lib)
    return out.value


def close_peer(ptr, device):
    lib = _lib.load()
    with torch.cuda.device(device):
        lib.rbx_peer_close(ctypes.c_void_p(ptr))


# -------------------------------------------------------------------------------------------------
# host-side bookkeeping shared by both modes (pure Python / integer arithmetic)
# -------------------------------------------------------------------------------------------------
def local_rows(R, world, rank):
    """Number of global rows r in [0, R) with r % world == rank."""
    return (R - rank + world - 1) // world if R > rank else 0


def shard_capacity(R, world):
    return (R + world - 1) // world


def owned_pad_rows(pad_rows, world, rank):
    """Local row numbers of the padding rows this rank owns (their gradient is defined as zero)."""
    return [p // world for p in pad_rows if p is not None and p >= 0 and p % world == rank]


class ShardedEmbeddingFM(object):
    """Fused multi-slot gather + FM + LR over a row-sharded table.

    R: rows of the global fused table, D: embedding dim.  `table`, `table_lr`, `g_table`,
    `g_table_lr` are this rank's shards ([cap, D] / [cap], cap = ceil(R / world)).
    alloc: how peer-visible memory is obtained -- "symm" (CUDA VMM via torch symmetric memory; peers map
    it with 2 MB pages) or "ipc" (legacy cudaIpc*; measured to collapse to ~10 GB/s for random access
    into multi-GB tables, kept for small tables / debugging)."""

    def __init__(self, R, D, mode="peer", group=None, device=None, with_lr=True, kern=None, max_ids=None, slack=1.5,
                 alloc="symm", layout="split", chunks=1):
        if mode not in ("peer", "push", "a2a", "stream"):
            raise RbxError("ShardedEmbeddingFM: mode must be 'stream', 'peer', 'push' or 'a2a'")
        if layout not in ("split", "rowlr", "rowpad"):
            raise RbxError("ShardedEmbeddingFM: layout must be 'split', 'rowlr' or 'rowpad'")
        if layout == "rowlr" and mode != "stream" and (mode != "peer" or not with_lr or D not in (4, 8, 16)):
            raise RbxError("layout='rowlr' (row + first-order weight in one 2D-float physical row) needs mode='peer', "
                           "with_lr=True and D in {4, 8, 16}")
        if layout == "rowpad" and mode != "stream":
            raise RbxError("layout='rowpad' (physical rows of D + 4 floats) exists in mode='stream' only")
        self.layout = layout
        self.chunks = max(1, int(chunks)) if mode == "stream" else 1
        self.R, self.D, self.mode, self.group, self.with_lr = int(R), int(D), mode, group, with_lr
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if kern is None:
            from . import ops as kern
        self.kern = kern
        if mode == "peer" and hasattr(kern, "shard_set_rank"):
            kern.shard_set_rank(self.rank)          # hint: which shard's rows are local (see rbx_shard_set_rank)
        self.alloc = alloc if (dist.is_initialized() and dist.get_world_size(group) > 1) else "ipc"
        self.cap = shard_capacity(self.R, self.world)
        self.n_local = local_rows(self.R, self.world, self.rank)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        n_t, n_l = self.cap * self.D, self.cap
        n_l4 = (n_l + 3) // 4 * 4
        self._block = None
        self._peers = []
        self._ws = self._saved = self._ptr_arrays = None
        if mode == "stream":
            if self.world & (self.world - 1) or self.world > 8:
                raise RbxError("stream mode needs a power-of-two world <= 8 (got %d)" % self.world)
            if max_ids is None:
                raise RbxError("stream mode needs max_ids (= max batch * categorical slots per rank)")
            self._init_stream(int(max_ids), float(slack))
            return
        if mode == "peer" and layout == "rowlr":
            if self.world & (self.world - 1) or self.world > 8:
                raise RbxError("peer mode needs a power-of-two world <= 8 (got %d)" % self.world)
            # one block: table_phys [cap, 2D] | g_phys [cap, 2D]; physical row = [e (D) | w_lr | zeros]
            RS = 2 * self.D
            self._offs = (0, self.cap * RS)
            self._block = self._shared_block(2 * self.cap * RS)
            flat = self._block.tensor
            tp = flat[:self.cap * RS].view(self.cap, RS)
            gp = flat[self.cap * RS:].view(self.cap, RS)
            self.table, self.table_lr = tp[:, :self.D], tp[:, self.D]
            self.g_table, self.g_table_lr = gp[:, :self.D], gp[:, self.D]
            self._gflat = flat[self.cap * RS:]
            self._flat = flat
            self._map_peers()
            self._saved = self._ws = None
            return
        if mode == "peer":
            if self.world & (self.world - 1) or self.world > 8:
                raise RbxError("peer mode needs a power-of-two world <= 8 (got %d)" % self.world)
            # one block: table | table_lr | g_table | g_table_lr   (offsets in floats, 16-byte aligned)
            self._offs = (0, n_t, n_t + n_l4, 2 * n_t + n_l4)
            self._block = self._shared_block(2 * n_t + 2 * n_l4)
            flat = self._block.tensor
        else:
            self._offs = (0, n_t, n_t + n_l4, 2 * n_t + n_l4)
            flat = torch.zeros(2 * n_t + 2 * n_l4, dtype=F32, device=self.device)
        o = self._offs
        self.table = flat[o[0]:o[0] + n_t].view(self.cap, self.D)
        self.table_lr = flat[o[1]:o[1] + n_l]
        self.g_table = flat[o[2]:o[2] + n_t].view(self.cap, self.D)
        self.g_table_lr = flat[o[3]:o[3] + n_l]
        self._gflat = flat[o[2]:o[3] + n_l]
        self._flat = flat
        self._ptr_arrays = None
        if mode == "peer":
            self._map_peers()
        self._saved = None
        self._ws = None
        if mode == "push":
            if max_ids is None:
                raise RbxError("push mode needs max_ids (= max batch * categorical slots per rank)")
            self._init_push(int(max_ids), float(slack))

    # -- stream mode ------------------------------------------------------------------------------------
    def _init_stream(self, max_ids, slack):
        """Table and gradient shards are LOCAL allocations (only their owner ever touches them); the exchange
        workspace -- flags, count / id inboxes (double-buffered by step parity), row buffer, gradient inbox -- is one
        peer-visible block."""
        W, D, cap_rows, dev = self.world, self.D, self.cap, self.device
        if self.layout == "split" or not self.with_lr:
            self.layout = "split"
            n_t, n_l4 = cap_rows * D, (cap_rows + 3) // 4 * 4
            flat = torch.zeros(2 * n_t + 2 * n_l4, dtype=F32, device=dev)
            self._tphys, self._gphys = flat[:n_t].view(cap_rows, D), flat[n_t + n_l4:2 * n_t + n_l4].view(cap_rows, D)
            self.table, self.g_table = self._tphys, self._gphys
            self.table_lr = flat[n_t:n_t + cap_rows]
            self.g_table_lr = flat[2 * n_t + n_l4:2 * n_t + n_l4 + cap_rows]
            self._gflat = flat[n_t + n_l4:]
        else:
            RS = 2 * D if self.layout == "rowlr" else D + 4
            flat = torch.zeros(2 * cap_rows * RS, dtype=F32, device=dev)
            self._tphys, self._gphys = flat[:cap_rows * RS].view(cap_rows, RS), flat[cap_rows * RS:].view(cap_rows, RS)
            self.table, self.table_lr = self._tphys[:, :D], self._tphys[:, D]
            self.g_table, self.g_table_lr = self._gphys[:, :D], self._gphys[:, D]
            self._gflat = flat[cap_rows * RS:]
        self._flat = flat
        self.max_ids = max_ids
        # `chunks` lanes: the batch is cut into that many sample ranges, each with its own workspace, barrier flags and
        # CUDA stream, so one range's NVLink-bound phases (serve, grad_push) overlap the other ranges' HBM-bound ones
        # (route, consume, apply) and the waits at the cross-rank barriers
        per_lane = (max_ids + self.chunks - 1) // self.chunks + 64 * 32
        self._lanes = [_StreamLane(self, per_lane, slack) for _ in range(self.chunks)]
        self.slot_cap = self._lanes[0].cap
        self.rowbuf = self._lanes[0].rowbuf
        self._ws = _LaneBlocks(self._lanes)
        self._lane_streams = None
        self.barrier()

    def _peer_views(self, block, off, n):
        """What the kernel provider gets for "every rank's copy of words [off, off + n) of `block`": a ctypes array of
        device pointers on CUDA; the list of the ranks' (shared-memory) tensors under the CPU stand-in of the tests."""
        if hasattr(block, "peer_tensors"):
            return [t[off:off + n] for t in block.peer_tensors]
        return (ctypes.c_void_p * self.world)(*[b + off * 4 for b in self._peer_bases(block)])

    def _mark(self, name):
        """Phase timing for bench.py / tools (off unless `self.profile` is a list): CUDA events between the phases."""
        prof = getattr(self, "profile", None)
        if prof is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            prof.append((name, ev))

    @staticmethod
    def phase_times(prof):
        """[(name, event)] -> {phase: total ms}; a phase is the span that ENDS at its mark."""
        out = {}
        for (_, e0), (n1, e1) in zip(prof[:-1], prof[1:]):
            if n1 != "begin":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out

    def _chunk_bounds(self, B):
        C = min(self.chunks, max(1, B))
        per = (B + C - 1) // C
        per = (per + 31) // 32 * 32                      # tile-aligned cuts keep every chunk's tiles whole
        return [(lo, min(lo + per, B)) for lo in range(0, B, per)]

    def _run_lanes(self, bounds, fn):
        """fn(lane, lo, hi) for every chunk; on CUDA each lane runs on its own stream, fenced against the caller's."""
        if len(bounds) == 1 or self.device.type != "cuda" or getattr(self, "profile", None) is not None:
            for lane, (lo, hi) in zip(self._lanes, bounds):
                fn(lane, lo, hi)
            return
        if self._lane_streams is None:
            self._lane_streams = [torch.cuda.Stream(device=self.device) for _ in self._lanes]
        main = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(main)
        for lane, st, (lo, hi) in zip(self._lanes, self._lane_streams, bounds):
            st.wait_event(start)
            with torch.cuda.stream(st):
                fn(lane, lo, hi)
        for st in self._lane_streams[:len(bounds)]:
            main.wait_stream(st)

    def _fwd_stream(self, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, want_E, n_slots):
        B, F = rows.shape
        if B * F > self.max_ids:
            raise RbxError("sharded stream: %d ids exceed max_ids=%d" % (B * F, self.max_ids))
        Ft = n_slots or (F + len(num_pos))
        dev, D = rows.device, self.D
        E = torch.empty((B, Ft, D), dtype=F32, device=dev) if want_E else None
        S = torch.empty((B, D), dtype=F32, device=dev)
        fm = torch.empty((B,), dtype=F32, device=dev)
        lr = torch.empty((B,), dtype=F32, device=dev) if self.with_lr else None
        bounds = self._chunk_bounds(B)

        def one(lane, lo, hi):
            lane.forward(rows[lo:hi], cat_pos, dense_x[lo:hi] if dense_x is not None else None, dense_w, dense_w_lr, num_pos,
                         lr_bias, Ft, (E[lo:hi] if E is not None else None, S[lo:hi], fm[lo:hi], lr[lo:hi] if lr is not None else None))
        self._run_lanes(bounds, one)
        self._saved = (bounds, B, F)
        self._bwd_pending = False
        return E, S, fm, lr

    def _bwd_stream(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                    g_dense_w, g_dense_w_lr, g_lr_bias, n_slots):
        if self._saved is None:
            raise RbxError("ShardedEmbeddingFM.backward (stream) needs the forward of the same batch first")
        bounds, B, F = self._saved
        if tuple(rows.shape) != (B, F):
            raise RbxError("ShardedEmbeddingFM.backward (stream): rows differ in shape from the forward's")
        Ft = n_slots or (F + len(num_pos))
        again, self._bwd_pending = self._bwd_pending, True
        cut = lambda t, lo, hi: None if t is None else t[lo:hi]

        def one(lane, lo, hi):
            lane.backward(rows[lo:hi], cat_pos, pad_rows, cut(dense_x, lo, hi), dense_w, num_pos, cut(E, lo, hi), S[lo:hi],
                          cut(dE, lo, hi), cut(d_fm, lo, hi), cut(d_lr, lo, hi), g_dense_w, g_dense_w_lr, g_lr_bias, Ft, again)
        self._run_lanes(bounds, one)


    # -- push-mode workspace: inboxes every peer can write ---------------------------------------------
    def _init_push(self, max_ids, slack):
        W, D = self.world, self.D
        if W > 8:
            raise RbxError("push mode covers up to 8 ranks (got %d)" % W)
        self.max_ids = max_ids
        self.slot_cap = min(max_ids, int(max_ids / W * slack) + 1024) if W > 1 else max_ids
        al = lambda n: (n + 3) // 4 * 4
        sizes = [("inbox_ids", al(W * self.slot_cap)), ("inbox_meta", al(W * 4)), ("rowbuf", al(max_ids * D)),
                 ("rowbuf_lr", al(max_ids)), ("ginbox", al(W * self.slot_cap * D)), ("ginbox_lr", al(W * self.slot_cap))]
        offs, tot = {}, 0
        for name, n in sizes:
            offs[name] = (tot, n)
            tot += n
        self._ws = self._shared_block(tot)
        flat = self._ws.tensor
        v = {name: flat[o:o + n] for name, (o, n) in offs.items()}
        self.inbox_ids = v["inbox_ids"].view(I32)[:W * self.slot_cap]
        self.inbox_meta = v["inbox_meta"].view(I32)[:W * 4]
        self.rowbuf = v["rowbuf"][:max_ids * D].view(max_ids, D)
        self.rowbuf_lr = v["rowbuf_lr"][:max_ids]
        self.ginbox = v["ginbox"][:W * self.slot_cap * D]
        self.ginbox_lr = v["ginbox_lr"][:W * self.slot_cap]
        self.gsend = torch.zeros((max_ids, D), dtype=F32, device=self.device)
        self.gsend_lr = torch.zeros(max_ids, dtype=F32, device=self.device)
        base = self._peer_bases(self._ws)
        self._ptr_arrays = {name: (ctypes.c_void_p * W)(*[b + o * 4 for b in base]) for name, (o, n) in offs.items()}
        self._pad_cache = {}
        self.barrier()

    def check_overflow(self):
        """Host check (synchronises): did any (owner, requester) bucket exceed its slot capacity?"""
        if self.mode == "push" and int(self.inbox_meta.view(self.world, 4)[:, 2].max()) != 0:
            raise RbxError("sharded push: an id bucket exceeded the slot capacity %d; raise `slack`" % self.slot_cap)
        if self.mode == "stream" and any(int(lane.overflow.item()) != 0 for lane in self._lanes):
            raise RbxError("sharded stream: an id lane exceeded the slot capacity %d; raise `slack`" % self.slot_cap)

    # -- peer mapping ------------------------------------------------------------------------------
    def _shared_block(self, numel):
        if self.device.type != "cuda":
            return FileBlock(numel, self.group)
        if self.alloc == "symm":
            return SymmBlock(numel, self.device, self.group)
        return PeerBlock(numel, self.device)

    def _peer_bases(self, block):
        """Device pointers (valid on this device) of every rank's copy of `block`."""
        if isinstance(block, SymmBlock):
            return list(block.ptrs)
        base = [None] * self.world
        base[self.rank] = block.ptr
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, block.handle(), group=self.group)
            for w, h in enumerate(handles):
                if w != self.rank:
                    base[w] = open_peer(h, self.device)
                    self._peers.append(base[w])
        return base

    def _map_peers(self):
        base = self._peer_bases(self._block)
        o = self._offs
        mk = lambda off: (ctypes.c_void_p * self.world)(*[b + off * 4 for b in base])
        if self.layout == "rowlr":
            self._ptr_arrays = {"table": mk(o[0]), "g_table": mk(o[1])}
            return
        self._ptr_arrays = {"table": mk(o[0]), "table_lr": mk(o[1]), "g_table": mk(o[2]), "g_table_lr": mk(o[3])}

    def close(self):
        for p in self._peers:
            close_peer(p, self.device)
        self._peers = []
        if self._ws is not None:
            self.barrier()
            self.inbox_ids = self.inbox_meta = self.rowbuf = self.rowbuf_lr = self.ginbox = self.ginbox_lr = None
            self._lanes = None
            self._ws.free()
            self._ws = None
        if self._block is not None:
            self.barrier()
            self.table = self.table_lr = self.g_table = self.g_table_lr = self._gflat = self._flat = None
            self._block.free()
            self._block = None

    # -- parameter plumbing -------------------------------------------------------------------------
    def load_global(self, table, table_lr=None):
        """Fill this rank's shard from a full [R, D] (and [R]) table (tests / checkpoints)."""
        mine = table[self.rank::self.world].to(self.device, F32)
        self.table[:mine.shape[0]].copy_(mine)                    # (strided views in the rowlr layout)
        if table_lr is not None:
            m1 = table_lr.reshape(-1)[self.rank::self.world].to(self.device, F32)
            self.table_lr[:m1.shape[0]].copy_(m1)
        self.barrier()

    def gather_global(self, which="g_table"):
        """All ranks' shards of `which` re-interleaved to the global row order (tests / checkpoints)."""
        local = getattr(self, which).contiguous()                 # (the rowlr layout hands out strided views)
        parts = [torch.empty_like(local) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(parts, local, group=self.group)
        else:
            parts = [local]
        out = torch.empty((self.cap * self.world,) + tuple(local.shape[1:]), dtype=F32, device=local.device)
        for w, p in enumerate(parts):
            out[w::self.world] = p
        return out[:self.R]

    def zero_grad(self):
        self._gflat.zero_()

    def device_barrier(self):
        """Stream-ordered cross-rank barrier (a 1-element all-reduce on the current stream): every
        rank's earlier kernels -- including their reductions into this rank's gradient shard -- have
        completed before anything enqueued after it starts.  No host synchronisation."""
        if self.world > 1:
            if getattr(self, "_tok", None) is None:
                self._tok = torch.zeros(1, dtype=F32, device=self.device)
            dist.all_reduce(self._tok, group=self.group)

    def barrier(self):
        if self.world > 1:
            if self._flat is not None and self._flat.is_cuda:
                torch.cuda.current_stream(self.device).synchronize()
            dist.barrier(group=self.group)

    # -- forward / backward ---------------------------------------------------------------------------
    def forward(self, rows, cat_pos, dense_x=None, dense_w=None, dense_w_lr=None, num_pos=(), lr_bias=None,
                want_E=True, n_slots=None):
        """rows: int32 [B, F] GLOBAL row ids of this rank's batch shard.  Returns (E, S, fm, lr)."""
        if self.mode == "stream":
            out = self._fwd_stream(rows, cat_pos, dense_x, dense_w, dense_w_lr, list(num_pos), lr_bias, want_E, n_slots)
        elif self.mode == "peer":
            out = self._fwd_peer(rows, cat_pos, dense_x, dense_w, dense_w_lr, list(num_pos), lr_bias, want_E, n_slots)
        elif self.mode == "push":
            out = self._fwd_push(rows, cat_pos, dense_x, dense_w, dense_w_lr, list(num_pos), lr_bias, want_E, n_slots)
        else:
            out = self._fwd_a2a(rows, cat_pos, dense_x, dense_w, dense_w_lr, list(num_pos), lr_bias, want_E, n_slots)
        return out

    def backward(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                 g_dense_w=None, g_dense_w_lr=None, g_lr_bias=None, n_slots=None):
        """Accumulates into this rank's AND (peer mode) the owners' gradient shards; the dense
        (replicated) gradients g_dense_* are local partial sums the caller all-reduces."""
        if self.mode == "stream":
            self._bwd_stream(rows, cat_pos, pad_rows, dense_x, dense_w, list(num_pos), E, S, dE, d_fm, d_lr,
                             g_dense_w, g_dense_w_lr, g_lr_bias, n_slots)
        elif self.mode == "peer":
            self._bwd_peer(rows, cat_pos, pad_rows, dense_x, dense_w, list(num_pos), E, S, dE, d_fm, d_lr,
                           g_dense_w, g_dense_w_lr, g_lr_bias, n_slots)
        elif self.mode == "push":
            self._bwd_push(rows, cat_pos, pad_rows, dense_x, dense_w, list(num_pos), E, S, dE, d_fm, d_lr,
                           g_dense_w, g_dense_w_lr, g_lr_bias, n_slots)
        else:
            self._bwd_a2a(rows, cat_pos, pad_rows, dense_x, dense_w, list(num_pos), E, S, dE, d_fm, d_lr,
                          g_dense_w, g_dense_w_lr, g_lr_bias, n_slots)

    # ---- peer mode: one fused kernel per direction, exchange inside ----------------------------------
    def _fwd_peer(self, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, want_E, n_slots):
        if self.layout == "rowlr":
            return self.kern.embed_fm_fwd_sharded_rowlr(self._ptr_arrays["table"], self.world, rows, cat_pos, dense_x,
                                                        dense_w, dense_w_lr, num_pos, lr_bias, self.R, self.D,
                                                        want_E=want_E, n_slots=n_slots)
        return self.kern.embed_fm_fwd_sharded(self._ptr_arrays["table"], self._ptr_arrays["table_lr"] if self.with_lr else None,
                                              self.world, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias,
                                              self.R, self.D, want_E=want_E, want_lr=self.with_lr, n_slots=n_slots)

    def _bwd_peer(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                  g_dense_w, g_dense_w_lr, g_lr_bias, n_slots):
        if self.layout == "rowlr":
            self.kern.embed_fm_bwd_sharded_rowlr(self._ptr_arrays["table"], self._ptr_arrays["g_table"], self.world, rows,
                                                 cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                                                 g_dense_w, g_dense_w_lr, g_lr_bias, self.R, self.D, n_slots=n_slots)
            return
        self.kern.embed_fm_bwd_sharded(self._ptr_arrays["table"], self._ptr_arrays["g_table"],
                                       self._ptr_arrays["g_table_lr"] if self.with_lr else None, self.world, rows,
                                       cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm,
                                       d_lr if self.with_lr else None, g_dense_w, g_dense_w_lr, g_lr_bias,
                                       self.R, self.D, n_slots=n_slots)

    # ---- push mode: the exchange done by the kernels, stores over NVLink into the peers' inboxes ---------
    def _fwd_push(self, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, want_E, n_slots):
        k, pa = self.kern, self._ptr_arrays
        B, F = rows.shape
        N = B * F
        if N > self.max_ids:
            raise RbxError("sharded push: %d ids exceed max_ids=%d" % (N, self.max_ids))
        send, pos, counts = k.shard_route(rows.reshape(-1), self.world)
        k.shard_push_ids(send, counts, pa["inbox_ids"], pa["inbox_meta"], self.rank, self.world, self.slot_cap)
        self.device_barrier()                      # every requester's ids are in my inbox
        k.shard_serve_rows(self.table, self.table_lr if self.with_lr else None, self.inbox_ids, self.inbox_meta,
                           pa["rowbuf"], pa["rowbuf_lr"] if self.with_lr else None, self.world, self.slot_cap)
        self.device_barrier()                      # every owner's rows are in my row buffer
        E, S, fm, lr = k.embed_fm_fwd(self.rowbuf[:N], self.rowbuf_lr[:N] if self.with_lr else None, pos.view(B, F), cat_pos,
                                      dense_x, dense_w, dense_w_lr, num_pos, lr_bias, want_E=want_E, want_lr=self.with_lr,
                                      n_slots=n_slots)
        self._saved = (pos, counts)
        return E, S, fm, lr

    def _bwd_push(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                  g_dense_w, g_dense_w_lr, g_lr_bias, n_slots):
        if self._saved is None:
            raise RbxError("ShardedEmbeddingFM.backward (push) needs the forward of the same batch first")
        k, pa, D = self.kern, self._ptr_arrays, self.D
        pos, counts = self._saved
        B, F = rows.shape
        N = B * F
        use_lr = self.with_lr and d_lr is not None
        gsend, gsend_lr = self.gsend[:N], (self.gsend_lr[:N] if use_lr else None)
        gsend.zero_()
        if use_lr:
            gsend_lr.zero_()
        k.embed_fm_bwd(self.rowbuf[:N], pos.view(B, F), cat_pos, None, dense_x, dense_w, num_pos, E, S, dE, d_fm,
                       d_lr if use_lr else None, gsend, gsend_lr, g_dense_w, g_dense_w_lr, g_lr_bias, D, N, n_slots=n_slots)
        k.shard_push_grads(gsend, gsend_lr, counts, pa["ginbox"], pa["ginbox_lr"] if use_lr else None, self.rank,
                           self.world, self.slot_cap)
        self.device_barrier()                      # every requester's gradients are in my inbox
        key = tuple(pad_rows or ())
        if key not in self._pad_cache:
            mine = owned_pad_rows(pad_rows or [], self.world, self.rank)
            self._pad_cache[key] = torch.tensor(mine, dtype=I32, device=self.device) if mine else None
        k.shard_apply_grads(self.ginbox, self.ginbox_lr if use_lr else None, self.inbox_ids, self.inbox_meta, self.g_table,
                            self.g_table_lr if use_lr else None, self.world, self.slot_cap, self._pad_cache[key])
        self.device_barrier()                      # inboxes are free for the next step

    # ---- a2a mode: route -> all_to_all ids -> gather -> all_to_all rows -> fused FM on the received buffer --
    def _exchange_counts(self, counts):
        """counts[w] = ids this rank sends to w  ->  (send_sizes, recv_sizes) as Python lists."""
        recv = torch.empty_like(counts)
        if self.world > 1:
            dist.all_to_all_single(recv, counts, group=self.group)
        else:
            recv.copy_(counts)
        both = torch.stack([counts, recv]).cpu().tolist()         # the one host sync of the step
        return [int(x) for x in both[0]], [int(x) for x in both[1]]

    def _a2a(self, out, inp, out_sizes, in_sizes):
        if self.world > 1:
            dist.all_to_all_single(out, inp, out_sizes, in_sizes, group=self.group)
        else:
            out.copy_(inp)
        return out

    def _fwd_a2a(self, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, want_E, n_slots):
        k, D = self.kern, self.D
        B, F = rows.shape
        N = B * F
        send, pos, counts = k.shard_route(rows.reshape(-1), self.world)
        send_sizes, recv_sizes = self._exchange_counts(counts)
        Nr = sum(recv_sizes)
        ids_recv = self._a2a(torch.empty(Nr, dtype=I32, device=rows.device), send, recv_sizes, send_sizes)
        rows_out = k.gather_rows(self.table, ids_recv)                                   # [Nr, D]
        got = self._a2a(torch.empty((N, D), dtype=F32, device=rows.device), rows_out, send_sizes, recv_sizes)
        got_lr = None
        if self.with_lr:
            lr_out = k.gather_rows(self.table_lr.view(-1, 1), ids_recv).view(-1)        # [Nr]
            got_lr = self._a2a(torch.empty(N, dtype=F32, device=rows.device), lr_out, send_sizes, recv_sizes)
        # un-permute folded into the fused kernel: its "table" is the received buffer, its "rows" the send positions
        E, S, fm, lr = k.embed_fm_fwd(got, got_lr, pos.view(B, F), cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias,
                                      want_E=want_E, want_lr=self.with_lr, n_slots=n_slots)
        self._saved = (pos, ids_recv, send_sizes, recv_sizes, got)
        return E, S, fm, lr

    def _bwd_a2a(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                 g_dense_w, g_dense_w_lr, g_lr_bias, n_slots):
        if self._saved is None:
            raise RbxError("ShardedEmbeddingFM.backward (a2a) needs the forward of the same batch first")
        k, D = self.kern, self.D
        pos, ids_recv, send_sizes, recv_sizes, got = self._saved
        B, F = rows.shape
        N, Nr = B * F, ids_recv.numel()
        dev = rows.device
        # gradients computed straight into send order (each position is written by exactly one (b, f))
        gsend = torch.zeros((N, D), dtype=F32, device=dev)
        gsend_lr = torch.zeros(N, dtype=F32, device=dev) if (self.with_lr and d_lr is not None) else None
        k.embed_fm_bwd(got, pos.view(B, F), cat_pos, None, dense_x, dense_w, num_pos, E, S, dE, d_fm,
                       d_lr if self.with_lr else None, gsend, gsend_lr, g_dense_w, g_dense_w_lr, g_lr_bias, D, N,
                       n_slots=n_slots)
        grecv = self._a2a(torch.empty((Nr, D), dtype=F32, device=dev), gsend, recv_sizes, send_sizes)
        k.scatter_add_rows(grecv, ids_recv, None, self.g_table)
        if gsend_lr is not None:
            grecv_lr = self._a2a(torch.empty(Nr, dtype=F32, device=dev), gsend_lr, recv_sizes, send_sizes)
            k.scatter_add_rows(grecv_lr.view(-1, 1), ids_recv, None, self.g_table_lr.view(-1, 1))
        # padding rows: nn.Embedding(padding_idx) defines their gradient as zero; the owner clears them
        mine = owned_pad_rows(pad_rows or [], self.world, self.rank)
        if mine:
            idx = torch.tensor(mine, dtype=torch.long, device=dev)
            self.g_table.index_fill_(0, idx, 0.0)
            self.g_table_lr.index_fill_(0, idx, 0.0)


class _LaneBlocks(object):
    """The lanes' peer-visible workspaces behind the one `free()` ShardedEmbeddingFM.close() calls."""

    def __init__(self, lanes):
        self.lanes = lanes

    def free(self):
        for lane in self.lanes:
            lane.free()


class _StreamLane(object):
    """One sample range of the streamed exchange: workspace (flags, count / id inboxes double-buffered by step parity,
    row buffer, gradient inbox -- one peer-visible block), slot bookkeeping and the phase sequence."""

    def __init__(self, owner, max_ids, slack):
        self.o = owner
        W, D, dev = owner.world, owner.D, owner.device
        self.max_ids = max_ids
        self.cap = cap = min(max_ids, int(max_ids / W * slack) + 1024) if W > 1 else max_ids
        al = lambda n: (n + 3) // 4 * 4
        sizes = [("flags", 8), ("meta", 16), ("inbox_ids", al(2 * W * cap)), ("rowbuf", al(W * cap * D)),
                 ("rowbuf_lr", al(W * cap)), ("ginbox", al(W * cap * D)), ("ginbox_lr", al(W * cap))]
        offs, tot = {}, 0
        for name, n in sizes:
            offs[name] = (tot, n)
            tot += n
        self.block = owner._shared_block(tot)
        wsf = self.block.tensor
        v = {name: wsf[o:o + n] for name, (o, n) in offs.items()}
        self.meta = v["meta"].view(I32).view(2, 8)
        self.inbox_ids = v["inbox_ids"].view(I32)[:2 * W * cap].view(2, W * cap)
        self.rowbuf = v["rowbuf"][:W * cap * D]
        self.rowbuf_lr = v["rowbuf_lr"][:W * cap]
        self.ginbox = v["ginbox"][:W * cap * D]
        self.ginbox_lr = v["ginbox_lr"][:W * cap]
        self.peers = {name: owner._peer_views(self.block, o, n) for name, (o, n) in offs.items()}
        for par in (0, 1):
            self.peers[("meta", par)] = owner._peer_views(self.block, offs["meta"][0] + 8 * par, 8)
            self.peers[("inbox_ids", par)] = owner._peer_views(self.block, offs["inbox_ids"][0] + par * W * cap, W * cap)
        self.cursor = torch.zeros(8, dtype=I32, device=dev)
        self.overflow = torch.zeros(1, dtype=I32, device=dev)
        self.tiles = None
        self.step = self.epoch = 0
        self.saved = None

    def free(self):
        self.meta = self.inbox_ids = self.rowbuf = self.rowbuf_lr = self.ginbox = self.ginbox_lr = self.peers = None
        self.block.free()

    def tile_state(self, B, F):
        o = self.o
        T = o.kern.xs_tile_samples(F, o.D)
        if T <= 0:
            raise RbxError("sharded stream mode: F=%d, D=%d is outside what the tile kernels cover" % (F, o.D))
        n_tiles = (B + T - 1) // T
        st = self.tiles
        if st is None or st[0] != (T, F) or st[1].shape[0] < n_tiles:
            dev = o.device
            st = ((T, F), torch.zeros((n_tiles, 8), dtype=I32, device=dev), torch.zeros((n_tiles, 8), dtype=I32, device=dev),
                  torch.zeros((n_tiles * T * F,), dtype=torch.int16, device=dev))
            self.tiles = st
        return st

    def barrier(self, par=None, with_counts=False):
        o = self.o
        self.epoch += 1
        o.kern.xs_barrier(self.peers["flags"], self.peers[("meta", par)] if with_counts else None,
                          self.cursor if with_counts else None, o.rank, o.world, self.epoch, o.device)

    def forward(self, rows, cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, Ft, out):
        o, k = self.o, self.o.kern
        B, F = rows.shape
        if B * F > self.max_ids:
            raise RbxError("sharded stream: a chunk of %d ids exceeds its lane's %d" % (B * F, self.max_ids))
        _, tile_base, tile_cnt, pair_sorted = self.tile_state(B, F)
        par = self.step & 1
        self.step += 1
        cap, W = self.cap, o.world
        o._mark("begin")
        k.xs_route(rows, o.R, o.D, o.rank, W, cap, self.cursor, tile_base, tile_cnt, pair_sorted, self.overflow,
                   self.peers[("inbox_ids", par)])
        o._mark("route")
        self.barrier(par, with_counts=True)              # every requester's ids and counts are in my inboxes
        o._mark("barrier_ids")
        in_row = o.layout != "split"
        k.xs_serve(o._tphys, o.D, o.table_lr if (o.with_lr and not in_row) else None, o.with_lr and in_row,
                   self.inbox_ids[par], self.meta[par], cap, o.rank, W, self.peers["rowbuf"],
                   self.peers["rowbuf_lr"] if o.with_lr else None)
        o._mark("serve")
        self.barrier()                                   # every owner's rows are in my row buffer
        o._mark("barrier_rows")
        k.xs_consume(self.rowbuf, self.rowbuf_lr if o.with_lr else None, tile_base, tile_cnt, pair_sorted,
                     cat_pos, dense_x, dense_w, dense_w_lr, num_pos, lr_bias, B, cap, o.D, W,
                     want_E=out[0] is not None, want_lr=o.with_lr, n_slots=Ft, out=out)
        o._mark("consume")
        self.saved = (par, B, F)

    def backward(self, rows, cat_pos, pad_rows, dense_x, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                 g_dense_w, g_dense_w_lr, g_lr_bias, Ft, again):
        o, k = self.o, self.o.kern
        par, B, F = self.saved
        if tuple(rows.shape) != (B, F):
            raise RbxError("ShardedEmbeddingFM.backward (stream): chunk shapes differ from the forward's")
        _, tile_base, tile_cnt, pair_sorted = self.tiles
        cap, W = self.cap, o.world
        use_lr = o.with_lr and d_lr is not None
        if again:
            self.barrier()       # a second backward on the same forward: the owners may still be reading the last one's inbox
        o._mark("begin")
        # numeric slots / bias are batch reductions with no exchange: they ride in the same launch when they fit
        fuse_num = len(num_pos) * (o.D // 4) <= 256 and len(num_pos) <= 64
        k.xs_grad_push(E, self.rowbuf, S, dE, d_fm, d_lr, rows, pad_rows, tile_base, tile_cnt,
                       pair_sorted, cat_pos, cap, o.D, Ft, o.rank, W, self.peers["ginbox"],
                       self.peers["ginbox_lr"] if use_lr else None,
                       dense_x=dense_x if fuse_num else None, dense_w=dense_w if fuse_num else None,
                       num_pos=num_pos if fuse_num else (), g_dense_w=g_dense_w if fuse_num else None,
                       g_dense_w_lr=g_dense_w_lr if fuse_num else None, g_lr_bias=g_lr_bias if fuse_num else None)
        o._mark("grad_push")
        if not fuse_num:         # (the F = 0 form of the fused backward)
            k.embed_fm_bwd(None, None, [], None, dense_x, dense_w, num_pos, None, S, dE, d_fm, d_lr, None, None,
                           g_dense_w, g_dense_w_lr, g_lr_bias, o.D, o.R, B=B, n_slots=Ft)
        o._mark("numeric_slots")
        self.barrier()                                   # every requester's gradients are in my inbox
        o._mark("barrier_grads")
        in_row = o.layout != "split"
        k.xs_apply(self.ginbox, self.ginbox_lr if use_lr else None, self.inbox_ids[par], self.meta[par], cap, W,
                   o._gphys, o.D, o.g_table_lr if (use_lr and not in_row) else None, use_lr and in_row)
        o._mark("apply")
