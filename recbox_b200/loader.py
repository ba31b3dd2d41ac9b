"""Input pipeline for the fused path.

The reference's ranking loader (`ranking/pytorch/dataloaders/h5_dataloader.py:25-47`) hstacks every column
of the dataset into ONE float64 `[N, n_cols]` matrix and hands `[B, n_cols]` float64 batches to
`RankingModel.train_step`, which slices 39 columns and copies each to the device (`ranking_model.py:106-116`),
where every embedding lookup casts its column back with `.long()` (`feature_embedding.py:201,204`).  Ids and
floats travel as 8-byte doubles: 320 B per Criteo sample over PCIe.

`PackedDataset` does the conversion ONCE when the data is loaded: categorical ids -> one pinned int32
`[N, F]` block (uint16 when every vocabulary is below 65 536: 108 B per Criteo sample), numeric values -> one pinned fp32 `[N, Fn]` block, label -> pinned fp32 `[N]` -- 160 B per
Criteo sample.  `PackedDataLoader` mirrors the reference `DataLoader`'s surface (`num_samples`,
`num_batches`, `len()`, iteration, `shuffle`) and yields `PackedBatch`es: three contiguous pinned slices
(shuffle: one host index-gather per block into a pinned ring), each one async H2D copy; the fused embedding
layer consumes the device blocks directly -- no per-column slicing, no cast kernels.  When the loader is bound
to the embedding layer (`bind(layer)`) the fused-table row offsets are added at load time too, and the ids
block IS the kernels' `rows` argument.

Sequence features and `meta` columns are not packed (they stay with the reference loader's dict form).
Host code only: no kernels here, `recbox_b200.ops` does the device work.
"""
import numpy as np
import torch

from ._lib import RbxError

I32, F32 = torch.int32, torch.float32


class PackedBatch(object):
    """One batch: ids int32 [B,F] (+ row offsets when `offsets` is set) -- or, in the compact form, ids16: int16 [B,F]
    holding uint16 LOCAL ids (every vocabulary < 65 536; the layer adds the row offsets on the device) --, dense fp32
    [B,Fn], labels fp32 [B]."""
    __slots__ = ("ids", "ids16", "dense", "labels", "cat_names", "num_names", "offsets", "feature_map")

    def __init__(self, ids, dense, labels, cat_names, num_names, offsets, feature_map, ids16=None):
        self.ids, self.ids16, self.dense, self.labels = ids, ids16, dense, labels
        self.cat_names, self.num_names, self.offsets, self.feature_map = cat_names, num_names, offsets, feature_map

    def __len__(self):
        for t in (self.ids, self.ids16, self.dense, self.labels):
            if t is not None:
                return t.shape[0]
        return 0

    @property
    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in (self.ids, self.ids16, self.dense, self.labels) if t is not None)

    def to(self, device, non_blocking=True):
        """Three async copies (pinned -> device) on the current stream."""
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return PackedBatch(mv(self.ids), mv(self.dense), mv(self.labels), self.cat_names, self.num_names, self.offsets,
                           self.feature_map, ids16=mv(self.ids16))

    def copy_into(self, dst, non_blocking=True):
        """Copy into a preallocated device PackedBatch of the same shape (double-buffered prefetch)."""
        for a, b in ((dst.ids, self.ids), (dst.ids16, self.ids16), (dst.dense, self.dense), (dst.labels, self.labels)):
            if a is not None:
                a.copy_(b, non_blocking=non_blocking)
        return dst

    def columns(self):
        """{feature: column view} in the reference's X_dict form (ranking_model.py:106-116)."""
        X = PackedColumns(self)
        return X


class PackedColumns(dict):
    """X_dict whose values are views into a PackedBatch; the fused layer recognises it and uses the blocks."""

    def __init__(self, batch):
        super(PackedColumns, self).__init__()
        self.packed = batch
        for i, n in enumerate(batch.cat_names):
            if batch.ids is not None:
                self[n] = batch.ids[:, i]
            else:                       # compact form: widened lazily, only if somebody indexes the column
                self[n] = _LazyU16Column(batch.ids16, i)
        for i, n in enumerate(batch.num_names):
            self[n] = batch.dense[:, i]

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _LazyU16Column):
            v = v.widen()
            dict.__setitem__(self, key, v)
        return v


class _LazyU16Column(object):
    """Column i of a compact id block, materialised as int32 on first access (generic, non-fused consumers)."""
    __slots__ = ("block", "i")

    def __init__(self, block, i):
        self.block, self.i = block, i

    def widen(self):
        return self.block[:, self.i].to(torch.int32) & 0xFFFF


def _pin(t):
    try:
        return t.pin_memory()
    except Exception:          # no CUDA runtime in this process (host-only tests): plain memory
        return t


class PackedDataset(object):
    """The dataset of h5_dataloader.py:25-47 in packed form.

    data: the reference's `[N, n_cols]` float64 array / tensor (columns = features in feature_map order, then the
    labels, as `load_data_array` builds it) or a dict {column: array} as `load_h5` returns it."""

    def __init__(self, feature_map, data, pin=True, compact="auto"):
        """compact: store the ids as uint16 (2 B instead of 4 B per id over PCIe) -- "auto": when every categorical
        vocabulary is below 65 536 (the Criteo config: 26 x 2 B + 13 x 4 B + 4 B = 108 B per sample)."""
        self.feature_map = feature_map
        feats = [(n, s) for n, s in feature_map.features.items() if s["type"] != "meta"]
        for n, s in feats:
            if s["type"] not in ("categorical", "numeric"):
                raise RbxError("PackedDataset packs categorical / numeric columns; feature %r is %r" % (n, s["type"]))
        self.cat_names = [n for n, s in feats if s["type"] == "categorical"]
        self.num_names = [n for n, s in feats if s["type"] == "numeric"]
        labels = list(feature_map.labels)
        if len(labels) != 1:
            raise RbxError("PackedDataset needs exactly one label column (ranking_model.py:118-122)")

        if isinstance(data, dict):
            col = lambda name: np.asarray(data[name]).reshape(-1)
        else:
            arr = data.numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
            if arr.ndim != 2:
                raise RbxError("PackedDataset: data must be [N, n_cols]")
            names = list(feature_map.features.keys()) + labels
            if arr.shape[1] != len(names):
                raise RbxError("PackedDataset: %d columns for %d features + labels" % (arr.shape[1], len(names)))
            index = {n: i for i, n in enumerate(names)}
            col = lambda name: arr[:, index[name]]
        N = len(col(labels[0]))
        ids = np.empty((N, len(self.cat_names)), dtype=np.int32)
        for i, n in enumerate(self.cat_names):
            c = col(n)
            v = feature_map.features[n].get("vocab_size")
            if N and (c.min() < 0 or (v is not None and c.max() >= v)):
                raise RbxError("PackedDataset: ids of %r fall outside [0, vocab_size)" % n)
            ids[:, i] = c                       # float64 -> int32 truncation == .long() on integral values
        dense = np.empty((N, len(self.num_names)), dtype=np.float32)
        for i, n in enumerate(self.num_names):
            dense[:, i] = col(n)                # == .float()
        lab = np.asarray(col(labels[0]), dtype=np.float32)
        vocabs = [feature_map.features[n].get("vocab_size") for n in self.cat_names]
        fits = bool(self.cat_names) and all(v is not None and v <= 65536 for v in vocabs)
        if compact is True and not fits:
            raise RbxError("PackedDataset(compact=True): a vocabulary is missing or above 65 536")
        self.compact = fits if compact == "auto" else bool(compact)
        self.ids = self.ids16 = None
        if self.compact:
            self.ids16 = torch.from_numpy(ids.astype(np.uint16).view(np.int16))     # uint16 bit patterns
        else:
            self.ids = torch.from_numpy(ids)
        self.dense = torch.from_numpy(dense)
        self.labels = torch.from_numpy(np.ascontiguousarray(lab))
        self.offsets = None
        if pin:
            self.dense, self.labels = _pin(self.dense), _pin(self.labels)
            if self.ids is not None:
                self.ids = _pin(self.ids)
            if self.ids16 is not None:
                self.ids16 = _pin(self.ids16)

    def __len__(self):
        return self.labels.shape[0]

    @property
    def bytes_per_sample(self):
        return (2 if self.compact else 4) * len(self.cat_names) + 4 * (len(self.num_names) + 1)

    def add_row_offsets(self, offsets):
        """Turn local ids into fused-table rows once (offsets[f] = first row of feature f's table).  The compact form
        keeps local ids (a fused-table row does not fit 16 bits): the layer adds its offsets on the device."""
        if self.offsets is not None:
            raise RbxError("row offsets were already added")
        offsets = [int(o) for o in offsets]
        if len(offsets) != len(self.cat_names):
            raise RbxError("need one offset per categorical feature")
        if self.compact:
            return
        self.ids += torch.tensor(offsets, dtype=I32)[None, :]
        self.offsets = offsets

    def batch(self, lo, hi):
        return PackedBatch(self.ids[lo:hi] if self.ids is not None else None, self.dense[lo:hi], self.labels[lo:hi],
                           self.cat_names, self.num_names, self.offsets, self.feature_map,
                           ids16=self.ids16[lo:hi] if self.ids16 is not None else None)


class PackedDataLoader(object):
    """Drop-in for h5_dataloader.py:50-59 `DataLoader(feature_map, data_path, batch_size, shuffle)` over a
    PackedDataset.  Yields pinned host `PackedBatch`es (or device ones when `device` is given: the copy of batch
    i+1 is issued on a side stream while the caller works on batch i)."""

    def __init__(self, feature_map, data, batch_size=32, shuffle=False, device=None, seed=None, ring=3, **kwargs):
        self.dataset = data if isinstance(data, PackedDataset) else PackedDataset(feature_map, data)
        self.feature_map = feature_map
        self.batch_size, self.shuffle, self.device = int(batch_size), bool(shuffle), device
        self.num_samples = len(self.dataset)
        self.num_batches = int(np.ceil(self.num_samples * 1.0 / self.batch_size))
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(seed)
        self._ring_n = max(2, int(ring))
        self._ring = None

    def __len__(self):
        return self.num_batches

    def bind(self, embedding_layer):
        """Add the fused-table row offsets of `embedding_layer` (FeatureEmbedding / FeatureEmbeddingDict) to the ids
        once, so batches feed the kernels without any conversion launch."""
        layer = getattr(embedding_layer, "embedding_layer", embedding_layer)
        self.dataset.add_row_offsets(layer.row_offsets(self.dataset.cat_names))
        return self

    # -- host side ---------------------------------------------------------------------------------
    def _host_batches(self):
        ds, bs, N = self.dataset, self.batch_size, self.num_samples
        if not self.shuffle:
            for lo in range(0, N, bs):
                yield ds.batch(lo, min(lo + bs, N))
            return
        perm = torch.randperm(N, generator=self._gen)
        if self._ring is None:
            mk = lambda t: _pin(torch.empty((bs,) + tuple(t.shape[1:]), dtype=t.dtype))
            ids_src = ds.ids16 if ds.compact else ds.ids
            self._ring = [(mk(ids_src), mk(ds.dense), mk(ds.labels)) for _ in range(self._ring_n)]
        ids_src = ds.ids16 if ds.compact else ds.ids
        for k, lo in enumerate(range(0, N, bs)):
            idx = perm[lo:lo + bs]
            n = idx.numel()
            bi, bd, bl = self._ring[k % self._ring_n]
            torch.index_select(ids_src, 0, idx, out=bi[:n])
            torch.index_select(ds.dense, 0, idx, out=bd[:n])
            torch.index_select(ds.labels, 0, idx, out=bl[:n])
            yield PackedBatch(None if ds.compact else bi[:n], bd[:n], bl[:n], ds.cat_names, ds.num_names, ds.offsets,
                              self.feature_map, ids16=bi[:n] if ds.compact else None)

    def __iter__(self):
        if self.device is None:
            for b in self._host_batches():
                yield b
            return
        # device prefetch: copy stream runs one batch ahead of the consumer
        dev = torch.device(self.device)
        copy = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        pending = None
        in_flight = []                      # copy-done events of the batches still reading a pinned ring slot
        host = self._host_batches()
        while True:
            if len(in_flight) >= self._ring_n - 1:
                in_flight.pop(0).synchronize()          # the slot the next host gather overwrites has left the host
            hb = next(host, None)
            if hb is None:
                break
            with torch.cuda.stream(copy):
                db = hb.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            in_flight.append(ev)
            if pending is not None:
                yield self._hand_over(pending, main)
            pending = (db, ev)
        if pending is not None:
            yield self._hand_over(pending, main)

    @staticmethod
    def _hand_over(pending, main):
        pb, pev = pending
        main.wait_event(pev)
        for t in (pb.ids, pb.ids16, pb.dense, pb.labels):
            if t is not None:
                t.record_stream(main)
        return pb


# =================================================================================================
# matching side: per-epoch negative sampling (h5_generator.py:72-95, 144-181)
# =================================================================================================
class EpochNegativeSampler(object):
    """TrainGenerator.negative_sampling on the GPU: `all_item_indexes = hstack([pos_item_indexes, negatives])` rebuilt
    every epoch in ONE launch instead of np.random.choice per query row (+ a multiprocessing pool and pickle round trips).

    query_indexes [N] and pos_item_indexes [N] are the columns TrainGenerator keeps (h5_generator.py:131,20);
    user2items_dict is `get_user2items_dict` (h5_generator.py:37-42) and is only needed with ignore_pos_items=True."""

    def __init__(self, num_items, query_indexes, pos_item_indexes, num_negs, user2items_dict=None, ignore_pos_items=False,
                 seed=0, device="cuda"):
        from . import ops
        self._ops = ops
        self.num_items, self.num_negs, self.seed, self.epoch = int(num_items), int(num_negs), int(seed), 0
        dev = torch.device(device)
        q = np.asarray(query_indexes).reshape(-1)
        self.n = len(q)
        self.pos = torch.as_tensor(np.asarray(pos_item_indexes).reshape(-1), dtype=torch.int64).to(dev)
        self.user_of_query = self.pos_ptr = self.pos_items = None
        if ignore_pos_items:
            if user2items_dict is None:
                raise RbxError("ignore_pos_items needs user2items_dict")
            users, inv = np.unique(q, return_inverse=True)
            lens = np.fromiter((len(user2items_dict[u]) for u in users), dtype=np.int64, count=len(users))
            ptr = np.zeros(len(users) + 1, dtype=np.int64)
            np.cumsum(lens, out=ptr[1:])
            items = np.empty(int(ptr[-1]), dtype=np.int64)
            for i, u in enumerate(users):
                items[ptr[i]:ptr[i + 1]] = np.sort(np.asarray(user2items_dict[u], dtype=np.int64))
            self.user_of_query = torch.from_numpy(inv.astype(np.int64)).to(dev)
            self.pos_ptr, self.pos_items = torch.from_numpy(ptr).to(dev), torch.from_numpy(items).to(dev)

    def sample(self):
        """-> all_item_indexes int64 [N, 1 + num_negs] on the device (column 0 = the positive); a new stream per epoch."""
        out, gave_up = self._ops.sample_negatives(self.n, self.num_negs, self.num_items, self.seed * 1000003 + self.epoch,
                                                  pos=self.pos, user_of_query=self.user_of_query, pos_ptr=self.pos_ptr,
                                                  pos_items=self.pos_items)
        self.epoch += 1
        self.gave_up = gave_up
        return out
