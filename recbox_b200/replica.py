"""Gradient exchange of data-parallel REPLICAS (SURVEY.md section 8e "replicas only": the table fits every GPU, each rank runs
the path on its own batch shard and the dense gradients are summed once per step -- the reference's multi-device mode,
DistributedDataParallel in third_party/recbole/trainer/trainer.py:48-64).

`ReplicaReducer` owns a symmetric (CUDA VMM, multicast-mapped) fp32 buffer and sums it across the ranks with the in-switch
all-reduce kernel of csrc/allreduce.cu (rbx_nvls_allreduce_f32: multimem.ld_reduce + multimem.st over NVLink / NVSwitch);
where multicast memory is not available (no NVSwitch, a gloo group in the CPU tests) it falls back to the collective of
torch.distributed and says so in `.path`."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import RbxError


class ReplicaReducer(object):
    def __init__(self, numel, device, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.numel = (int(numel) + 3) // 4 * 4
        self.device = torch.device(device)
        self.hdl, self.mc = None, 0
        self.path = "torch.distributed all_reduce"
        if self.device.type == "cuda" and self.world > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self.buffer = symm_mem.empty(self.numel, dtype=torch.float32, device=self.device)
                self.hdl = symm_mem.rendezvous(self.buffer, self.group)
                self.mc = int(self.hdl.multicast_ptr or 0)
                if self.mc:
                    self.path = "rbx_nvls_allreduce_f32 (multimem.ld_reduce / multimem.st over NVSwitch multicast memory)"
            except Exception as e:                      # no symmetric memory on this system: plain allocation + NCCL
                self.hdl, self.mc = None, 0
                self.path = "torch.distributed all_reduce (symmetric memory unavailable: %s)" % str(e)[:80]
        if self.hdl is None:
            self.buffer = torch.empty(self.numel, dtype=torch.float32, device=self.device)
        self.buffer.zero_()

    def all_reduce(self, average=False):
        """Sum `self.buffer` over the ranks, in place, on the current stream."""
        if self.world == 1:
            return self.buffer
        if self.mc:
            lib = _lib.load()
            with torch.cuda.device(self.device):
                self.hdl.barrier(channel=0)            # every rank's contribution is written
                rc = lib.rbx_nvls_allreduce_f32(ctypes.c_void_p(self.mc), self.numel, self.rank, self.world,
                                                ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
                if rc != 0:
                    raise RbxError("rbx_nvls_allreduce_f32 failed (%d): %s" % (rc, lib.rbx_last_error().decode()))
                self.hdl.barrier(channel=1)            # every slice is reduced and broadcast
        else:
            dist.all_reduce(self.buffer, group=self.group)
        if average:
            self.buffer.div_(self.world)
        return self.buffer

    def reduce_tensor(self, t, average=False):
        """All-reduce an arbitrary contiguous fp32 tensor through the symmetric buffer (copy in, reduce, copy out)."""
        n = t.numel()
        if not t.is_contiguous() or t.dtype != torch.float32:
            raise RbxError("ReplicaReducer.reduce_tensor: needs a contiguous fp32 tensor")
        if n > self.numel:
            raise RbxError("ReplicaReducer: tensor of %d floats exceeds the %d-float buffer" % (n, self.numel))
        flat = t.reshape(-1)
        self.buffer[:n].copy_(flat)
        if n < self.numel:
            self.buffer[n:].zero_()
        self.all_reduce(average)
        flat.copy_(self.buffer[:n])
        return t


class SparseRowExchange(object):
    """Gradient exchange of replicas whose batch touches only a sliver of the table ("switching to a
    sparse all_gather of (row-id, row-grad) pairs when B_loc * F << rows" -- the 10 M-row item table of configs[2], the 1 M-row
    table of configs[4]): every rank lists the rows its batch touched (rbx_unique_ids), gathers their gradient rows
    (rbx_gather_rows), all ranks all_gather the fixed-capacity (ids, rows) blocks, and each adds the OTHER ranks' rows into
    its dense gradient table (rbx_scatter_add_rows; ids of -1 pad the blocks and are skipped).  No host synchronisation: the
    capacity is the batch's id count, the unique count stays on the device.  Afterwards every replica holds the summed
    gradient on the union of the touched rows, which `exchange` returns for the touched-rows optimizer.

    kern: provider of unique_ids / gather_rows / scatter_add_rows (recbox_b200.ops; a CPU stand-in in the gloo tests)."""

    def __init__(self, R, D, max_ids, device, group=None, kern=None):
        if kern is None:
            from . import ops as kern
        self.kern = kern
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.R, self.D, self.cap = int(R), int(D), int(max_ids)
        dev = torch.device(device)
        self.ids_all = torch.empty((self.world, self.cap), dtype=torch.int32, device=dev)
        self.rows_all = torch.empty((self.world, self.cap, self.D), dtype=torch.float32, device=dev)
        self._slot = torch.arange(self.cap, device=dev)

    def exchange(self, g_table, rows):
        """g_table [R, D]: this rank's dense gradient table (its own contributions already scattered in); rows: the int32 row
        ids this rank's batch touched (any shape, duplicates welcome, at most max_ids of them).
        -> int32 [world * cap] ids touched by ANY rank (-1 = padding), for TouchedRowsOptimizer.step."""
        flat = rows.reshape(-1)
        if flat.numel() > self.cap:
            raise RbxError("SparseRowExchange: %d ids exceed the capacity %d" % (flat.numel(), self.cap))
        uniq, _, _, n_out = self.kern.unique_ids(flat, self.R, want_first=False, want_inverse=False, sync=False)
        mine = self.ids_all[self.rank]
        mine.fill_(-1)
        k = uniq.numel()
        mine[:k] = torch.where(self._slot[:k] < n_out[0], uniq.to(torch.int32), torch.full_like(uniq, -1, dtype=torch.int32))
        self.rows_all[self.rank] = self.kern.gather_rows(g_table, mine)
        if self.world > 1:
            dist.all_gather_into_tensor(self.ids_all.view(-1), mine.clone(), group=self.group)
            dist.all_gather_into_tensor(self.rows_all.view(-1), self.rows_all[self.rank].reshape(-1).clone(), group=self.group)
            for w in range(self.world):
                if w != self.rank:
                    self.kern.scatter_add_rows(self.rows_all[w], self.ids_all[w], -1, g_table)
        return self.ids_all.view(-1)
