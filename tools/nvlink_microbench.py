"""All-to-all WRITE bandwidth over NVLink / NVSwitch as the sharded exchange sees it (torchrun, one rank per GPU):
every rank pushes `mb` MB to each peer's inbox at once --
  kernel : rbx_shard_push_grads (plain coalesced 16-byte stores from one kernel, grid.y = destination)
  ce     : torch copy_ into the peers' symmetric-memory buffers on one stream per destination (copy engines)
Prints GB/s per GPU per direction (payload crossing NVLink = (W-1)/W of what is pushed).  Reference: 770 GB/s measured
peer copy (B200_PROFILING.md)."""
import argparse
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=14.4)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from recbox_b200 import ops, sharded
    import torch.distributed._symmetric_memory as symm_mem
    D = 16
    cap = int(args.mb * 1e6 / (D * 4))
    blk = sharded.SymmBlock(world * cap * D, dev)
    ptrs = (ctypes.c_void_p * world)(*blk.ptrs)
    gsend = torch.randn(world * cap, D, device=dev)
    counts = torch.full((world,), cap, dtype=torch.int32, device=dev)
    out = {}

    def timed(fn, name):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = {"ms": float(ms), "gbs_out_per_gpu": (world - 1) * cap * D * 4 / float(ms) / 1e6}

    timed(lambda: ops.shard_push_grads(gsend, None, counts, ptrs, None, rank, world, cap), "kernel_st_v4")
    streams = [torch.cuda.Stream() for _ in range(world)]
    peers = [blk.hdl.get_buffer(w, (world * cap * D,), torch.float32) for w in range(world)]
    flat = gsend.view(-1)

    def ce():
        main_s = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main_s)
        for w in range(world):
            if w == rank:
                continue
            with torch.cuda.stream(streams[w]):
                streams[w].wait_event(ev)
                peers[w][rank * cap * D:(rank + 1) * cap * D].copy_(flat[w * cap * D:(w + 1) * cap * D], non_blocking=True)
        for w in range(world):
            main_s.wait_stream(streams[w])
    timed(ce, "copy_engine")
    # the replica-mode gradient exchange of configs[1]: NCCL all-reduce of the fused 68 MB gradient buffer
    gbuf = torch.randn(17000204, device=dev)
    timed(lambda: dist.all_reduce(gbuf), "nccl_all_reduce_68MB")
    out["nccl_all_reduce_68MB"]["busbw_gbs"] = 2 * (world - 1) / world * gbuf.numel() * 4 / out["nccl_all_reduce_68MB"]["ms"] / 1e6
    out["nccl_all_reduce_68MB"].pop("gbs_out_per_gpu")
    if rank == 0:
        print(json.dumps({"world": world, "mb_per_peer": args.mb, **out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
