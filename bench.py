#!/usr/bin/env python
"""bench.py -- samples/sec of the RecBox embedding + FM hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--ids uniform|zipf]

A "step" is one pass of the hot path over one Criteo-shaped synthetic batch (configs[1]: DeepFM,
26 categorical + 13 numeric fields, 26 x 38 462 rows = 1 000 012-row fused table, D = 16,
B = 65 536 per GPU):
    forward  : fused multi-slot gather + numeric Linear(1,D) + FM product_sum + LR  -> E, S, fm, lr
    backward : zero the dense grad tables, then scatter-add dE + d_fm*(S-e) (and d_lr) into them,
               plus the numeric-slot / bias batch reductions
The dense MLP tail (a true GEMM) is outside the path: its input gradient dE
is a device-resident stand-in.  `value` times the step with inputs resident in HBM; `e2e` times
the same step fed from a pinned HOST float64 batch matrix (what the reference's DataLoader hands
to train_step) through H2D + rbx_split_batch_f64, with the logits read back to the host.
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU ops on all host
cores) on the same step.  Under torchrun each rank runs an independent replica of the path on its
own batch shard (weak scaling, no data-path collective; DESIGN.md "Multi-GPU").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(B=65536, F=26, Fn=13, D=16, V=38462)
# algorithmic bytes per sample, fp32 rows / int32 ids
def bytes_fwd(F, Fn, D): return F * (4 + 4 * D) + (F + Fn) * 4 * D + 4 * F + 4 * Fn + 4
def bytes_bwd(F, Fn, D): return 4 * F + (F + Fn) * 4 * D + F * 4 * D + F * 8 * D + 8 * F + 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu
    capture of this workload (profiles/r*_traffic.json, newest round); None if not captured."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        try:
            with open(path) as f:
                d = json.load(f)
            if kernel in d:
                return d[kernel]["dram_bytes_read"] + d[kernel]["dram_bytes_write"]
        except Exception:
            pass
    return None


def make_ids(B, F, V, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "zipf":
        ids = np.minimum(rng.zipf(1.05, size=(B, F)), V - 1)
    else:
        ids = rng.integers(1, V, size=(B, F))          # pad row 0 never sampled (SURVEY 8d cfg 2)
    return ids.astype(np.int64)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML in-process every 2 ms
    (the timed region is tens of milliseconds; spawning nvidia-smi takes longer than that), nvidia-smi as fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = [bool(bits & getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 bool(bits & getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 bool(bits & getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 bool(bits & getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4))]
        return [mhz, self.max_mhz] + ["Active" if f else "Not Active" for f in flags]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    time.sleep(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = [n for i, n in enumerate(self.NAMES) if any(str(s[2 + i]).lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's restatement of the same step, on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_step_factory(B, ids_kind, seed):
    from collections import OrderedDict
    from oracle import recbox_oracle as oracle     # checker / baseline only (never the product path)
    F, Fn, D, V = CFG["F"], CFG["Fn"], CFG["D"], CFG["V"]
    g = torch.Generator().manual_seed(seed)
    feats, W, W1, X = OrderedDict(), OrderedDict(), OrderedDict(), OrderedDict()
    ids = make_ids(B, F, V, ids_kind, seed)
    for n in range(Fn):
        name = "I%d" % (n + 1)
        feats[name] = {"type": "numeric", "source": ""}
        W[name] = (torch.randn(D, 1, generator=g) * 0.1).requires_grad_(True)
        W1[name] = (torch.randn(1, 1, generator=g) * 0.1).requires_grad_(True)
        X[name] = torch.rand(B, generator=g, dtype=torch.float64)
    for f in range(F):
        name = "C%d" % (f + 1)
        feats[name] = {"type": "categorical", "source": "", "vocab_size": V, "padding_idx": 0}
        W[name] = (torch.randn(V, D, generator=g) * 0.01).requires_grad_(True)
        W1[name] = (torch.randn(V, 1, generator=g) * 0.01).requires_grad_(True)
        X[name] = torch.from_numpy(ids[:, f].astype(np.float64))
    bias = torch.zeros(1, requires_grad=True)
    dE = torch.randn(B, F + Fn, D, generator=g) * 1e-3
    d_out = torch.randn(B, 1, generator=g) * 1e-3
    params = list(W.values()) + list(W1.values()) + [bias]

    def step():
        for p in params:
            p.grad = None                                          # optimizer.zero_grad()
        E = oracle.dict2tensor(oracle.embed_dict(X, feats, W))     # FeatureEmbedding
        y = oracle.factorization_machine(X, E, feats, W1, bias)    # FactorizationMachine
        torch.autograd.backward([E, y], [dE, d_out])               # dense embedding grads
        return y
    return step


def run_cpu(steps, warmup, B, ids_kind):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_factory(B, ids_kind, 20242)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return B / dt, dt


def main_reference(args, rank, world):
    if rank != 0:
        return
    B = CFG["B"]
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    sps, dt = run_cpu(steps, warmup, B, args.ids)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "samples/sec on Criteo-shaped synthetic (embedding + FM hot path, fwd+bwd)",
        "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.ids, 1),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": "%d step(s) of the full B=%d batch through oracle/recbox_oracle.py (torch %s CPU ops)" % (steps, B, torch.__version__)},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(ids_kind, world):
    return {"workload": "BASELINE configs[1] DeepFM hot path: 26 cat + 13 dense, 26x38462 = 1000012-row fused table, "
                        "D=16, B=65536 per GPU; fused gather+FM+LR fwd, dense-grad zero + scatter-add bwd; MLP tail outside the path",
            "global_batch": CFG["B"] * world, "ids": ids_kind, "batches_rotated": 4,
            "l2": "working set per step (E 163 MB + dE 163 MB + table/grad 136 MB) exceeds the 126 MB L2; 4 id batches rotate",
            "parallelism": "dp%d replicas, no data-path collective" % world}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main_b200(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from recbox_b200 import ops
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    l2_persist = ops.l2_set_persisting_bytes(int(os.environ.get("RBX_L2_PERSIST_MB", "0")) << 20)
    B, F, Fn, D, V = CFG["B"], CFG["F"], CFG["Fn"], CFG["D"], CFG["V"]
    Ft, R = F + Fn, F * V
    g = torch.Generator().manual_seed(20240 + 2 + rank)
    table = (torch.randn(R, D, generator=g) * 0.01).to(dev)
    table_lr = (torch.randn(R, generator=g) * 0.01).to(dev)
    dense_w = (torch.randn(Fn, D, generator=g) * 0.1).to(dev)
    dense_w_lr = (torch.randn(Fn, generator=g) * 0.1).to(dev)
    bias = torch.zeros(1, device=dev)
    field_off = [f * V for f in range(F)]
    cat_pos = list(range(Fn, Ft))            # Criteo order: I1..I13 then C1..C26
    num_pos = list(range(Fn))
    pad_row = field_off                      # padding_idx = 0 of every table
    NB = 4
    host_batches, rows_l, dense_l = [], [], []
    col_kind = [2] * Fn + [1] * F + [3]
    col_slot = list(range(Fn)) + list(range(F)) + [0]
    for i in range(NB):
        ids = make_ids(B, F, V, args.ids, 1000 * rank + i)
        dx = torch.rand(B, Fn, generator=g, dtype=torch.float64)
        lab = (torch.rand(B, 1, generator=g) < 0.5).double()
        M = torch.cat([dx, torch.from_numpy(ids).double(), lab], 1).contiguous().pin_memory()
        host_batches.append(M)
        r, d, _ = ops.split_batch(M.to(dev), col_kind, col_slot, field_off, F, Fn)
        rows_l.append(r)
        dense_l.append(d)
    dE = (torch.randn(B, Ft, D, generator=g) * 1e-3).to(dev)
    d_fm = (torch.randn(B, generator=g) * 1e-3).to(dev)
    d_lr = d_fm.clone()
    # dense gradients of every fused parameter live in ONE allocation -> one zero-fill launch
    sizes = [R * D, R, Fn * D, Fn, 1]
    offs = [0]
    for n in sizes:
        offs.append((offs[-1] + n + 3) // 4 * 4)
    gbuf = torch.empty(offs[-1], device=dev)
    g_table = gbuf[offs[0]:offs[0] + R * D].view(R, D)
    g_table_lr = gbuf[offs[1]:offs[1] + R]
    g_dense_w = gbuf[offs[2]:offs[2] + Fn * D].view(Fn, D)
    g_dense_w_lr = gbuf[offs[3]:offs[3] + Fn]
    g_bias = gbuf[offs[4]:offs[4] + 1]

    def fwd(rows, dx):
        return ops.embed_fm_fwd(table, table_lr, rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias)

    def zero():
        gbuf.zero_()

    def bwd(rows, dx, E, S):
        ops.embed_fm_bwd(table, rows, cat_pos, pad_row, dx, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                         g_table, g_table_lr, g_dense_w, g_dense_w_lr, g_bias, D, R)

    def step(i, evs=None):
        rows, dx = rows_l[i % NB], dense_l[i % NB]
        if evs: evs[0].record()
        E, S, fm, lr = fwd(rows, dx)
        if evs: evs[1].record()
        zero()
        if evs: evs[2].record()
        bwd(rows, dx, E, S)
        if evs: evs[3].record()
        return fm, lr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    for i in range(W):
        step(i)
    # ---- value: device-resident, CUDA events, max over ranks ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_beg.record()
    for i in range(K):
        step(i, evs[i])
    t_end.record()
    barrier()
    ms_total = t_beg.elapsed_time(t_end)
    t_f = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    t_z = sum(e[1].elapsed_time(e[2]) for e in evs) / K
    t_b = sum(e[2].elapsed_time(e[3]) for e in evs) / K

    # ---- e2e: pinned HOST batch -> H2D -> (split) -> step -> logits back to host --------------------
    # Every step's input crosses PCIe inside the timed region.  The loader-side prefetch is the usual
    # one: batch i+1 is copied on a copy stream while batch i computes (two device buffers, events
    # both ways); the logits go back on a third stream.
    #   mode "f64":    the reference loader's float64 [B,40] batch matrix (h5_dataloader.py:46) + rbx_split_batch_f64
    #   mode "packed": recbox_b200.loader.PackedDataLoader's blocks (int32 rows [B,F] + fp32 dense [B,Fn], converted
    #                  once at load time) -- the kernels read the copied blocks directly
    main = torch.cuda.current_stream()
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
    out_host = [torch.empty(B, dtype=torch.float32).pin_memory() for _ in range(2)]
    logit_dev = [torch.empty(B, dtype=torch.float32, device=dev) for _ in range(2)]
    packed_host = [(r.cpu().pin_memory(), d.cpu().pin_memory()) for r, d in zip(rows_l, dense_l)]

    def e2e_measure(mode):
        if mode == "f64":
            dev_in = [(torch.empty_like(host_batches[0], device=dev),) for _ in range(2)]
            src = [(m,) for m in host_batches]
        else:
            dev_in = [(torch.empty_like(rows_l[0]), torch.empty_like(dense_l[0])) for _ in range(2)]
            src = packed_host
        ev_in = [torch.cuda.Event() for _ in range(2)]      # batch landed in dev_in[j]
        ev_free = [torch.cuda.Event() for _ in range(2)]    # dev_in[j] consumed
        ev_out = [torch.cuda.Event() for _ in range(2)]     # logits of slot j computed
        ev_read = [torch.cuda.Event() for _ in range(2)]    # logits of slot j copied to the host

        def copy_in(i):
            j = i % 2
            with torch.cuda.stream(h2d):
                h2d.wait_event(ev_free[j])
                for d_, s_ in zip(dev_in[j], src[i % NB]):
                    d_.copy_(s_, non_blocking=True)
                ev_in[j].record(h2d)

        def compute(i):
            j = i % 2
            main.wait_event(ev_in[j])
            if mode == "f64":
                rows, dx, lab = ops.split_batch(dev_in[j][0], col_kind, col_slot, field_off, F, Fn)
                ev_free[j].record(main)
            else:
                rows, dx = dev_in[j]
            E, S, fm, lr = fwd(rows, dx)
            main.wait_event(ev_read[j])                      # the previous logits of this slot left the device
            torch.add(fm, lr, out=logit_dev[j])
            ev_out[j].record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev_out[j])
                out_host[j].copy_(logit_dev[j], non_blocking=True)
                ev_read[j].record(d2h)
            zero()
            bwd(rows, dx, E, S)
            if mode != "f64":
                ev_free[j].record(main)                      # the backward still reads the copied blocks

        def run(n):
            copy_in(0)
            for i in range(n):
                if i + 1 < n:
                    copy_in(i + 1)
                compute(i)
            main.wait_stream(d2h)
        for e in ev_free + ev_read:
            e.record(main)
        e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        run(W)
        barrier()
        e_beg.record()
        run(K)
        e_end.record()
        barrier()
        return e_beg.elapsed_time(e_end), sum(t.numel() * t.element_size() for t in src[0])

    ms_e2e, e2e_bytes, ms_e2e_packed, e2e_packed_bytes = 0.0, 0, 0.0, 0
    if not args.no_e2e:
        ms_e2e, e2e_bytes = e2e_measure("f64")
        ms_e2e_packed, e2e_packed_bytes = e2e_measure("packed")
    sampler.stop_flag = True
    sampler.join(timeout=2)

    times = torch.tensor([ms_total, ms_e2e, t_f, t_z, t_b, ms_e2e_packed], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, t_f, t_z, t_b, ms_e2e_packed = times.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        bf, bb = bytes_fwd(F, Fn, D) * B, bytes_bwd(F, Fn, D) * B
        kern = {"embed_fm_fwd": {"ms": t_f, "alg_bytes": bf, "gbs": bf / t_f / 1e6},
                "grad_zero_fill(memset)": {"ms": t_z, "alg_bytes": (R * D + R) * 4, "gbs": (R * D + R) * 4 / t_z / 1e6},
                "embed_fm_bwd(+dense_w_bwd)": {"ms": t_b, "alg_bytes": bb, "gbs": bb / t_b / 1e6}}
        dom = "embed_fm_bwd(+dense_w_bwd)" if t_b >= t_f else "embed_fm_fwd"
        ach = kern[dom]["gbs"]
        line = {
            "metric": "samples/sec on Criteo-shaped synthetic (embedding + FM hot path, fwd+bwd)",
            "value": B * world * K / (ms_total * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.ids, world),
            "e2e": None if args.no_e2e else {"value": B * world * K / (ms_e2e * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": B * 4, "ms_per_step": ms_e2e / K,
                    "input": "reference loader's float64 [B,40] batch matrix (h5_dataloader.py:46), pinned"},
            "e2e_packed": None if args.no_e2e else {"value": B * world * K / (ms_e2e_packed * 1e-3), "unit": "samples/s",
                    "h2d_bytes_per_step": e2e_packed_bytes, "d2h_bytes_per_step": B * 4, "ms_per_step": ms_e2e_packed / K,
                    "input": "recbox_b200.loader.PackedDataLoader blocks (int32 rows + fp32 dense, converted once at load), pinned"},
            "gpu_launches": 2 * K, "library_launches": K,   # ours: k_embed_fm_fwd + k_embed_fm_bwd; torch fill zeroes the grads
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic(dom) if args.ids == "uniform" else None, "peak_source": peak_src,
                         "pair_achieved": (bf + bb) / (t_f + t_b) / 1e6, "pair_frac": (bf + bb) / (t_f + t_b) / 1e6 / peak},
            "kernels": kern, "l2_persisting_bytes": l2_persist,
            "clocks": sampler.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            sps, dt = run_cpu(3, 1, B, args.ids)
            line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "3 steps of the full B=%d batch through oracle/recbox_oracle.py (%.0f ms/step)" % (B, dt * 1e3)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# B200 arm, BASELINE configs[3]: 100M-row table row-sharded over the ranks (weak scaling: B per GPU fixed)
# ------------------------------------------------------------------------------------------------
def main_sharded(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    from recbox_b200 import ops, sharded
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, F, Fn, D, V = CFG["B"], CFG["F"], CFG["Fn"], args.dim, args.rows_per_field
    Ft, R = F + Fn, F * V
    if args.shard_mode == "auto":       # one GPU: the local fused kernels; several: the streamed exchange
        args.shard_mode = "peer" if world == 1 else "stream"
    if args.shard_layout == "auto":     # ROW+LR wins on multi-GB tables at every N (profiles/r1_sharded_runs.jsonl, r2e / r1w)
        args.shard_layout = "rowlr" if (args.shard_mode in ("peer", "stream") and not args.no_lr and D in (4, 8, 16)) else "split"
    sh = sharded.ShardedEmbeddingFM(R, D, mode=args.shard_mode, device=dev, max_ids=B * F, alloc=args.peer_alloc, with_lr=not args.no_lr,
                                    layout=args.shard_layout, chunks=args.shard_chunks, slack=1.25)
    gen = torch.Generator(device=dev).manual_seed(20240 + 4 + rank)
    sh.table.normal_(0, 0.01, generator=gen)
    sh.table_lr.normal_(0, 0.01, generator=gen)
    g = torch.Generator().manual_seed(20240 + 4)
    dense_w = (torch.randn(Fn, D, generator=g) * 0.1).to(dev)
    dense_w_lr = (torch.randn(Fn, generator=g) * 0.1).to(dev)
    bias = torch.zeros(1, device=dev)
    field_off = [f * V for f in range(F)]
    cat_pos, num_pos, pad_row = list(range(Fn, Ft)), list(range(Fn)), field_off
    NB = 4
    rows_l, dense_l = [], []
    for i in range(NB):
        ids = torch.from_numpy(make_ids(B, F, V, args.ids, 1000 * rank + i))
        rows_l.append((ids + torch.tensor(field_off)[None]).to(torch.int32).to(dev))
        dense_l.append(torch.rand(B, Fn, generator=g).to(dev))
    dE = (torch.randn(B, Ft, D, generator=g) * 1e-3).to(dev)
    d_fm = (torch.randn(B, generator=g) * 1e-3).to(dev)
    d_lr = d_fm.clone()
    gw, gw1, gb = torch.zeros(Fn, D, device=dev), torch.zeros(Fn, device=dev), torch.zeros(1, device=dev)
    sh.barrier()

    def step(i, evs=None):
        rows, dx = rows_l[i % NB], dense_l[i % NB]
        if evs: evs[0].record()
        E, S, fm, lr = sh.forward(rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias)
        if evs: evs[1].record()
        sh.backward(rows, cat_pos, pad_row, dx, dense_w, num_pos, E, S, dE, d_fm, d_lr, gw, gw1, gb)
        if evs: evs[2].record()
        if args.shard_mode != "stream":
            sh.device_barrier()        # remote reductions of this step have landed in every owner's shard
        if evs: evs[3].record()
        return fm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    for i in range(W):
        step(i)
    sh.check_overflow()
    sampler = ClockSampler(local_rank)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for i in range(K):
        step(i, evs[i])
    t1.record()
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_total = t0.elapsed_time(t1)
    t_f = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    t_b = sum(e[1].elapsed_time(e[2]) for e in evs) / K
    t_s = sum(e[2].elapsed_time(e[3]) for e in evs) / K
    phases = None
    if args.shard_mode == "stream":       # a second, separately timed pass with an event after every phase (rank 0's view)
        sh.profile = []
        for i in range(K):
            step(i)
        torch.cuda.synchronize()
        phases = {k: v / K for k, v in sh.phase_times(sh.profile).items()}
        sh.profile = None
        barrier()
    times = torch.tensor([ms_total, t_f, t_b, t_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, t_f, t_b, t_s = times.tolist()
    if rank == 0:
        peak, peak_src = peaks()
        bf, bb = bytes_fwd(F, Fn, D) * B, bytes_bwd(F, Fn, D) * B
        remote = (world - 1) / world
        nv_f = remote * B * F * (4 * D + 4)               # row + lr value read from peers
        nv_b = remote * B * F * (4 * D + 4)               # row grad + lr grad reduced into peers
        line = {
            "metric": "samples/sec on Criteo-shaped synthetic (embedding + FM hot path, fwd+bwd)",
            "value": B * world * K / (ms_total * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: DeepFM hot path, %d-row fused table (26 x %d) row-sharded over %d GPU(s), "
                                   "D=%d, B=65536 per GPU; mode=%s%s" % (R, V, world, D, args.shard_mode,
                                                                            (", layout=%s" % args.shard_layout if args.shard_layout != "split" else "") +
                                                                            (", %d pipelined sample ranges" % args.shard_chunks if args.shard_chunks > 1 else "")),
                       "global_batch": B * world, "ids": args.ids, "batches_rotated": NB,
                       "l2": "table shard (%.1f GB) and E/dE streams exceed the 126 MB L2" % (sh.cap * D * 4 / 1e9),
                       "parallelism": "dp%d batch shards + row-sharded table (r %% %d), exchange inside the fused kernels over NVLink" % (world, world)
                       if args.shard_mode == "peer" else
                       ("dp%d batch shards + row-sharded table (r %% %d); streamed exchange: ids / rows / row gradients cross NVLink as contiguous runs "
                        "written by the kernels, flag barriers over peer memory" % (world, world) if args.shard_mode == "stream"
                        else "dp%d + row-sharded table, NCCL all_to_all" % world)},
            "gpu_launches": {"peer": 3, "stream": 9}.get(args.shard_mode, 12) * K,
            "roofline": {"bound": "hbm" if world == 1 else "nvlink", "kernel": "embed_fm_bwd_sharded" if t_b >= t_f else "embed_fm_fwd_sharded",
                         "achieved": max(bf / t_f, bb / t_b) / 1e6 if world == 1 else (nv_b / t_b if t_b >= t_f else nv_f / t_f) / 1e6,
                         "peak": peak if world == 1 else 770.0, "unit": "GB/s",
                         "frac": (max(bf / t_f, bb / t_b) / 1e6 / peak) if world == 1 else ((nv_b / t_b if t_b >= t_f else nv_f / t_f) / 1e6 / 770.0),
                         "traffic": None,
                         "peak_source": peak_src if world == 1 else "B200_PROFILING.md measured peer copy 770 GB/s per direction"},
            "kernels": {"fwd": {"ms": t_f, "hbm_alg_gbs": bf / t_f / 1e6, "nvlink_gbs": nv_f / t_f / 1e6},
                        "bwd": {"ms": t_b, "hbm_alg_gbs": bb / t_b / 1e6, "nvlink_gbs": nv_b / t_b / 1e6},
                        "step_barrier": {"ms": t_s}},
            "clocks": sampler.summary(),
        }
        if phases is not None:
            line["phases_ms_rank0"] = phases
        print(json.dumps(line))
    sh.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "sharded"],
                    help="cfg2: BASELINE configs[1], replicated 1M-row table (default, the metric's config); "
                         "sharded: configs[3], 100M-row table row-sharded over the ranks")
    ap.add_argument("--shard-mode", default="auto", choices=["auto", "stream", "push", "peer", "a2a"])
    ap.add_argument("--peer-alloc", default="symm", choices=["ipc", "symm"])
    ap.add_argument("--shard-layout", default="auto", choices=["auto", "split", "rowlr", "rowpad"],
                    help="rowlr: embedding row + first-order weight in one physical row (one NVLink request per slot)")
    ap.add_argument("--shard-chunks", type=int, default=1, help="stream mode: sample ranges pipelined on their own streams")
    ap.add_argument("--no-lr", action="store_true", help="sharded workload without the first-order (LR) table")
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--rows-per-field", type=int, default=3846154)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    elif args.workload == "sharded":
        main_sharded(args, rank, world, local_rank)
    else:
        main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
