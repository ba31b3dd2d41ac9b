#!/usr/bin/env python
"""bench.py -- samples/sec of the RecBox embedding + FM hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--ids uniform|zipf]
                  [--workload cfg2|sharded|dssm|sasrec]

A "step" is one pass of the hot path over one Criteo-shaped synthetic batch (configs[1]: DeepFM,
26 categorical + 13 numeric fields, 26 x 38 462 rows = 1 000 012-row fused table, D = 16,
B = 65 536 per GPU):
    forward  : fused multi-slot gather + numeric Linear(1,D) + FM product_sum + LR  -> E, S, fm, lr
    backward : zero the fused dense gradient buffer (our kernel, side stream), then scatter-add dE + d_fm*(S-e) (and d_lr)
               into it, plus the numeric-slot / bias batch reductions
The dense MLP tail (a13) has its own numbers (tools/train_step_bench.py); here its input gradient dE is a
device-resident stand-in.
  value   the step with inputs resident in HBM, replayed from CUDA graphs, EXACTLY K steps between two CUDA events
  e2e     the same step through the LAYER API (FeatureEmbedding + FactorizationMachine modules, autograd), fed every step from
          pinned host memory (packed uint16 ids + fp32 dense), logits read back to the host -- copies inside the timed region
  sharded BASELINE configs[3] (100 M-row table row-sharded over the ranks) at the same N, with an in-run parity check
`--impl reference` times the reference's OWN FeatureEmbedding + FactorizationMachine (baseline/_ref, copied unmodified by
baseline/fetch_ref.py; the oracle port only when that copy is absent) on all host cores.  Under torchrun (N > 1) the ranks are
replicas that train ONE model: each runs the path on its own batch shard and the fused gradient buffer is all-reduced inside the
timed step (our in-switch kernel or NCCL).
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


def _teardown(dist):
    """Process-group teardown that cannot outlive the bench line: the JSON is already printed and flushed; if NCCL's teardown
    stalls (seen once with CUDA graphs still alive, run r2t) the process leaves with status 0 after 30 s."""
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Timer(30.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    try:
        dist.destroy_process_group()
    finally:
        t.cancel()


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
# reference arm / cpu_baseline: the reference's OWN modules (baseline/_ref, copied unmodified by baseline/fetch_ref.py)
# on the host cores; the oracle's restatement (kind "port") only when that copy is absent
# ------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def criteo_feature_map(cls, name="bench_criteo"):
    """FeatureMap of BASELINE configs[1]: I1..I13 numeric, C1..C26 categorical (38 462 rows each, padding_idx 0), built the
    way SURVEY.md Appendix A drives the reference's FeatureMap by hand; `cls` is the reference's class or ours."""
    fm = cls(name, ROOT)
    for n in range(CFG["Fn"]):
        fm.features["I%d" % (n + 1)] = {"source": "", "type": "numeric"}
    for f in range(CFG["F"]):
        fm.features["C%d" % (f + 1)] = {"source": "", "type": "categorical", "vocab_size": CFG["V"], "padding_idx": 0}
    fm.labels = ["label"]
    fm.num_fields = fm.get_num_fields()
    fm.set_column_index()
    fm.default_emb_dim = CFG["D"]
    return fm


def host_batch_matrix(B, ids_kind, seed, g):
    """The reference loader's batch: one float64 [B, 40] matrix, features in FeatureMap order then the label
    (h5_dataloader.py:36-47)."""
    ids = make_ids(B, CFG["F"], CFG["V"], ids_kind, seed)
    dx = torch.rand(B, CFG["Fn"], generator=g, dtype=torch.float64)
    lab = (torch.rand(B, 1, generator=g) < 0.5).double()
    return torch.cat([dx, torch.from_numpy(ids).double(), lab], 1).contiguous()


def ref_step_factory(B, ids_kind, seed):
    """One pass of the hot path through the reference's FeatureEmbedding + FactorizationMachine (unmodified source under
    baseline/_ref): forward E and the FM logit, backward with the same upstream gradients the B200 arm uses."""
    os.environ["RECBOX_REFERENCE"] = REF_DIR
    from oracle import ref_shim                      # import aliases only (fuxictr.* -> recbox.ranking.*, stub h5py / faiss)
    L = ref_shim.install()
    from recbox.ranking.features import FeatureMap as RefFeatureMap
    fm = criteo_feature_map(RefFeatureMap)
    torch.manual_seed(seed)
    emb, fml = L.FeatureEmbedding(fm, CFG["D"]), L.FactorizationMachine(fm)
    g = torch.Generator().manual_seed(seed)
    M = host_batch_matrix(B, ids_kind, seed, g)
    dE = torch.randn(B, CFG["F"] + CFG["Fn"], CFG["D"], generator=g) * 1e-3
    d_out = torch.randn(B, 1, generator=g) * 1e-3
    params = list(emb.parameters()) + list(fml.parameters())
    cols = {n: fm.get_column_index(n) for n in fm.features}

    def step():
        for p in params:
            p.grad = None                                          # optimizer.zero_grad()
        X = {n: M[:, c] for n, c in cols.items()}                  # RankingModel.get_inputs (ranking_model.py:106-116)
        E = emb(X)
        y = fml(X, E)
        torch.autograd.backward([E, y], [dE, d_out])               # dense embedding grads
        return y
    return step


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
    """-> (samples/s, s per step, kind, description): the reference's own modules when baseline/_ref is present."""
    torch.set_num_threads(os.cpu_count() or 1)
    kind, what = "port", "oracle/recbox_oracle.py (CPU restatement; baseline/_ref absent)"
    step = None
    if os.path.isdir(os.path.join(REF_DIR, "recbox", "ranking")):
        try:
            step = ref_step_factory(B, ids_kind, 20242)
            kind, what = "reference", "the reference's own FeatureEmbedding + FactorizationMachine (baseline/_ref, unmodified)"
        except Exception as e:             # say so instead of silently timing something else
            what = "oracle/recbox_oracle.py (CPU restatement; baseline/_ref failed to import: %s)" % str(e)[:120]
    if step is None:
        step = cpu_step_factory(B, ids_kind, 20242)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return B / dt, dt, kind, what


def main_reference(args, rank, world):
    if rank != 0:
        return
    B = CFG["B"]
    steps, warmup = max(1, min(args.steps, 20)), max(1, min(args.warmup, 3))
    sps, dt, kind, what = run_cpu(steps, warmup, B, args.ids)
    cores = torch.get_num_threads()
    cfg = workload_config(args.ids, 1)
    cfg["parallelism"] = "one host process, %d threads (the reference's CPU path has no multi-device mode)" % cores
    line = {
        "impl": "reference", "metric": "samples/sec on Criteo-shaped synthetic (embedding + FM hot path, fwd+bwd)",
        "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": "%d step(s) of the full B=%d batch through %s (torch %s CPU ops)" % (steps, B, what, torch.__version__)},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(ids_kind, world, saved_e=False):
    par = ("1 GPU" if world == 1 else
           "dp%d replicas of the 1M-row table, each on its own batch shard; the fused dense gradient buffer (68 MB) is "
           "all-reduced over NVLink / NVSwitch INSIDE the timed step (SURVEY 8e 'replicas only'; kernels.all_reduce.path says how)" % world)
    e_src = "read from the forward's saved E" if saved_e else "re-read from the table, not from the forward's E stream"
    return {"workload": "BASELINE configs[1] DeepFM hot path: 26 cat + 13 dense, 26x38462 = 1000012-row fused table, "
                        "D=16, B=65536 per GPU; fused gather+FM+LR fwd, dense-grad zero (side stream, under the fwd) + scatter-add bwd "
                        "(the FM term's e " + e_src + "); MLP tail outside the path",
            "global_batch": CFG["B"] * world, "ids": ids_kind, "batches_rotated": 4,
            "l2": "working set per step (E 163 MB + dE 163 MB + table/grad 136 MB) exceeds the 126 MB L2; 4 id batches rotate",
            "parallelism": par}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def main_b200(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU path; use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from recbox_b200 import layers, loader, ops
    from recbox_b200.features import FeatureMap
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
        M = host_batch_matrix(B, args.ids, 1000 * rank + i, g).pin_memory()
        host_batches.append(M)
        r, d, _ = ops.split_batch(M.to(dev), col_kind, col_slot, field_off, F, Fn)
        rows_l.append(r)
        dense_l.append(d)
    dE = (torch.randn(B, Ft, D, generator=g) * 1e-3).to(dev)
    d_fm = (torch.randn(B, generator=g) * 1e-3).to(dev)
    d_lr = d_fm.clone()
    # dense gradients of every fused parameter live in ONE allocation -> one zero-fill launch, one all-reduce
    sizes = [R * D, R, Fn * D, Fn, 1]
    offs = [0]
    for n in sizes:
        offs.append((offs[-1] + n + 3) // 4 * 4)
    # replicas (N > 1): the buffer lives in symmetric multicast memory and is summed by the in-switch all-reduce kernel
    # (recbox_b200.replica.ReplicaReducer -> rbx_nvls_allreduce_f32); --allreduce nccl keeps torch.distributed's collective
    reducer = None
    if world > 1 and (args.allreduce == "nvls" or (args.allreduce == "auto" and world >= 8)):
        from recbox_b200 import replica
        reducer = replica.ReplicaReducer(offs[-1], dev)
        gbuf = reducer.buffer
    else:
        gbuf = torch.zeros(offs[-1], device=dev)
    g_table = gbuf[offs[0]:offs[0] + R * D].view(R, D)
    g_table_lr = gbuf[offs[1]:offs[1] + R]
    g_dense_w = gbuf[offs[2]:offs[2] + Fn * D].view(Fn, D)
    g_dense_w_lr = gbuf[offs[3]:offs[3] + Fn]
    g_bias = gbuf[offs[4]:offs[4] + 1]
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    ev_fork, ev_zero = torch.cuda.Event(), torch.cuda.Event()

    def fwd(rows, dx):
        return ops.embed_fm_fwd(table, table_lr, rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias)

    def zero_async():
        """optimizer.zero_grad() of the fused gradient buffer on the side stream: forked from the step's stream (so it
        follows the last reader of the previous step's gradients -- the backward / the all-reduce) and run under the forward.
        --zero inline issues it on the step's own stream in front of the forward instead."""
        if args.zero == "inline":
            ops.zero_(gbuf)
            ev_zero.record(torch.cuda.current_stream())
            return
        ev_fork.record(torch.cuda.current_stream())
        side.wait_event(ev_fork)
        with torch.cuda.stream(side):
            ops.zero_(gbuf)
            ev_zero.record(side)

    def bwd(rows, dx, E, S):
        if not args.bwd_saved_e:           # e re-read from the (mostly L2-resident) table instead of the forward's E stream:
            E = None                       # same algorithmic bytes (F * 4D either way), fewer DRAM bytes (r2ba: 355 vs 351 M samples/s)
        ops.embed_fm_bwd(table, rows, cat_pos, pad_row, dx, dense_w, num_pos, E, S, dE, d_fm, d_lr,
                         g_table, g_table_lr, g_dense_w, g_dense_w_lr, g_bias, D, R)

    def kernels_step(i, evs=None):
        """forward + zero-fill + backward on batch i (three launches of ours)."""
        rows, dx = rows_l[i % NB], dense_l[i % NB]
        zero_async()
        if evs: evs[0].record()
        E, S, fm, lr = fwd(rows, dx)
        if evs: evs[1].record()
        torch.cuda.current_stream().wait_event(ev_zero)
        if evs: evs[2].record()
        bwd(rows, dx, E, S)
        if evs: evs[3].record()
        return fm, lr

    # the three launches of each of the NB rotating batches as a CUDA graph: the step is ~0.19 ms, so the gaps between
    # eagerly issued launches (ctypes + event calls) are a measurable part of it
    graphed = None
    if not args.eager:
        from recbox_b200 import graphs
        try:
            graphed = [graphs.GraphedStep(lambda i=i: kernels_step(i), warmup=2) for i in range(NB)]
        except Exception as e:
            sys.stderr.write("bench: GraphedStep failed (%r); the device-resident loop issues its launches eagerly\n" % (e,))
            graphed = None
            torch.cuda.synchronize()

    def step(i, evs=None):
        if evs is None and graphed is not None:
            out = graphed[i % NB]()
        else:
            out = kernels_step(i, evs)
        if world > 1:                          # replicas train ONE model: dense gradient exchange inside the step
            reducer.all_reduce() if reducer is not None else dist.all_reduce(gbuf)
        if evs: evs[4].record()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    for i in range(W):
        step(i)
    # ---- value: device-resident, CUDA events around EXACTLY K steps, max over ranks ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_beg.record()
    for i in range(K):
        step(i)
    t_end.record()
    barrier()
    ms_total = t_beg.elapsed_time(t_end)
    # ---- per-kernel times: a second pass over the same K steps, issued eagerly with an event after every launch ----
    # (the GPU first spins for ~0.3 ms per step, so the host is ahead of it for the whole pass and the gaps between the
    # events are device time, not the time the host needs to issue a launch)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    for i in range(3):
        step(i, evs[i])
    torch.cuda.synchronize()
    torch.cuda._sleep(int(0.3e-3 * K * 1.9e9))
    for i in range(K):
        step(i, evs[i])
    barrier()
    t_f = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    t_w = sum(e[1].elapsed_time(e[2]) for e in evs) / K          # backward waiting for the zero-fill (0 when it hid)
    t_b = sum(e[2].elapsed_time(e[3]) for e in evs) / K
    t_ar = sum(e[3].elapsed_time(e[4]) for e in evs) / K
    ms_evented = evs[0][0].elapsed_time(evs[-1][4]) / K          # the evented pass's own step time (its kernels sum below it)
    # ---- the same two kernels in the regime `value` is measured in (graph replay, no events between the launches): K replays
    # of a graph that holds only forward + zero-fill; the backward's share is the step minus that
    t_f_graph = t_b_graph = None
    if graphed is not None and world == 1:
        def fwd_only(i):
            zero_async()
            out = fwd(rows_l[i % NB], dense_l[i % NB])
            torch.cuda.current_stream().wait_event(ev_zero)
            return out
        try:
            gf = [graphs.GraphedStep(lambda i=i: fwd_only(i), warmup=1) for i in range(NB)]
            for i in range(3):
                gf[i % NB]()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            f0.record()
            for i in range(K):
                gf[i % NB]()
            f1.record()
            torch.cuda.synchronize()
            t_f_graph = f0.elapsed_time(f1) / K
            t_b_graph = ms_total / K - t_f_graph
            del gf
        except Exception as e:
            sys.stderr.write("bench: forward-only graph failed (%r)\n" % (e,))
            torch.cuda.synchronize()
    # the zero-fill kernel alone (own events on its stream, separate pass so that it does not perturb the timed region)
    z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        z0.record(side)
        for _ in range(10):
            ops.zero_(gbuf)
        z1.record(side)
    torch.cuda.synchronize()
    t_z = z0.elapsed_time(z1) / 10

    # ---- e2e: pinned HOST batch -> H2D -> step -> logits back to host, every step, inside the timed region ----------
    # Through the public layer API a RecBox model calls (FeatureEmbedding + FactorizationMachine modules, autograd
    # backward with the MLP tail's stand-in gradient), fed by
    #   "packed":  recbox_b200.loader.PackedDataLoader's blocks -- uint16 ids [B,26] (every vocabulary < 65 536) + fp32
    #              dense [B,13] + fp32 label, converted once at load time: 108 B / sample   <- `e2e`
    #   "f64":     the reference loader's float64 [B,40] batch matrix (h5_dataloader.py:46): 320 B / sample  <- `e2e_f64`
    # and once below the layer API, straight on the ops (int32 packed blocks), to show what the Python layer costs.
    # The loader-side prefetch is the usual one: batch i+1 is copied on a copy stream while batch i computes (two device
    # buffers, events both ways); the logits go back on a third stream.
    fmap = criteo_feature_map(FeatureMap)
    torch.manual_seed(20240 + 2)
    emb = layers.FeatureEmbedding(fmap, D).to(dev)
    fml = layers.FactorizationMachine(fmap).to(dev)
    params = list(emb.parameters()) + list(fml.parameters())
    d_out = d_fm.view(-1, 1)

    class _Model(object):                                        # what layers.get_inputs needs of a RankingModel
        feature_map, device = fmap, dev
    ds = loader.PackedDataset(fmap, torch.cat(host_batches, 0))  # compact: uint16 ids
    host_packed = [ds.batch(i * B, (i + 1) * B) for i in range(NB)]
    h2d, d2h = torch.cuda.Stream(), torch.cuda.Stream()
    out_host = [torch.empty(B, 1, dtype=torch.float32).pin_memory() for _ in range(2)]
    logit_dev = [torch.empty(B, 1, dtype=torch.float32, device=dev) for _ in range(2)]
    packed32_host = [(r.cpu().pin_memory(), d.cpu().pin_memory()) for r, d in zip(rows_l, dense_l)]

    issue = {"packed": "eager", "f64": "eager", "ops": "eager"}

    def e2e_measure(mode):
        if mode == "f64":
            dev_in = [torch.empty_like(host_batches[0], device=dev) for _ in range(2)]
            nbytes = host_batches[0].numel() * 8
        elif mode == "packed":
            dev_in = [host_packed[0].to(dev) for _ in range(2)]
            nbytes = host_packed[0].nbytes
        else:
            dev_in = [(torch.empty_like(rows_l[0]), torch.empty_like(dense_l[0])) for _ in range(2)]
            nbytes = sum(t.numel() * t.element_size() for t in packed32_host[0])
        ev_in = [torch.cuda.Event() for _ in range(2)]      # batch landed in dev_in[j]
        ev_free = [torch.cuda.Event() for _ in range(2)]    # dev_in[j] consumed
        ev_out = [torch.cuda.Event() for _ in range(2)]     # logits of slot j computed
        ev_read = [torch.cuda.Event() for _ in range(2)]    # logits of slot j copied to the host

        def copy_in(i):
            j = i % 2
            with torch.cuda.stream(h2d):
                h2d.wait_event(ev_free[j])
                if mode == "f64":
                    dev_in[j].copy_(host_batches[i % NB], non_blocking=True)
                elif mode == "packed":
                    host_packed[i % NB].copy_into(dev_in[j])
                else:
                    for d_, s_ in zip(dev_in[j], packed32_host[i % NB]):
                        d_.copy_(s_, non_blocking=True)
                ev_in[j].record(h2d)

        # the layer-API step of slot j as a CUDA graph (recbox_b200.graphs.GraphedStep): ~80 per-feature Parameters cross
        # autograd per step (1.2 ms of host work measured eagerly, profiles/r2_layer_overhead.txt) although two kernels do
        # the work; the replay re-issues the captured launches with no Python in between.  "eager" keeps the plain calls.
        graphed = [None, None]

        def layer_step(j):
            for p in params:
                p.grad = None                                # optimizer.zero_grad() (ranking_model.py:192)
            X = layers.get_inputs(_Model, dev_in[j])         # PackedColumns / PackedInputs views, no copies
            E = emb(X)
            y = fml(X, E)
            logit_dev[j].copy_(y.detach())
            torch.autograd.backward([E, y], [dE, d_out])

        if mode in ("packed", "f64") and not args.eager:
            from recbox_b200 import graphs
            try:
                for j in range(2):
                    graphed[j] = graphs.GraphedStep(lambda j=j: layer_step(j), warmup=2)
                issue[mode] = "cuda-graph replay"
            except Exception as e:
                sys.stderr.write("bench: GraphedStep failed (%r); the layer-API e2e leg runs eagerly\n" % (e,))
                graphed = [None, None]
                torch.cuda.synchronize()

        def compute(i):
            j = i % 2
            main.wait_event(ev_in[j])
            main.wait_event(ev_read[j])                      # the previous logits of this slot left the device
            if mode == "ops":
                rows, dx = dev_in[j]
                zero_async()
                E, S, fm, lr = fwd(rows, dx)
                torch.add(fm.view(-1, 1), lr.view(-1, 1), out=logit_dev[j])
                main.wait_event(ev_zero)
                bwd(rows, dx, E, S)
            elif graphed[j] is not None:
                graphed[j]()
            else:
                layer_step(j)
            ev_out[j].record(main)
            with torch.cuda.stream(d2h):
                d2h.wait_event(ev_out[j])
                out_host[j].copy_(logit_dev[j], non_blocking=True)
                ev_read[j].record(d2h)
            if world > 1:                                    # the replicas' gradient exchange belongs to the step; issued eagerly
                if mode == "ops":                            # after the replay (a graph that holds NCCL kernels hung the
                    reducer.all_reduce() if reducer is not None else dist.all_reduce(gbuf)   # process-group teardown in run r2t)
                else:
                    layers.sync_replica_gradients(params, reducer=reducer)
            ev_free[j].record(main)                          # the backward still reads the copied blocks

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
        return e_beg.elapsed_time(e_end), nbytes

    e2e_ms = {}
    if not args.no_e2e:
        for mode in ("packed", "f64", "ops"):
            e2e_ms[mode] = e2e_measure(mode)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    names = ["packed", "f64", "ops"]
    times = torch.tensor([ms_total, t_f, t_w, t_b, t_ar, t_z] + [e2e_ms[m][0] if m in e2e_ms else 0.0 for m in names],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, t_f, t_w, t_b, t_ar, t_z = times.tolist()[:6]
    e2e_t = dict(zip(names, times.tolist()[6:]))

    sharded_line = None
    if not args.no_sharded:     # configs[3] at the same N, in the same invocation (driver-witnessed exchange + parity)
        try:
            sharded_line = run_sharded(args, rank, world, dev, min(K, 30), 3)
            if sharded_line is not None:
                for k in ("metric", "unit", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "n_gpus"):
                    sharded_line.pop(k, None)
        except Exception as e:    # the headline line must survive a failure of the secondary workload; the failure is reported
            sharded_line = {"error": repr(e)[:300]}

    if rank == 0:
        peak, peak_src = peaks()
        bf, bb = bytes_fwd(F, Fn, D) * B, bytes_bwd(F, Fn, D) * B
        kern = {"embed_fm_fwd": {"ms": t_f, "alg_bytes": bf, "gbs": bf / t_f / 1e6},
                "grad_zero_fill (rbx_zero_f32, side stream under the forward)": {"ms": t_z, "alg_bytes": offs[-1] * 4, "gbs": offs[-1] * 4 / t_z / 1e6,
                                                                                  "exposed_ms": t_w},
                "embed_fm_bwd(+dense_w_bwd)": {"ms": t_b, "alg_bytes": bb, "gbs": bb / t_b / 1e6}}
        if world > 1:
            kern["all_reduce(grad buffer)"] = {"ms": t_ar, "bytes": offs[-1] * 4, "path": reducer.path if reducer is not None else "NCCL all_reduce",
                                               "busbw_gbs": 2 * (world - 1) / world * offs[-1] * 4 / t_ar / 1e6}
        if t_f_graph is not None:
            kern["graph_regime"] = {
                "what": "the regime `value` is measured in (CUDA-graph replay, launches back to back, no events between them): "
                        "forward(+zero-fill under it) = K replays of a forward-only graph; backward = step - that",
                "embed_fm_fwd(+zero_fill)": {"ms": t_f_graph, "alg_bytes": bf + offs[-1] * 4, "gbs": (bf + offs[-1] * 4) / t_f_graph / 1e6},
                "embed_fm_bwd(+dense_w_bwd)": {"ms": t_b_graph, "alg_bytes": bb, "gbs": bb / t_b_graph / 1e6,
                                               "frac": bb / t_b_graph / 1e6 / peak},
                "step_gbs": (bf + bb + offs[-1] * 4) / (ms_total / K) / 1e6, "step_frac": (bf + bb + offs[-1] * 4) / (ms_total / K) / 1e6 / peak}
        kern["evented_pass_ms_per_step"] = ms_evented
        dom = "embed_fm_bwd(+dense_w_bwd)" if t_b >= t_f else "embed_fm_fwd"
        ach = kern[dom]["gbs"]
        how = "CUDA events around every launch of an eagerly issued pass over the same K steps (includes the event / launch gaps)"
        if t_f_graph is not None and dom == "embed_fm_bwd(+dense_w_bwd)":
            # the dominant kernel's time in the regime the step is timed in: both terms are CUDA-event times over K launches,
            # and forward + backward add up to the step by construction
            ach = kern["graph_regime"][dom]["gbs"]
            how = ("graph regime: step time - time of a forward-only graph, both by CUDA events over K replays "
                   "(the evented eager pass reads %.1f us for this kernel, event gaps included)" % (t_b * 1e3))

        def e2e_obj(mode, what):
            ms, nbytes = e2e_t[mode], e2e_ms[mode][1]
            return {"value": B * world * K / (ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": B * 4, "ms_per_step": ms / K, "input": what, "issue": issue[mode]}
        line = {
            "metric": "samples/sec on Criteo-shaped synthetic (embedding + FM hot path, fwd+bwd)",
            "value": B * world * K / (ms_total * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args.ids, world, args.bwd_saved_e),
            "e2e": None if args.no_e2e else e2e_obj("packed", "layer API (FeatureEmbedding + FactorizationMachine modules, autograd) fed by "
                    "recbox_b200.loader.PackedDataset blocks: uint16 ids + fp32 dense + fp32 label (108 B/sample), pinned"),
            "e2e_f64": None if args.no_e2e else e2e_obj("f64", "layer API fed by the reference loader's float64 [B,40] batch matrix "
                    "(h5_dataloader.py:46; 320 B/sample), pinned -- the compatibility number"),
            "e2e_ops": None if args.no_e2e else e2e_obj("ops", "below the layer API: recbox_b200.ops on int32 packed blocks (160 B/sample), pinned"),
            # ours per step: k_embed_fm_fwd + k_zero_f32 + k_embed_fm_bwd; library: the NCCL all-reduce at N > 1
            "gpu_launches": (3 + (1 if (reducer is not None and reducer.mc) else 0)) * K,
            "library_launches": (K if reducer is None else 2 * K) if world > 1 else 0,      # NCCL all-reduce, or the two barrier launches around ours
            "issue": "cuda-graph replay (one graph per rotating batch)" if graphed is not None else "eager",
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic(dom) if args.ids == "uniform" else None, "peak_source": peak_src, "how": how,
                         "pair_achieved": (bf + bb) / (t_f + t_b) / 1e6, "pair_frac": (bf + bb) / (t_f + t_b) / 1e6 / peak,
                         "step_achieved": (bf + bb + offs[-1] * 4) / (ms_total / K) / 1e6,
                         "step_frac": (bf + bb + offs[-1] * 4) / (ms_total / K) / 1e6 / peak},
            "kernels": kern, "l2_persisting_bytes": l2_persist,
            "clocks": sampler.summary(),
        }
        if sharded_line is not None:
            line["sharded"] = sharded_line
        if world == 1 and not args.no_cpu_baseline:
            sps, dt, kind, what = run_cpu(3, 1, B, args.ids)
            line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": "3 steps of the full B=%d batch through %s (%.0f ms/step)" % (B, what, dt * 1e3)}
        print(json.dumps(line))
    if world > 1:
        _teardown(dist)


# ------------------------------------------------------------------------------------------------
# B200 arm, BASELINE configs[3]: 100M-row table row-sharded over the ranks (weak scaling: B per GPU fixed)
# ------------------------------------------------------------------------------------------------
def sharded_parity(args, rank, world, dev):
    """In-run parity of the exchange at THIS world size: a small row-sharded table (same mode / layout as the timed run)
    against the single-table fused kernels (the oracle-pinned path, tests/test_kernels_gpu.py) on the same seeded inputs --
    gathered rows bit-exact, sums and gradients within 1e-5 (transitive parity: sharded == single-table == oracle).
    -> True / False over all ranks."""
    import torch.distributed as dist
    from recbox_b200 import ops, sharded
    F, Fn, D, V, B = 8, 3, args.dim, 997, 2048
    Ft, R = F + Fn, F * V
    g = torch.Generator().manual_seed(4242)                       # every rank builds the same global problem
    table = (torch.randn(R, D, generator=g) * 0.1).to(dev)
    table_lr = (torch.randn(R, generator=g) * 0.1).to(dev)
    dense_w = (torch.randn(Fn, D, generator=g) * 0.1).to(dev)
    dense_w_lr = (torch.randn(Fn, generator=g) * 0.1).to(dev)
    bias = torch.full((1,), 0.25, device=dev)
    field_off = [f * V for f in range(F)]
    ids = torch.randint(0, V, (B * world, F), generator=g)
    rows_all = (ids + torch.tensor(field_off)[None]).to(torch.int32).to(dev)
    dx_all = torch.rand(B * world, Fn, generator=g).to(dev)
    dE = torch.randn(B * world, Ft, D, generator=g).to(dev)
    d_fm = torch.randn(B * world, generator=g).to(dev)
    d_lr = torch.randn(B * world, generator=g).to(dev)
    cat_pos, num_pos, pad_row = list(range(Fn, Ft)), list(range(Fn)), field_off
    sl = slice(rank * B, (rank + 1) * B)
    ok = True
    sh = sharded.ShardedEmbeddingFM(R, D, mode=args.shard_mode, device=dev, max_ids=B * F, alloc=args.peer_alloc, layout=args.shard_layout,
                                    chunks=args.shard_chunks, slack=3.0)
    try:
        sh.load_global(table, table_lr)
        rows, dx = rows_all[sl].contiguous(), dx_all[sl].contiguous()
        E, S, fm, lr = sh.forward(rows, cat_pos, dx, dense_w, dense_w_lr, num_pos, bias)
        gw, gw1, gb = torch.zeros(Fn, D, device=dev), torch.zeros(Fn, device=dev), torch.zeros(1, device=dev)
        sh.zero_grad()
        sh.barrier()
        sh.backward(rows, cat_pos, pad_row, dx, dense_w, num_pos, E, S, dE[sl].contiguous(), d_fm[sl].contiguous(), d_lr[sl].contiguous(),
                    gw, gw1, gb)
        sh.barrier()
        sh.check_overflow()
        if world > 1:
            for t in (gw, gw1, gb):
                dist.all_reduce(t)
        gt, gt1 = sh.gather_global("g_table"), sh.gather_global("g_table_lr")
        Er, Sr, fmr, lrr = ops.embed_fm_fwd(table, table_lr, rows_all, cat_pos, dx_all, dense_w, dense_w_lr, num_pos, bias)
        rt, rt1 = torch.zeros_like(table), torch.zeros_like(table_lr)
        rw, rw1, rb = torch.zeros_like(gw), torch.zeros_like(gw1), torch.zeros_like(gb)
        ops.embed_fm_bwd(table, rows_all, cat_pos, pad_row, dx_all, dense_w, num_pos, Er, Sr, dE, d_fm, d_lr, rt, rt1, rw, rw1, rb, D, R)

        def close(a, b):
            return bool(((a - b).abs() <= 1e-5 * b.abs() + 2e-5 * b.abs().max()).all())
        ok = torch.equal(E, Er[sl]) and close(S, Sr[sl]) and close(fm, fmr[sl]) and close(lr, lrr[sl])
        ok = ok and all(close(a, b) for a, b in ((gt, rt), (gt1, rt1), (gw, rw), (gw1, rw1), (gb, rb)))
    except Exception as e:                       # a failed check must not take the bench line with it; it is reported
        sys.stderr.write("sharded_parity[rank %d]: %r\n" % (rank, e))
        ok = False
    finally:
        sh.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def run_sharded(args, rank, world, dev, K, W):
    """configs[3] on the ranks of an initialised process group -> the JSON object (rank 0; None elsewhere)."""
    import torch.distributed as dist
    from recbox_b200 import sharded
    B, F, Fn, D, V = CFG["B"], CFG["F"], CFG["Fn"], args.dim, args.rows_per_field
    Ft, R = F + Fn, F * V
    if args.shard_mode == "auto":       # the fused kernels with the exchange inside them: best measured at N = 1, 2 and 8
        args.shard_mode = "peer"        # (profiles/r2_sharded_runs.jsonl: 673 us at N = 8 vs 800 us streamed)
    if args.shard_layout == "auto":     # ROW+LR wins on multi-GB tables at every N (profiles/r1_sharded_runs.jsonl, r2e / r1w)
        args.shard_layout = "rowlr" if (args.shard_mode in ("peer", "stream") and not args.no_lr and D in (4, 8, 16)) else "split"
    parity_ok = sharded_parity(args, rank, world, dev) if not args.no_lr else None
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

    for i in range(W):
        step(i)
    sh.check_overflow()
    sampler = ClockSampler(dev.index)
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
        if world > 1:                          # slowest and fastest rank per phase
            names = sorted(phases)
            mine = torch.tensor([phases[n] for n in names], dtype=torch.float64, device=dev)
            hi, lo = mine.clone(), mine.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            phases = {n: {"rank0": phases[n], "min": float(lo[i]), "max": float(hi[i])} for i, n in enumerate(names)}
    times = torch.tensor([ms_total, t_f, t_b, t_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, t_f, t_b, t_s = times.tolist()
    line = None
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
            "mode": args.shard_mode, "layout": args.shard_layout, "parity_ok": parity_ok,
            "fwd_us": t_f * 1e3, "bwd_us": t_b * 1e3,
            "nvlink_gbs": {"fwd": nv_f / t_f / 1e6, "bwd": nv_b / t_b / 1e6, "peak": 770.0,
                           "what": "payload bytes (remote rows + first-order weights) per direction per GPU / phase time; peak = measured "
                                   "peer copy per direction (B200_PROFILING.md)"},
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
    sh.close()
    del sh
    torch.cuda.empty_cache()
    return line


def main_sharded(args, rank, world, local_rank):
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run_sharded(args, rank, world, dev, args.steps, max(args.warmup, 3))
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        _teardown(dist)


def main_matching(args, rank, world, local_rank):
    """BASELINE configs[2] / configs[4] (tools/matching_workloads.py)."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import matching_workloads as mw
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = (mw.run_dssm if args.workload == "dssm" else mw.run_sasrec)(args, rank, world, dev, sys.modules[__name__])
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        _teardown(dist)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ids", default="uniform", choices=["uniform", "zipf"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs under ncu)")
    ap.add_argument("--zero", default="side", choices=["side", "inline"],
                    help="gradient zero-fill on a side stream under the forward (default) or on the step's stream before it")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nvls", "nccl"],
                    help="replica gradient exchange at N > 1: in-switch all-reduce kernel over multicast memory or NCCL; auto = the "
                         "faster one measured at that N (r2ad / r2ae: NCCL 157 / 201 us at N = 2 / 4, the kernel 205 us at N = 8 vs NCCL 284)")
    ap.add_argument("--bwd-saved-e", action="store_true",
                    help="cfg2: the backward reads the forward's saved E instead of re-gathering e from the table (the round-1 form)")
    ap.add_argument("--eager", action="store_true", help="layer-API e2e legs without CUDA-graph replay (host-bound: ~80 Parameters cross autograd)")
    ap.add_argument("--no-sharded", action="store_true", help="cfg2 run without the attached configs[3] (100M-row sharded table) object")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "sharded", "dssm", "sasrec"],
                    help="cfg2: BASELINE configs[1], replicated 1M-row table (default, the metric's config); "
                         "sharded: configs[3], 100M-row table row-sharded over the ranks; dssm: configs[2]; sasrec: configs[4]")
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
    elif args.workload in ("dssm", "sasrec"):
        main_matching(args, rank, world, local_rank)
    else:
        main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
