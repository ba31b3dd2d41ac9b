import sys, os, torch, time
sys.path.insert(0, "/root/repo")
from recbox_b200 import ops
torch.manual_seed(0)
dev = "cuda"
items = torch.randn(10_000_000, 64, device=dev)
for U, k in ((1024, 100), (128, 100), (4096, 100)):
    q = torch.randn(U, 64, device=dev)
    for eng in ("2", "1"):
        os.environ["RBX_TOPK_ENGINE"] = eng
        for persist in (("1", "0") if eng == "2" else ("1",)):
            os.environ["RBX_TOPK_PERSIST"] = persist
            ops.topk_ip(q, items, k); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s, i = ops.topk_ip(q, items, k); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b)
            print("U=%d k=%d engine=%s persist=%s: %.2f ms  %.1f TFLOP/s" % (U, k, eng, persist, ms, 2.0 * U * 1e7 * 64 / ms / 1e9), flush=True)
            if eng == "2" and persist == "1": ref = i.clone()
            else: print("   index agreement with persistent: %.6f" % float((i == ref).float().mean()))
