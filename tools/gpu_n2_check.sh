# 2-GPU check: sharded parity (world 2) + the default multi-GPU bench path exactly as the driver launches it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -2 gpurun_out/r2d_bench_n2.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2d_bench_ref_n2.json 2> gpurun_out/r2d_bench_ref_n2.err
python - <<PY
import json
for f in ("r2d_bench_n2.json", "r2d_bench_ref_n2.json"):
    try:
        d = json.loads([l for l in open("gpurun_out/" + f) if l.startswith("{")][-1])
        print(f, "n_gpus", d["n_gpus"], "value %.4gM" % (d["value"] / 1e6), "e2e %.4gM" % (d["e2e"]["value"] / 1e6), d.get("clocks"), d.get("impl"))
    except Exception as e:
        print(f, "ERR", e)
PY
