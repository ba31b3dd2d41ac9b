# Development iteration on one B200: all GPU tests (no -x, failures listed), per-kernel rooflines, a bench line.
tag=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/${tag}_pytest.log; tail -40 gpurun_out/${tag}_pytest.log
timeout 600 python tools/kernel_rooflines.py $tag > gpurun_out/${tag}_kernels.md 2> gpurun_out/${tag}_kernels.err
cat gpurun_out/${tag}_kernels.md; tail -5 gpurun_out/${tag}_kernels.err
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("value %.4gM e2e %.4gM e2e_packed %.4gM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e_packed"]["value"]/1e6), d["roofline"]["frac"], d["roofline"]["pair_frac"], d["kernels"])
PY
