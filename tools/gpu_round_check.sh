# Round-end style check on one B200: GPU tests, smoke, both bench arms, ncu launch list + full capture.
tag=${1:-r1s}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --ids zipf --no-cpu-baseline > gpurun_out/${tag}_bench_zipf.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_embed_fm -s 6 -c 2 -o gpurun_out/${tag}_prof -f \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
python - <<PY
import json
for f in ("${tag}_bench_reference.json", "${tag}_bench.json", "${tag}_bench_zipf.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        print(f, "value %.3gM" % (d["value"] / 1e6), "e2e %.3gM" % (d["e2e"]["value"] / 1e6), d.get("roofline"), d.get("cpu_baseline"), d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/${tag}_bench.err
