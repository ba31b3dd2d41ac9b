# last check of the round after the unique rewrite: dedup / optimizer parity, memcheck on it, launch list, smoke, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dedup_optim_gpu.py tests/test_retrieval_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_dedup_optim_gpu.py -m gpu -q -x -p no:cacheprovider -k "unique or collate" 2>&1 | grep -E "passed|failed|ERROR SUMMARY"
bash tools/gpu_ncu_small.sh | grep -v k_sample | tail -14
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/r1z2_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r1z2_bench.json').read().strip().splitlines()[-1]); print('value %.4gM e2e %.4gM packed %.4gM frac %.3f pair %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e_packed']['value']/1e6, d['roofline']['frac'], d['roofline']['pair_frac']))"
