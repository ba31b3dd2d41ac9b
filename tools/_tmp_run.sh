mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r1v_pytest.log; cat gpurun_out/r1v_pytest.log
MODES="peer" bash tools/gpu_shard_bench.sh r1v_split 2
MODES="peer" bash tools/gpu_shard_bench.sh r1v_rowlr 2 --shard-layout rowlr
