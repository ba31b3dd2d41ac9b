mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_layers_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r1t_pytest.log; cat gpurun_out/r1t_pytest.log
timeout 300 bash tools/sweep.sh r1t base noagg
BENCH_ARGS="--ids zipf" timeout 300 bash tools/sweep.sh r1t_zipf base noagg
