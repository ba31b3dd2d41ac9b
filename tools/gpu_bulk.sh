# TMA bulk-reduce backward variant: parity under the variant library, then base vs variant bench lines.
mkdir -p gpurun_out
RBX_LIB_PATH=$PWD/build/variants/bulk.so timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r1x_bulk_pytest.log; tail -15 gpurun_out/r1x_bulk_pytest.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_loader_gpu.py tests/test_retrieval_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r1x_pytest.log; tail -8 gpurun_out/r1x_pytest.log
bash tools/sweep.sh r1x base bulk bulk_b3 bulk_u2
BENCH_ARGS="--ids zipf" bash tools/sweep.sh r1x_zipf base bulk
