mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1u_pytest.log; cat gpurun_out/r1u_pytest.log
timeout 300 bash tools/sweep.sh r1u base nofuse fuse_noagg
BENCH_ARGS="--ids zipf" timeout 300 bash tools/sweep.sh r1u_zipf base nofuse
