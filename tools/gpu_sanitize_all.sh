# compute-sanitizer memcheck over every single-GPU parity test except the full-size ones
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py tests/test_loader_gpu.py tests/test_dedup_optim_gpu.py tests/test_retrieval_gpu.py \
  -m gpu -q -p no:cacheprovider > gpurun_out/r1z_memcheck_all.log 2>&1
echo "memcheck exit code $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r1z_memcheck_all.log | head -20
