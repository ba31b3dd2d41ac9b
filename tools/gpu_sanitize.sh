# compute-sanitizer memcheck over the kernels written this round (small parity cases only)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_kernels_gpu.py tests/test_retrieval_gpu.py tests/test_dedup_optim_gpu.py -m gpu -q -x -p no:cacheprovider \
  -k "interact_random or interact_known or power_sums or pool_materialised or split_batch or topk_ip_matches or topk_ip_ties or rank_metrics or sample_negatives or unique" \
  > gpurun_out/r1z_memcheck.log 2>&1
echo "memcheck exit code $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r1z_memcheck.log | head -20
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 \
  python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "interact_random and (39-16 or 8-4 or 33-32)" \
  > gpurun_out/r1z_racecheck.log 2>&1
echo "racecheck exit code $?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/r1z_racecheck.log | head -10
