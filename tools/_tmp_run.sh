MODES="peer" bash tools/gpu_shard_bench.sh r1p_nolr 2 --peer-alloc symm --no-lr
MODES="peer" bash tools/gpu_shard_bench.sh r1p_d64 2 --peer-alloc symm --dim 64 --rows-per-field 961538
MODES="peer" bash tools/gpu_shard_bench.sh r1p_d64_nolr 2 --peer-alloc symm --dim 64 --rows-per-field 961538 --no-lr
