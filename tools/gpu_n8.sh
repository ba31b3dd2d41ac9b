# 8-GPU sharded-table measurement (configs[3]): peer mode, both shard layouts.
mkdir -p gpurun_out
MODES="peer" bash tools/gpu_shard_bench.sh r1w_rowlr 8 --shard-layout rowlr --steps 30
MODES="peer" bash tools/gpu_shard_bench.sh r1w_split 8 --steps 30
MODES="push" bash tools/gpu_shard_bench.sh r1w_push 8 --steps 30
