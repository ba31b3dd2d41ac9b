# Does the ROW+LR physical layout (row and first-order weight in one 128-byte line) pay on ONE GPU too?  1M-row table.
mkdir -p gpurun_out
MODES="peer" bash tools/gpu_shard_bench.sh r2e_split 1 --rows-per-field 38462 --steps 50
MODES="peer" bash tools/gpu_shard_bench.sh r2e_rowlr 1 --rows-per-field 38462 --shard-layout rowlr --steps 50
MODES="peer" bash tools/gpu_shard_bench.sh r2e_split100m 1 --steps 30
MODES="peer" bash tools/gpu_shard_bench.sh r2e_rowlr100m 1 --shard-layout rowlr --steps 30
python tools/train_step_bench.py r2e > gpurun_out/r2e_train_step.log 2>&1; tail -6 gpurun_out/r2e_train_step.log
