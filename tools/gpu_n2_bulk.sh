# 2-GPU: sharded parity + peer/rowlr step time, default library vs the TMA bulk-reduce backward variant
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -5
RBX_LIB_PATH=$PWD/build/variants/bulk.so timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -8
MODES="peer" bash tools/gpu_shard_bench.sh r1y_base 2 --shard-layout rowlr --steps 30
RBX_LIB_PATH=$PWD/build/variants/bulk.so MODES="peer" bash tools/gpu_shard_bench.sh r1y_bulk 2 --shard-layout rowlr --steps 30
RBX_LIB_PATH=$PWD/build/variants/bulk.so MODES="peer" bash tools/gpu_shard_bench.sh r1y_bulk_split 2 --steps 30
