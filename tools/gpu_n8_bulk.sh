# 8-GPU: peer/rowlr step time with the TMA bulk-reduce backward (peer rows only / all rows)
mkdir -p gpurun_out
RBX_LIB_PATH=$PWD/build/variants/bulk_remote.so MODES="peer" bash tools/gpu_shard_bench.sh r1z_bulk_remote 8 --shard-layout rowlr --steps 30
RBX_LIB_PATH=$PWD/build/variants/bulk.so MODES="peer" bash tools/gpu_shard_bench.sh r1z_bulk_all 8 --shard-layout rowlr --steps 30
