#!/bin/bash
# One entry point for the GPU-box runs of this repo (replaces the per-experiment scripts of round 1).
#   tools/gpu_run.sh <target> [extra args]     -- run on the box via: gpurun [--gpus N] -- 'bash tools/gpu_run.sh <target>'
# Everything a run produces goes to gpurun_out/<tag>_* (tag = $RBX_TAG, default r2).
set -u
TAG=${RBX_TAG:-r2}
OUT=gpurun_out
mkdir -p $OUT
N=$(python -c 'import torch; print(torch.cuda.device_count())')
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
case "$1" in
  tests)        # the whole -m gpu suite
    python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_gputests.txt ;;
  tests-sharded)
    python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_sharded_tests.txt ;;
  bench)        # the driver's line, N = visible GPUs
    shift
    if [ "$N" -gt 1 ]; then $TR bench.py --gpus $N "$@"; else python bench.py "$@"; fi 2>&1 | tail -3 | tee -a $OUT/${TAG}_bench.jsonl ;;
  sharded)      # configs[3] lines: tools/gpu_run.sh sharded <mode> <layout> [more bench args]
    shift; MODE=$1; LAYOUT=$2; shift 2
    if [ "$N" -gt 1 ]; then $TR bench.py --gpus $N --workload sharded --shard-mode $MODE --shard-layout $LAYOUT --steps 30 --warmup 5 "$@"
    else python bench.py --workload sharded --shard-mode $MODE --shard-layout $LAYOUT --steps 30 --warmup 5 "$@"; fi 2>&1 | tail -2 | tee -a $OUT/${TAG}_sharded.jsonl ;;
  nvlink)       # all-to-all write bandwidth microbenchmark
    shift; $TR tools/nvlink_microbench.py "$@" 2>&1 | tail -2 | tee -a $OUT/${TAG}_nvlink.jsonl ;;
  *) echo "unknown target $1"; exit 2 ;;
esac
