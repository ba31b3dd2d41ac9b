# usage: tools/gpu_shard_bench.sh <tag> <ngpus> [extra bench args]
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
out=gpurun_out/${tag}_shard_n${n}.jsonl
: > $out
for mode in ${MODES:-push peer a2a}; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --workload sharded --shard-mode $mode --steps 30 --warmup 5 "$@" >> $out 2>> gpurun_out/${tag}_shard.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --workload sharded --shard-mode $mode --steps 30 --warmup 5 "$@" >> $out 2>> gpurun_out/${tag}_shard.err
  fi
done
tail -5 gpurun_out/${tag}_shard.err
python - <<PY
import json
for l in open('$out'):
    if not l.startswith('{'): continue
    d = json.loads(l); k = d['kernels']
    print(d['config']['workload'][-12:], 'n=%d value %.1fM samples/s  step %.1f us  fwd %.1f us (nvlink %.0f GB/s)  bwd %.1f us (nvlink %.0f GB/s)  barrier %.1f us' % (
        d['n_gpus'], d['value']/1e6, d['ms_per_step']*1e3, k['fwd']['ms']*1e3, k['fwd']['nvlink_gbs'], k['bwd']['ms']*1e3, k['bwd']['nvlink_gbs'], k['step_barrier']['ms']*1e3))
PY
