# Round check + per-kernel rooflines on one B200.  usage: bash tools/gpu_round_check2.sh <tag>
tag=${1:-r1u}
bash tools/gpu_round_check.sh $tag
python tools/kernel_rooflines.py $tag > gpurun_out/${tag}_kernels.md 2> gpurun_out/${tag}_kernels.err
cat gpurun_out/${tag}_kernels.md; tail -3 gpurun_out/${tag}_kernels.err
