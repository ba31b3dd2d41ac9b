mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>&1 | head -12
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/r1j_sharded2.log
cat gpurun_out/r1j_sharded2.log
