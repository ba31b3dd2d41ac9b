set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r1i_pytest.log
cat gpurun_out/r1i_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3
