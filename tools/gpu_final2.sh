mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r1zz_pytest.log; tail -6 gpurun_out/r1zz_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/kernel_rooflines.py r1zz > gpurun_out/r1zz_kernels.md 2> gpurun_out/r1zz_kernels.err; grep -E "inner_product|unique|split" gpurun_out/r1zz_kernels.md
