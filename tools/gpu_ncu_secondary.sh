mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_ip_fwd_warp|k_ip_bwd_warp|k_ip_filter_mma" -c 6 -o gpurun_out/r1z_secondary -f python tools/ncu_secondary.py > gpurun_out/r1z_secondary.log 2>&1
tail -3 gpurun_out/r1z_secondary.log
ls -la gpurun_out/r1z_secondary.ncu-rep
