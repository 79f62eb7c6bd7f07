# Round-2 closing measurement pass after the cluster-kernel rework (run under gpurun): bench line, configuration sweep,
# stage table and ncu capture of the 4-CTA cluster kernel on one 128 x 128 tile.
# Needs the developer timing build of the cluster kernels: make -C fpm-opencv_b200 lib/libfpmb200_fastd.so (built here, it
# travels with the snapshot).  Protocol stress of the cluster kernels: timeout 120 python tools/dev_stress_cluster.py 20
set -x
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err || exit 1
timeout 300 python tools/dev_sweep.py cfg1_mono_np64 cfg2_fLEDc_np128 cfg3b_cellScope_np64 cfg3_cellScope_np256 cfg4_dogStomach_np128 cfg5_cellscope2_np128 cfg5b_cellscope2_np256 cfg7_mono_np90 cfg8_cellScope_np100 cfg4s_dogStomach_np200 > gpurun_out/r02_final_config_sweep.txt 2>&1
( FPM_LIB=libfpmb200_fastd.so FPM_TILES=1 FPM_CLUSTER=4 timeout 60 python tools/dev_stages.py cfg2_fLEDc_np128 cfg5_cellscope2_np128
  FPM_LIB=libfpmb200_fastd.so FPM_TILES=1 FPM_CLUSTER=2 timeout 60 python tools/dev_stages.py cfg2_fLEDc_np128
  FPM_LIB=libfpmb200_fastd.so FPM_TILES=1 timeout 60 python tools/dev_stages.py cfg5b_cellscope2_np256 cfg3_cellScope_np256 ) > gpurun_out/r02_cluster_stage_cycles.txt 2>&1
CMD="python tools/dev_ncu.py cfg2_fLEDc_np128 1 10 2"
timeout 60 $CMD > gpurun_out/plain_ncu_cmd.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:fpm_update_cluster_kernel -s 1 -c 1 -f -o gpurun_out/r02_cluster_kernel $CMD > gpurun_out/ncu_cluster.log 2>&1
ls -la gpurun_out | tail -6
