set -x
mkdir -p gpurun_out
( FPM_TILES=148 FPM_CLUSTER=1 timeout 200 python tools/dev_stages.py cfg4_dogStomach_np128 cfg1_mono_np64 cfg5_cellscope2_np128
  FPM_TILES=1,16 timeout 200 python tools/dev_stages.py cfg5b_cellscope2_np256 cfg3_cellScope_np256
  FPM_TILES=1 FPM_CLUSTER=4 timeout 200 python tools/dev_stages.py cfg2_fLEDc_np128 cfg5_cellscope2_np128 ) > gpurun_out/r01_stage_cycles.txt 2>&1
timeout 300 python tools/dev_sweep.py > gpurun_out/r01_config_sweep.txt 2>&1
CMD="python tools/dev_sweep.py cfg5b_cellscope2_np256"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fpm_update_cluster_kernel -s 2 -c 1 -f -o gpurun_out/r01_cluster_kernel $CMD > gpurun_out/r01_ncu_cluster.log 2>&1
echo rc=$?
tail -3 gpurun_out/r01_ncu_cluster.log
cat gpurun_out/r01_config_sweep.txt
