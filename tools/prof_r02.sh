# Round-2 measurement pass (run under gpurun): bench (both arms), stage tables, config sweep, ncu launch list.
# (the ncu --set full capture of the update kernel comes from tools/dev_ncu.py: under bench.py's 592-tile launch the
#  replay passes fail to save / restore 10 GB of buffers)
set -x
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err || exit 1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2>/dev/null
( FPM_TILES=148 FPM_CLUSTER=1 timeout 200 python tools/dev_stages.py cfg4_dogStomach_np128 cfg1_mono_np64 cfg5_cellscope2_np128 ) > gpurun_out/r02_final_stage_cycles.txt 2>&1
timeout 500 python tools/dev_sweep.py cfg1_mono_np64 cfg2_fLEDc_np128 cfg3b_cellScope_np64 cfg3_cellScope_np256 cfg4_dogStomach_np128 cfg5_cellscope2_np128 cfg5b_cellscope2_np256 cfg7_mono_np90 cfg8_cellScope_np100 cfg4s_dogStomach_np200 > gpurun_out/r02_final_config_sweep.txt 2>&1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fov-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final_launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
ls -la gpurun_out | tail -6
