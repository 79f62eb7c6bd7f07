set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r01_bench_plain.json 2> gpurun_out/r01_bench_plain.err || exit 1
tail -c 600 gpurun_out/r01_bench_plain.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/r01_ncu_launch.log 2>&1
echo launches rc=$?
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fpm_update_kernel -c 1 -f -o gpurun_out/r01_update_kernel $CMD > gpurun_out/r01_ncu_full.log 2>&1
echo full rc=$?
ls -la gpurun_out | tail
