# Developer script: GPU parity suite, stage tables for the timing build and experiment builds, short bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/gpu_tests.log; cat gpurun_out/gpu_tests.log
for lib in libfpmb200_timing.so $EXP_LIBS; do
  echo "=== $lib"; FPM_LIB=$lib FPM_TILES=148 FPM_CLUSTER=1 timeout 120 python tools/dev_stages.py cfg4_dogStomach_np128 ${EXTRA_CFGS}
done > gpurun_out/stages.txt 2>&1
cat gpurun_out/stages.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python -c "import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['config']['kernel'])"
