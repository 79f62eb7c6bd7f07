( timeout 120 python tools/dev_stress_cluster.py 20 2>&1 | cut -c1-110 | tail -3; echo "stress rc $?"
  timeout 30 python tools/dev_cluster.py cfg4_dogStomach_np128 40 2,1 10 2>&1 | tail -2 | cut -c1-200
  timeout 30 python tools/dev_cluster.py cfg4_dogStomach_np128 74 2,1 10 2>&1 | tail -2 | cut -c1-200
  timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
  timeout 100 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | cut -c1-200
) > gpurun_out/tests_r02f.log 2>&1
tail -24 gpurun_out/tests_r02f.log
