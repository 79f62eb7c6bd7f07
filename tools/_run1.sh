( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
  timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -6
) > gpurun_out/tests_r02c.log 2>&1
tail -20 gpurun_out/tests_r02c.log | cut -c1-250
