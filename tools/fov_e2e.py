"""Full-FOV end-to-end wall time of `fpmMain` at several GPU counts of one box (the bench's full_fov_e2e leg, run alone).
usage: python tools/fov_e2e.py 1,2,4,8 [iters]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fpm-opencv_b200")]
import bench
g = bench.geometry()
stacks = bench.distinct_stacks(g, 8)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for n in [int(x) for x in sys.argv[1].split(",")]:
    r = bench.fov_e2e_leg(n, g, stacks, iters)
    print(json.dumps(r), flush=True)
