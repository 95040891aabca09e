"""Stand-alone resampling and per-timestep transition kernels against the HBM roofline (8 B per particle for
resampling, 8 du + 16 B per particle-step for the transition).  usage: python scripts/hbm_kernels_bench.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

res = bench.secondary_hbm_kernels()
for o in res['kernels']:
    print(json.dumps(o))
