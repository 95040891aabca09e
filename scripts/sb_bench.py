"""configs[2] (Gaussian Schroedinger bridge particle Gibbs) alone, for ncu launch lists.  usage: python scripts/sb_bench.py [chains]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
print(json.dumps(bench.secondary_sb(chains=int(sys.argv[1]) if len(sys.argv) > 1 else 16384, steps=2)))
