"""Grid SDF query microbenchmark of bench.py (256 worlds x 64^3 grids x 131072 points), for ncu."""
import sys
sys.path.insert(0, '.')
import torch
import bench
from diffsdfsim_b200 import ops
out = bench.sdf_query_roofline(torch.device('cuda', 0))
print({k: out[k] for k in ('frac', 'avg_launch_ms')}, out['value_only'])
