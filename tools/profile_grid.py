"""Profiling driver: two grid log-prob launches of the bench model (symsol, F=2048, level-5 grid x 1 image) in mode $MODE.
Run under ncu:  ncu --set full --clock-control none --import-source on -k regex:flow_ -c 1 -o gpurun_out/x python tools/profile_grid.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rotationnormflow_b200 import grid as rgrid

mode = os.environ.get("MODE", "tc")
cfg, flow = bench.build_flow("symsol", feature_dim=2048)
flow = flow.cuda().eval()
G = rgrid.healpix_grid(5)
feat = torch.relu(torch.randn(1, 2048)).cuda()
for _ in range(2):
    out = flow.grid_log_prob(G, feat, mlp_mode=mode)
torch.cuda.synchronize()
print(mode, out["argmax"])
