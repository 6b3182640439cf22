"""Profiling driver for the other BASELINE configurations (GPU box, under ncu):
   CFG=1: raw.yml Flow.forward on 100 000 rotations (flow_t4_kernel<false>, tail-bound launch)
   CFG=3: modelnet_fisher.yml (F=2080, 24 conditional affines, matrix-Fisher base) grid log-prob, level-5 grid x 2 images
Run:  CFG=3 ncu --set full --clock-control none --import-source on -k regex:flow_t4 -s 1 -c 1 -o gpurun_out/x python tools/profile_cfg.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import rnf_oracle as orc
from rotationnormflow_b200 import grid as rgrid

which = os.environ.get("CFG", "3")
if which == "1":
    cfg, flow = bench.build_flow("raw")
    flow = flow.cuda().eval()
    R = orc.random_rotations(100_000, torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        for _ in range(2):
            out = flow(R)
    torch.cuda.synchronize()
    print("cfg1", float(out[1].mean()))
else:
    cfg, flow = bench.build_flow("modelnet_fisher")
    flow = flow.cuda().eval()
    G = rgrid.healpix_grid(5)
    g = torch.Generator().manual_seed(3)
    feat = torch.relu(torch.randn(2, orc.feature_dim_of(cfg), generator=g)).cuda()
    A = (torch.randn(2, 3, 3, generator=g) * 3).cuda()
    with torch.no_grad():
        for _ in range(2):
            out = flow.grid_log_prob(G, feat, fisher_A=A)
    torch.cuda.synchronize()
    print("cfg3", out["argmax"])
