"""Profiling driver: Flow.inverse of the symsol2 stack (F=512), 16 images x 32768 samples, twice.
Run under ncu:  ncu --set full --clock-control none --import-source on -k regex:flow_row -c 1 -o gpurun_out/inv python tools/profile_inverse.py"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rotationnormflow_b200 as rnf
from rotationnormflow_b200 import grid as rgrid

torch.manual_seed(0); np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    flow = rnf.get_flow(rnf.load_config("symsol2")).cuda().eval()
n_img, n_per = 16, 32768
base = rgrid.generate_queries(n_img * n_per, "random", device="cuda")
feat = torch.relu(torch.randn(n_img, 512)).cuda()
idx = torch.arange(n_img * n_per, device="cuda", dtype=torch.int32) // n_per
for _ in range(2):
  with torch.no_grad():
    R, l = flow.inverse(base, feat, feature_index=idx, mlp_mode=os.environ.get("MODE", "tc"))
torch.cuda.synchronize()
print(R.shape, float(l.mean()))
