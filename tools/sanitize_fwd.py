"""Small forward call of the tc path for compute-sanitizer runs (GPU box)."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rotationnormflow_b200 as rnf
from oracle import rnf_oracle as orc
n = int(os.environ.get("N", "600"))
cfg = rnf.load_config("symsol", layers=3, feature_dim=32)
torch.manual_seed(0); np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    flow = rnf.get_flow(cfg).cuda().eval()
gen = torch.Generator().manual_seed(1)
R = orc.random_rotations(n, gen).float().cuda()
feat = torch.relu(torch.randn(1, 32, generator=gen)).cuda()
with torch.no_grad():
    Rg, lg = flow(R, feat, feature_index=torch.zeros(n, dtype=torch.int32, device="cuda"), mlp_mode="tc")
torch.cuda.synchronize()
print("forward ok", float(lg.mean()))
