"""Evaluations of the mixture map per sample and layer executed by the inverse kernel, and its rate, as the mixtures get harder:
the fc_last weights of a seeded symsol2 model are scaled by s (s = 1: random init, the bench; larger s: centres up to |w'| = 0.7 and
peaky weights, as in a trained model).  GPU box."""
import contextlib, ctypes as C, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rotationnormflow_b200 as rnf
from rotationnormflow_b200 import _cabi, grid as rgrid
from oracle import rnf_oracle as orc

_cabi.load()
lib = C.CDLL(_cabi.library_path())
lib.rnf_debug_set_probe_counter.argtypes = [C.c_void_p]
lib.rnf_debug_set_probe_counter.restype = None
n_img, n_per = 8, 32768
for scale in (1.0, 3.0, 6.0, 12.0):
    torch.manual_seed(0); np.random.seed(0)
    cfg = rnf.load_config("symsol2")
    with contextlib.redirect_stdout(io.StringIO()):
        flow = rnf.get_flow(cfg)
    with torch.no_grad():
        for k, v in flow.state_dict().items():
            if "fc_last" in k:
                v.mul_(scale)
    flow = flow.cuda().eval()
    flow.invalidate_cache()
    n_mob = sum(1 for k in flow.state_dict() if k.endswith("conditioner.fc_last.bias"))
    base = rgrid.generate_queries(n_img * n_per, "random", device="cuda")
    feat = torch.relu(torch.randn(n_img, 512)).cuda()
    idx = torch.arange(n_img * n_per, device="cuda", dtype=torch.int32) // n_per
    counter = torch.zeros(1, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        flow.inverse(base, feat, feature_index=idx)
        lib.rnf_debug_set_probe_counter(C.c_void_p(counter.data_ptr()))
        flow.inverse(base, feat, feature_index=idx)
        torch.cuda.synchronize()
        lib.rnf_debug_set_probe_counter(C.c_void_p(0))
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(3):
            R, l = flow.inverse(base, feat, feature_index=idx)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 3
        # parity on a subset against the fp64 oracle
        o = orc.OracleFlow(cfg, {k: v.cpu() for k, v in flow.state_dict().items()}, torch.float64)
        sub = slice(0, 256)
        Ro, lo = o.inverse(base[sub].cpu().double(), feat[:1].cpu().double().expand(256, -1))
        close = ((R[sub].cpu().double() - Ro).abs().amax((1, 2)) < 1e-5).float().mean().item()
    ev = counter.item() / (n_img * n_per * n_mob)
    print(f"fc_last x {scale:4.1f}: {ev:.2f} evaluations per sample-layer, {n_img * n_per / ms / 1e3:.1f} M samples/s, rows within 1e-5 of the fp64 oracle {close:.3f}")
