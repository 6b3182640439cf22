"""Where does the time of the literal drop-in call go?  (GPU box)  Times each stage of
flow(samples, feature.repeat(N, 1)) for one image: the caller's repeat, the run-length pass (csrc/dedup.cu), the per-image
conditioner, the flow kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rotationnormflow_b200 as rnf
from rotationnormflow_b200 import engine, grid as rgrid

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5, w=3):
    for _ in range(w):
        fn()
    tot = 0.0
    per = []
    for _ in range(n):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        tot += s.elapsed_time(e)
        per.append(round(s.elapsed_time(e), 3))
    ALL.append(per)
    return tot / n


ALL = []
rows = 500_000
for F in (2048, 512):
    cfg = rnf.load_config("symsol", feature_dim=F)
    torch.manual_seed(0)
    flow = rnf.get_flow(cfg).to(dev).eval()
    chunk = rgrid.healpix_grid(5, 0, rows, device=dev)
    f1 = torch.relu(torch.randn(1, F, device=dev))
    feats = f1.repeat(rows, 1)
    with torch.no_grad():
        t_rep = timed(lambda: f1.repeat(rows, 1))
        t_dd = timed(lambda: engine.dedup_rows(feats, engine.DEDUP_CAP))
        idx, first, count = engine.dedup_rows(feats, engine.DEDUP_CAP)
        from rotationnormflow_b200.flow import _program
        prog = _program(flow, list(flow.layers), flow._perm_rows(), flow.feature_dim, dev)
        t_cond = timed(lambda: prog.condition_runs(feats, first, count, engine.DEDUP_CAP))
        cond = prog.condition_runs(feats, first, count, engine.DEDUP_CAP)
        t_run = timed(lambda: prog.run(chunk, cond, engine.DEDUP_CAP, idx, 0, False, "tc"))
        t_all = timed(lambda: flow(chunk, feats))
        t_exp = timed(lambda: flow(chunk, f1.expand(rows, F)))
        t_grid = timed(lambda: flow.grid_log_prob(chunk, f1))
    print(f"F={F}: repeat {t_rep:.3f} ms ({rows*F*4/t_rep/1e6:.0f} GB/s)  dedup {t_dd:.3f} ms ({rows*F*4/t_dd/1e6:.0f} GB/s)  condition_runs(cap {engine.DEDUP_CAP}) {t_cond:.3f} ms"
          f"  flow kernel {t_run:.3f} ms  | flow(R, repeated) {t_all:.3f} ms  flow(R, expanded) {t_exp:.3f} ms  grid_log_prob {t_grid:.3f} ms")
    print("   per-iteration ms:", ALL[-3:])
    del flow, feats, chunk
    torch.cuda.empty_cache()
