"""Round duration as a function of the tiles in flight per SM (flow_t4, RNF_T4_ACTIVE override) and config-1 rate (GPU box)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import bench
from oracle import rnf_oracle as orc
cfg, flow = bench.build_flow("raw")
flow = flow.cuda().eval()
def rate(n):
    R = orc.random_rotations(n, torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        for _ in range(3): flow(R)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): flow(R)
        e.record(); torch.cuda.synchronize()
    return n / (s.elapsed_time(e) / 10 * 1e-3) / 1e6
act = os.environ.get("RNF_T4_ACTIVE", "auto")
a = int(act) if act != "auto" else 4
full = 148 * a * 128 * 8            # eight full rounds of `a` tiles per SM
print(f"active={act}: 100k rows {rate(100000):.1f} M rot/s | 8 full rounds of {a}: {rate(full):.1f} M rot/s | 50k {rate(50000):.1f} | 200k {rate(200000):.1f} | 300k {rate(300000):.1f}")
'''
for act in ("auto", "4", "3", "2", "1"):
    env = dict(os.environ)
    if act == "auto":
        env.pop("RNF_T4_ACTIVE", None)
    else:
        env["RNF_T4_ACTIVE"] = act
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print((r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1])
