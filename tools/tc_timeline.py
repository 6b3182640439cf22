"""Debug tool: phase timeline of the tensor-core flow kernel (CTA 0, both tiles, Mobius steps 40..47).
Build with  RNF_NVCC_EXTRA=-DRNF_TC_TRACE=1 python -m rotationnormflow_b200.build --force  and run on the GPU box."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rotationnormflow_b200 import _cabi, grid as rgrid
from rotationnormflow_b200.flow import _program

cfg, flow = bench.build_flow()
flow = flow.cuda().eval()
lib = _cabi.load()
raw = C.CDLL(_cabi.library_path())
trace = torch.zeros(2 * 8 * 32, dtype=torch.int64, device="cuda")
raw.rnf_debug_set_trace.argtypes = [C.c_void_p]
raw.rnf_debug_set_trace(C.c_void_p(trace.data_ptr()))
G = rgrid.healpix_grid(5)
feat = torch.relu(torch.randn(1, 2048)).cuda()
out = flow.grid_log_prob(G, feat, mlp_mode=os.environ.get("RNF_TRACE_MODE", "tc"))
torch.cuda.synchronize()
t = trace.cpu().reshape(2, 8, 32)
names = {0: "start", 1: "turn", 2: "prologue", 3: "bar1", 4: "iss1", 5: "mma1", 6: "epi1", 7: "bar2", 8: "iss2", 9: "mma2", 10: "epi2",
         11: "bar3", 12: "iss3", 13: "mma3", 14: "epi3", 15: "bar4", 16: "iss4", 17: "mmaA", 18: "mix", 19: "xchg", 20: "end"}
t0 = int(t[t > 0].min())
for s in range(2, 6):
    for tile in range(2):
        row = t[tile, s]
        base = int(row[0])
        print(f"step {40+s} tile {tile}: start @{base - t0:7d} | " + " ".join(f"{names[i]}+{int(row[i]) - int(row[i-1])}" for i in range(1, 21)))
# absolute phase boundaries: chain = [start, iss4], wait = [iss4, mmaA], mixture = [mmaA, mix], tail = [mix, end]
print("absolute timeline (cycles since first stamp): chain_start, fc_last_issued, mixture_start, mixture_end, layer_end")
for s in range(2, 6):
    for tile in range(2):
        row = t[tile, s]
        print(f"step {40+s} tile {tile}: " + " ".join(f"{int(row[i]) - t0:7d}" for i in (0, 16, 17, 18, 20)))
