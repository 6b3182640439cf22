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
trace = torch.zeros(4 * 8 * 32, dtype=torch.int64, device="cuda")
raw.rnf_debug_set_trace.argtypes = [C.c_void_p]
raw.rnf_debug_set_trace(C.c_void_p(trace.data_ptr()))
G = rgrid.healpix_grid(5)
feat = torch.relu(torch.randn(1, 2048)).cuda()
out = flow.grid_log_prob(G, feat, mlp_mode=os.environ.get("RNF_TRACE_MODE", "tc"))
torch.cuda.synchronize()
mode = os.environ.get("RNF_TRACE_MODE", "tc")
if mode == "tc":
    names = {0: "start", 1: "yblk", 2: "iss0", 3: "mma0", 4: "epi0", 5: "iss1", 6: "mma1", 7: "epi1", 8: "iss2", 9: "mma2", 10: "epi2",
             11: "iss3", 12: "mma3", 13: "epi3", 14: "chunk0", 15: "chunk1", 16: "chunk2", 17: "chunk3", 18: "mix", 19: "end"}
else:
    names = {0: "start", 1: "turn", 2: "prologue", 3: "bar1", 4: "iss1", 5: "mma1", 6: "epi1", 7: "bar2", 8: "iss2", 9: "mma2", 10: "epi2",
             11: "bar3", 12: "iss3", 13: "mma3", 14: "epi3", 15: "bar4", 16: "iss4", 17: "mmaA", 18: "mix", 19: "xchg", 20: "end"}
n_tiles = 4 if mode == "tc" else 2
t = trace.cpu().reshape(-1, 8, 32)[:n_tiles]
t0 = int(t[t > 0].min())
last = max(names)
for s in range(2, 6):
    for tile in range(n_tiles):
        row = t[tile, s]
        base = int(row[0])
        print(f"step {40+s} tile {tile}: start @{base - t0:7d} | " + " ".join(f"{names[i]}+{int(row[i]) - int(row[i-1])}" for i in range(1, last + 1))
              + f" | total {int(row[last]) - base}"
              + (" | weight waits " + " ".join(str(int(row[21 + 2 * k]) - int(row[20 + 2 * k])) for k in range(5)) + f" | chunk-2 MMA wait {int(row[31]) - int(row[30])}" if mode == "tc" else ""))
