"""Debug tool: phase timeline of the tensor-core flow kernel (CTA 0, both tiles, Mobius steps 40..47).
Build with  RNF_NVCC_EXTRA=-DRNF_TC_TRACE=1 python -m rotationnormflow_b200.build --force  and run on the GPU box."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rotationnormflow_b200 import _cabi, grid as rgrid
from rotationnormflow_b200.flow import _program

cfg, flow = bench.build_flow("symsol", feature_dim=2048)
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
    # worker view (warp quarter 0 of each tile): gemmN = hand-over -> woken up after GEMM N (issue by the service warp + MMAs),
    # epiN = epilogue of GEMM N incl. both half hand-overs, chunkN = wait + drain of fc_last chunk N (incl. the arithmetic before it)
    names = {0: "start", 1: "yblk", 3: "gemm0", 4: "epi0", 6: "gemm1", 7: "epi1", 9: "gemm2", 10: "epi2",
             12: "gemm3", 13: "epi3", 14: "chunk0", 15: "chunk1", 16: "chunk2", 17: "chunk3", 18: "mix", 19: "end"}
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
        idx = sorted(names)
        print(f"step {40+s} tile {tile}: start @{base - t0:7d} | " + " ".join(f"{names[i]}+{int(row[i]) - int(row[j])}" for j, i in zip(idx[:-1], idx[1:]))
              + f" | total {int(row[last]) - base}")
