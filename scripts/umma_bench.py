"""tcgen05.mma cost vs shape and commit frequency (design aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
iters = 2000
grid = 148
for M in (128,):
    for N in (32, 64, 128, 256):
        for commit_every in (0, 4, 104):
            out = torch.zeros(grid, dtype=torch.int64, device=dev)
            for rep in range(2):
                _lib.check(lib.pnr_umma_bench(M, N, iters, 8, grid, out.data_ptr(), commit_every, _lib.stream_ptr(dev)), "umma_bench")
                torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 8)
            print(f"M={M:3d} N={N:3d} commit_every={commit_every}  cycles/MMA={cyc:7.1f}  MAC/clk={M * N * 16 / cyc:7.0f}")
