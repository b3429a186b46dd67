"""DSMEM one-way transfer cost between the CTAs of a pair (design aid for the epilogue exchange)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
iters = 2000
for mode, warps in ((0, 1), (0, 4), (0, 8), (1, 1)):
    for nbytes in (2048, 8192, 16384, 32768):
        out = torch.zeros(2, dtype=torch.int64, device=dev)
        for rep in range(2):
            _lib.check(lib.pnr_dsmem_bench(mode, nbytes, iters, warps, out.data_ptr(), _lib.stream_ptr(dev)), "dsmem")
            torch.cuda.synchronize()
        cyc = out.float().max().item() / iters
        print(f"mode={'st.v4+fence' if mode == 0 else 'cp.async.bulk'} warps={warps} bytes={nbytes:6d}  cycles/transfer={cyc:8.1f}  B/clk={nbytes / cyc:6.1f}")
