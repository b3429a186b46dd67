"""Per-SM bulk-copy ingest vs producers / consumers / ring depth / stage size (design aid for the weight pipeline)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
src_stages = 428
buf = torch.zeros(src_stages * 16384 + 65536, dtype=torch.uint8, device=dev)
off = (-buf.data_ptr()) % 1024
n_stages = 428 * 4
for grid in (148,):
    for n_prod, n_cons, stage_bytes, depth in ((3, 1, 16384, 3), (3, 1, 16384, 4), (3, 1, 16384, 5), (3, 1, 16384, 8), (3, 1, 16384, 10), (3, 1, 16384, 12), (3, 1, 8192, 10), (3, 1, 8192, 20), (1, 1, 16384, 3), (1, 1, 16384, 6), (2, 1, 16384, 6), (1, 2, 16384, 6), (2, 2, 16384, 6), (3, 1, 16384, 3), (3, 1, 16384, 6),
                                               (4, 1, 16384, 8), (4, 4, 16384, 8), (1, 1, 32768, 3), (2, 1, 32768, 4), (1, 1, 65536, 3)):
        out = torch.zeros(grid, dtype=torch.int64, device=dev)
        for rep in range(2):
            _lib.check(lib.pnr_ingest_bench(buf.data_ptr() + off, src_stages - 4, n_stages, depth, grid, out.data_ptr(), n_prod, n_cons, stage_bytes, _lib.stream_ptr(dev)), "ingest")
            torch.cuda.synchronize()
        cyc = out.float().mean().item()
        print(f"grid={grid:4d} prod={n_prod} cons={n_cons} stage={stage_bytes:6d} depth={depth:2d}  cycles/stage={cyc / n_stages:7.1f} B/clk/SM={stage_bytes * n_stages / cyc:6.1f}", flush=True)
