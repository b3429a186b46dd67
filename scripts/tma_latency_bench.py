"""Latency of the weight stream's loads in isolation (pnr_tma_latency_bench, csrc/lab.cu): cycles from issue to completion of k 16 KiB
loads per CTA, 1 pair (idle chip) vs 74 pairs (every SM streaming), for the three load forms."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
src_stages = 436                     # 6.8 MB: the size of one packed MLP
buf = torch.zeros(src_stages * 16384 + 65536, dtype=torch.uint8, device=dev)
off = (-buf.data_ptr()) % 1024
iters = 2000
names = {0: "bulk 1-D, own barrier", 1: "tensor-map box, own barrier", 2: "cta_group::2 boxes of both CTAs -> leader's barrier"}
for pairs in (1, 74):
    for mode in (0, 1, 2):
        for k in (1, 2, 3, 5):
            out = torch.zeros(pairs * 2, dtype=torch.int64, device=dev)
            for rep in range(2):
                _lib.check(lib.pnr_tma_latency_bench(buf.data_ptr() + off, src_stages, iters, mode, k, pairs, out.data_ptr(), _lib.stream_ptr(dev)), "tma_latency")
                torch.cuda.synchronize()
            v = out.float()
            v = v[v > 0]
            print(f"pairs={pairs:3d} {names[mode]:52s} k={k}  cycles issue->all landed = {v.mean().item() / iters:7.0f}", flush=True)
