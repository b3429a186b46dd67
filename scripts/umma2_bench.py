"""The CTA-pair MMA (tcgen05.mma.cta_group::2, M = 256) in isolation: cycles per MMA vs N, commit frequency, A-ring depth and
concurrent shared-memory traffic (design aid; pnr_umma2_bench in csrc/lab.cu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
iters = 4000
pairs = 74
src = torch.zeros(pairs * 4 << 20, dtype=torch.uint8, device=dev)
for N in (64, 128, 256):
    for commit_every in (4,):
        for a_slots in (5,):
            for bg in (0, 1, 5, 3, 7):
                out = torch.zeros(pairs * 10, dtype=torch.int64, device=dev)
                for rep in range(2):
                    _lib.check(lib.pnr_umma2_bench(N, iters, commit_every, a_slots, bg, pairs, src.data_ptr(), out.data_ptr(), _lib.stream_ptr(dev)), "umma2_bench")
                    torch.cuda.synchronize()
                total = out[:pairs].float().mean().item()
                cyc = total / (iters * 8)
                copies = out[pairs:pairs + pairs * 8].float().reshape(pairs * 2, 4).sum(1).mean().item()      # per CTA (four loader warps)
                print(f"N={N:3d} commit_every={commit_every} a_slots={a_slots} bg={bg}  cycles/MMA={cyc:7.1f}  MAC/clk/SM={128 * N * 16 / cyc:7.0f}  ({100 * 128 * N * 16 / cyc / 4096:5.1f} % of 4096)"
                      + (f"  concurrent bulk copies into smem: {copies * 8192 / total:6.1f} B/clk/SM" if bg & 1 else "") + ("   [no MMAs issued: baseline of the background traffic]" if bg & 4 else ""), flush=True)
