"""Diagnostic for the MN-major weight-gradient GEMM: one-hot probes reveal which (row, feature) of the operands land where."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixel_nerf_yolo_b200 import _lib
lib = _lib.load()
M, N, K = 64, 256, 256
st = _lib.stream_ptr(torch.device("cuda", 0))
ws = torch.empty(lib.pnr_lab_gemm_workspace_bytes(M, N, K) + 1024, dtype=torch.uint8, device="cuda")
def run(dY, X):
    dW = torch.zeros(N, K, device="cuda")
    dYd, Xd = dY.cuda(), X.cuda()
    _lib.check(lib.pnr_lab_wgrad(dYd.data_ptr(), Xd.data_ptr(), dW.data_ptr(), M, N, K, ws.data_ptr(), ws.numel(), st), "wgrad")
    torch.cuda.synchronize()
    return dW.cpu()
print("env", {k: v for k, v in os.environ.items() if k.startswith("PNR_WGRAD")})
for (m0, n0, k0) in [(0, 0, 0), (0, 1, 0), (0, 8, 0), (0, 64, 0), (0, 128, 0), (0, 0, 1), (0, 0, 9), (0, 0, 70), (0, 0, 130), (1, 0, 0), (7, 3, 5), (8, 3, 5), (17, 70, 200), (63, 255, 255)]:
    dY = torch.zeros(M, N); X = torch.zeros(M, K)
    dY[m0, n0] = 1.0; X[m0, k0] = 1.0
    out = run(dY, X)
    nz = out.nonzero().tolist()
    print("probe m,n,k =", (m0, n0, k0), "-> nonzero at", nz[:6], "values", [round(out[i, j].item(), 3) for i, j in nz[:6]])
g = torch.Generator().manual_seed(0)
dY, X = torch.randn(M, N, generator=g), torch.randn(M, K, generator=g)
ref = dY.bfloat16().double().t() @ X.bfloat16().double()
out = run(dY, X).double()
print("random: max err", (out - ref).abs().max().item(), "ref max", ref.abs().max().item())
