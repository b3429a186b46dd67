"""Per-tensor gradient error of the tcgen05 training path vs torch autograd over the oracle (diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import pixel_nerf_yolo_b200.synth as synth
from oracle import pixelnerf_oracle as O
num_objs, num_views, P = 1, 3, 300
scene = H.make_scene_dict(num_objs=num_objs, num_views=num_views, feat=16)
g = torch.Generator().manual_seed(P)
xyz = (torch.rand(num_objs, P, 3, generator=g) - 0.5) * 0.8
dirs = torch.nn.functional.normalize(torch.randn(num_objs, P, 3, generator=g), dim=-1)
gout = torch.randn(num_objs, P, 4, generator=g)
sc = H.oracle_scene(scene)
sc.latent = sc.latent.clone().requires_grad_(True)
mc = {k: v.clone().requires_grad_(True) for k, v in synth.mlp_state(1).items()}
xo = xyz.clone().requires_grad_(True)
EMU = bool(int(os.environ.get("EMU", "1")))
ref = O.field_forward(sc, mc, xo, dirs, emulate_bf16=EMU)
(ref * gout).sum().backward()
for prec in ("bf16",):
    net = H.build_net(scene, precision="bf16", train=True)
    lat = scene["latent"].cuda().clone().requires_grad_(True)
    net.encoder.set_latent(lat)
    net.train_precision = prec
    xc = xyz.cuda().requires_grad_(True)
    out = net(xc, coarse=True, viewdirs=dirs.cuda())
    (out * gout.cuda()).sum().backward()
    print(prec, "out err", (out.detach().cpu() - ref.detach()).abs().max().item())
    for name, p in net.mlp_coarse.named_parameters():
        r = mc[name].grad
        err = (p.grad.cpu() - r).abs().max().item()
        print(f"  {prec} {name:24s} err/scale {err / r.abs().max().item():.3e}  rel-norm {((p.grad.cpu() - r).norm() / r.norm()).item():.3e}")
    print(f"  {prec} latent err/scale {(lat.grad.cpu() - sc.latent.grad).abs().max().item() / sc.latent.grad.abs().max().item():.3e}",
          f"xyz err/scale {(xc.grad.cpu() - xo.grad).abs().max().item() / xo.grad.abs().max().item():.3e}")
