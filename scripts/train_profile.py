import sys, os, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import helpers as H
from pixel_nerf_yolo_b200.render import NeRFRenderer
dev = torch.device("cuda", 0)
scene = H.make_scene_dict(num_objs=4, num_views=3, feat=64, size=128)
net = H.build_net(scene, precision="bf16", train=True); net.train_precision = os.environ.get("TP", "tf32")
lat = scene["latent"].to(dev).clone().requires_grad_(True); net.encoder.set_latent(lat)
r = NeRFRenderer(64, 32, 16, white_bkgd=True).train().to(dev)
rays = H.rays_subset(4, 128, seed=1).to(dev); gt = torch.rand(4, 128, 3, device=dev)
for it in range(2):
    for p in net.parameters(): p.grad = None
    res = r(net, rays); loss = ((res.coarse.rgb - gt) ** 2).mean() + ((res.fine.rgb - gt) ** 2).mean(); loss.backward()
torch.cuda.synchronize(); print("ok", loss.item())
