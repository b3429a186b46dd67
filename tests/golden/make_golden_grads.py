"""Golden GRADIENTS of the training step (BASELINE config 3 shape), produced by the UNMODIFIED reference's autograd.

Run in the build container only:   python tests/golden/make_golden_grads.py

Same shims, synthetic weights, feature maps, rays and noise as make_golden.py (case "sb2": 2 objects x 24 rays,
3 source views).  Loss = MSE(coarse.rgb, gt) + MSE(fine.rgb, gt) as in train/trainlib/PixelNerfTrainer.py:141-157
(lambda_coarse = lambda_fine = 1).  Stored: the loss, every parameter gradient's norm and sum, small gradients in
full, a fixed slice of the big ones, and the gradient that reaches the encoder output (norm + a channel slice).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402


def grad_summary(out, prefix, named_grads):
    for name, g in named_grads:
        g = g.detach()
        out[f"{prefix}{name}.norm"] = np.float64(g.double().norm().item())
        out[f"{prefix}{name}.sum"] = np.float64(g.double().sum().item())
        if g.numel() <= 512 * 42:
            out[f"{prefix}{name}.full"] = g.numpy()
        else:
            out[f"{prefix}{name}.slice"] = g[:8, :96].numpy().copy()


def main():
    MG._install_shims()
    import pixel_nerf_yolo_b200.synth as synth
    from render import NeRFRenderer

    torch.set_num_threads(8)
    num_objs, size, feat, nrays = 2, 128, 16, 24
    scene = synth.scene_config1(seed=5, num_views=3, C=512, size=size, feat=feat, num_objs=num_objs)
    net = MG.build_reference_net(synth, 1, 2, scene, num_objs).train()
    lat = scene["latent"].clone().requires_grad_(True)
    net.encoder.latent = lat
    all_rays = torch.cat([synth.target_rays(size, 15.0 + 20 * s, -10.0) for s in range(num_objs)])
    pick = torch.from_numpy(np.random.default_rng(7).choice(size * size, nrays, replace=False)).long()
    rays = all_rays[:, pick]
    noise = MG.np_noise(9, num_objs * nrays)
    gt = torch.from_numpy(np.random.default_rng(13).random((num_objs, nrays, 3), dtype=np.float32))
    renderer = NeRFRenderer.from_conf(MG._Conf(MG.RENDER_CONF), eval_batch_size=50000).train()
    with MG._NoisePatch(noise):
        res = renderer(net, rays, want_weights=True)
    loss = torch.nn.functional.mse_loss(res.coarse.rgb, gt) + torch.nn.functional.mse_loss(res.fine.rgb, gt)
    loss.backward()
    out = {"ray_idx": pick.numpy(), "gt": gt.numpy(), "loss": np.float64(loss.item()),
           "coarse_rgb": res.coarse.rgb.detach().numpy(), "fine_rgb": res.fine.rgb.detach().numpy()}
    grad_summary(out, "coarse.", [(n, p.grad) for n, p in net.mlp_coarse.named_parameters()])
    grad_summary(out, "fine.", [(n, p.grad) for n, p in net.mlp_fine.named_parameters()])
    out["latent.norm"] = np.float64(lat.grad.double().norm().item())
    out["latent.sum"] = np.float64(lat.grad.double().sum().item())
    out["latent.slice"] = lat.grad[:, :16].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "reference_grads.npz"), **out)
    print("wrote reference_grads.npz; loss", loss.item(), "latent grad norm", out["latent.norm"],
          "coarse lin_out.weight grad norm", out["coarse.lin_out.weight.norm"],
          "coarse lin_in.weight grad norm", out["coarse.lin_in.weight.norm"])


if __name__ == "__main__":
    main()
