#!/usr/bin/env python
"""Headline benchmark: rays/s of a 3-view 128x128 pixelNeRF render (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one full target image (16 384 rays, 64 coarse + 96 fine samples per ray, 3 source views,
resnet34-sized 512-channel feature maps, random-init weights) through NeRFRenderer.forward.
`value` is device-timed with the rays already in HBM; `e2e` is the same render through the public API
(`renderer.bind_parallel(net, ..., simple_output=True)(rays)`) with the rays in pinned host memory and the
pixels copied back to the host inside the timed region.  Under torchrun (N>1) every rank renders its own
target view (weak scaling: N views for N GPUs) and the packed (rgb, depth) outputs are all-gathered once
per step with NCCL inside the timed region; the time is the max over ranks.

`--impl reference` times the reference's CPU implementation of the same path on the host cores.  The
reference is pure Python/PyTorch and cannot travel to the GPU box, so this arm runs the oracle port
(oracle/pixelnerf_oracle.py, pinned to the reference's outputs by tests/test_oracle_golden.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIZE, NS, KC, KF, KFD = 128, 3, 64, 32, 16
H, D_IN, C_LAT, D_OUT = 512, 42, 512, 4
POINTS_PER_RAY = KC + (KC + KF)                                    # 160
# MACs per (point, view) row and per point, FLOPs = 2 * MACs
FLOP_PER_POINT = 2 * (NS * (D_IN * H + 3 * (C_LAT * H + 2 * H * H)) + 2 * 2 * H * H + H * D_OUT)
FLOP_PER_RAY = FLOP_PER_POINT * POINTS_PER_RAY                     # 2.6218 GFLOP
WORKLOAD = "config2: 3-view 128x128 full-image render, 16384 rays, 64 coarse + 32 fine (16 importance + 16 depth), bf16 ResnetFC"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d.get("bf16_tflops_sustained"), bf16_burst=d.get("bf16_tflops"),
                    hbm=d.get("hbm_gbs"), source="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the field kernel per launch (mean of the coarse and the fine launch
    of this workload), copied from the committed `ncu --set full` capture; None if no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))["field_pair_kernel"]["dram_bytes_per_launch"]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (recipe of B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_inputs(rank, device):
    """Scene (3 source views through the random-init resnet34 encoder), weights and this rank's target view."""
    import pixel_nerf_yolo_b200.synth as synth
    from pixel_nerf_yolo_b200.conf import ConfigTree
    from pixel_nerf_yolo_b200.model import make_model
    from pixel_nerf_yolo_b200.render import NeRFRenderer
    conf = {
        "use_encoder": True, "use_global_encoder": False, "use_xyz": True, "use_code": True,
        "code": {"num_freqs": 6, "freq_factor": 1.5, "include_input": True}, "use_viewdirs": True, "use_code_viewdirs": False,
        "mlp_coarse": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4, "combine_layer": 3, "combine_type": "average"},
        "mlp_fine": {"type": "resnet", "n_blocks": 5, "d_hidden": 512, "d_out": 4, "combine_layer": 3, "combine_type": "average"},
        "encoder": {"backbone": "resnet34", "pretrained": False, "num_layers": 4, "index_padding": "zeros"},
    }
    torch.manual_seed(0)
    net = make_model(ConfigTree.from_dict(conf)).eval()
    net.mlp_coarse.load_state_dict(synth.mlp_state(1))
    net.mlp_fine.load_state_dict(synth.mlp_state(2))
    net = net.to(device)
    scene = synth.scene_config1(seed=0, num_views=NS, C=C_LAT, size=SIZE)
    g = torch.Generator().manual_seed(0)
    images = (torch.rand(1, NS, 3, SIZE, SIZE, generator=g) * 2 - 1).to(device)
    with torch.no_grad():
        net.encode(images, scene["poses"].to(device), scene["focal"].to(device))         # real resnet34 trunk, random init
        lat = net.encoder.latent
        net.encoder.set_latent(lat / lat.std())                                          # O(1) features
    renderer = NeRFRenderer(KC, KF, KFD, depth_std=0.01, white_bkgd=True, eval_batch_size=50000).eval().to(device)
    rays = synth.target_rays(SIZE, theta=15.0 + 15.0 * rank, phi=-10.0)                  # (1, 16384, 8), host
    return net, renderer, rays, scene


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the pixelnerf_b200 path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from pixel_nerf_yolo_b200.dist import all_gather_rows
    net, renderer, rays_host, scene = build_inputs(rank, device)
    render_par = renderer.bind_parallel(net, None, simple_output=True).eval()
    rays_dev = rays_host.to(device)
    rays_pinned = rays_host.pin_memory()
    n_rays = rays_host.shape[1]
    out_pinned = torch.empty(n_rays, 4).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)                      # > 126 MB L2
    bounds = [(i * n_rays, (i + 1) * n_rays) for i in range(world)]

    # time the dominant kernel (fused field kernel) live, on the launching stream: the single-call render records CUDA events
    # around its two field-kernel launches (pnr_render_args.field_events)
    field_ms = []

    def arm_field_events():
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        renderer.field_events = evs
        field_ms.append((evs[0], evs[1]))
        field_ms.append((evs[2], evs[3]))

    def step_device():
        arm_field_events()
        rgb, depth = render_par(rays_dev)
        packed = torch.cat((rgb[0], depth[0].unsqueeze(-1)), dim=-1)                      # (B, 4) = 16 B/ray
        if world > 1:
            packed = all_gather_rows(packed, bounds)                                     # the path's one collective
        return packed

    def step_e2e():
        r = rays_pinned.to(device, non_blocking=True)
        rgb, depth = render_par(r)
        packed = torch.cat((rgb[0], depth[0].unsqueeze(-1)), dim=-1)
        out_pinned.copy_(packed, non_blocking=True)
        return packed

    def timed_loop(fn, steps):
        total = 0.0
        for _ in range(steps):
            flush.fill_(1)                                                                # evict L2 between steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step_device()
            step_e2e()
        field_ms.clear()
        sampler = ClockSampler(local)
        sync_all()
        if rank == 0:
            sampler.start()
        ms_dev = timed_loop(step_device, args.steps)
        sync_all()
        field_pairs = list(field_ms)
        renderer.field_events = None
        launches = renderer.last_launches
        ms_e2e = timed_loop(step_e2e, args.steps)
        sync_all()
        clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    field_total_ms = sum(a.elapsed_time(b) for a, b in field_pairs)
    n_field = len(field_pairs)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(scene, net, sample_rays=args.cpu_rays)
    if rank == 0:
        pk = peaks()
        total_rays = world * n_rays * args.steps
        flop_launch = FLOP_PER_RAY * n_rays * args.steps / max(n_field, 1)               # algorithmic FLOPs per field launch
        achieved = (flop_launch / (field_total_ms / max(n_field, 1) * 1e-3)) / 1e12 if n_field else None
        line = {
            "metric": "rays/sec (samples/sec) 3-view render at 1/2/4/8 B200 vs host-CPU ref", "value": total_rays / (ms_dev * 1e-3),
            "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "samples_per_sec": total_rays * POINTS_PER_RAY / (ms_dev * 1e-3),
            "config": {"workload": WORKLOAD, "rays_per_gpu": n_rays, "source_views": NS, "samples_per_ray": POINTS_PER_RAY,
                       "feature_maps": "3x512x64x64 bf16 channels-last (random-init resnet34 on synthetic images, /std)",
                       "parallelism": f"rays sharded: {world} target view(s), one per GPU, one NCCL all-gather of (rgb,depth) per step",
                       "l2": "flushed between timed steps (256 MiB write); feature maps+weights are re-read from HBM each step"},
            "e2e": {"value": total_rays / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": n_rays * 8 * 4,
                    "d2h_bytes_per_step": n_rays * 4 * 4},
            "gpu_launches": launches * args.steps * 2,   # device loop + e2e loop
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / pk["bf16_sustained"]) if achieved else None, "traffic": ncu_traffic(),
                         "kernel": "pair::field_pair_kernel<3> (fused projection + gather + PE + ResnetFC, tcgen05 cta_group::2), 2 launches per step",
                         "peak_kind": f"{pk['source']} sustained bf16 (kernel timed inside a long step); burst {pk['bf16_burst']}",
                         "frac_of_burst": (achieved / pk["bf16_burst"]) if achieved else None,
                         "kernel_ms_per_step": field_total_ms / args.steps},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def oracle_inputs(scene, latent_cpu):
    from oracle import pixelnerf_oracle as O
    import pixel_nerf_yolo_b200.synth as synth
    sc = O.encode_cameras(latent_cpu, scene["poses"], scene["focal"], scene["image_wh"])
    return O, sc, synth.mlp_state(1), synth.mlp_state(2)


def time_oracle(O, sc, mc, mf, rays, seed=0, chunk=1024):
    """Wall time of the oracle render of `rays`, in ray chunks of `chunk` (the reference chunks its model calls too,
    nerf.py:209-222; rays are independent, so the arithmetic per ray is the same) to bound host memory."""
    total = 0.0
    for s in range(0, rays.shape[1], chunk):
        r = rays[:, s:s + chunk]
        noise = O.RenderNoise.draw(r.shape[1], KC, KF, KFD, generator=torch.Generator().manual_seed(seed + s))
        t0 = time.perf_counter()
        O.render(sc, mc, mf, r, noise, n_coarse=KC, n_fine=KF, n_fine_depth=KFD, white_bkgd=True)
        total += time.perf_counter() - t0
    return total


def cpu_baseline(scene, net, sample_rays):
    """The oracle port on the host cores, bounded sample of the same workload."""
    import pixel_nerf_yolo_b200.synth as synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    O, sc, mc, mf = oracle_inputs(scene, net.encoder.latent.detach().float().cpu())
    rays = synth.target_rays(SIZE)[:, :: max(1, (SIZE * SIZE) // sample_rays)][:, :sample_rays]
    time_oracle(O, sc, mc, mf, rays[:, :64])                                              # warm-up
    dt = time_oracle(O, sc, mc, mf, rays)
    return {"value": rays.shape[1] / dt, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{rays.shape[1]} rays of the same 128x128 view (every {(SIZE * SIZE) // sample_rays}th ray), {dt:.1f} s, torch {torch.__version__} fp32, {cores} threads"}


def run_reference(args):
    """Reference arm: the reference's CPU path (oracle port) on all host threads; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import pixel_nerf_yolo_b200.synth as synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    scene = synth.scene_config1(seed=0, num_views=NS, C=C_LAT, size=SIZE)
    O, sc, mc, mf = oracle_inputs(scene, scene["latent"])
    n = args.cpu_rays
    rays = synth.target_rays(SIZE)[:, :: max(1, (SIZE * SIZE) // n)][:, :n]
    for _ in range(args.warmup):
        time_oracle(O, sc, mc, mf, rays[:, :64])
    dt = sum(time_oracle(O, sc, mc, mf, rays, seed=i) for i in range(args.steps))
    val = n * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec (samples/sec) 3-view render at 1/2/4/8 B200 vs host-CPU ref", "value": val,
        "unit": "rays/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "step": f"bounded sample: {n} rays of the 16384-ray image per step"},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n} rays x {args.steps} steps, oracle port of the reference path, torch {torch.__version__} fp32"},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=None,
                    help="rays in the bounded CPU sample (default 4096 for cpu_baseline = 10-15 s of CPU work, 512 per step for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.cpu_rays is None:
        args.cpu_rays = 4096 if args.impl == "ours" else 512
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
