"""Baseline harness: runs the UNMODIFIED reference (kofinandi/pixel-nerf-yolo) for bench.py's reference arm, `cpu_baseline` and
`gpu_eager_baseline`.  `baseline/_ref/` (git-ignored, shipped to the GPU box by gpurun) holds a copy of the reference's own
`src/{model,render,util}` and `conf/` made by `baseline/make_ref.py`; nothing here is imported by the product package."""
