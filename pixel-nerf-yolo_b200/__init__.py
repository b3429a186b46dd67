"""B200-native rendering hot path of pixelNeRF-YOLO (drop-in for NeRFRenderer / PixelNeRFNet).

Host code is PyTorch (device memory, streams, torch.distributed); all arithmetic on the path runs in
hand-written sm_100a CUDA kernels reached through the C-ABI library declared in
``include/pixelnerf_b200.h``.  There is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
