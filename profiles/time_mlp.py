"""Time the FaceNeRF forward kernel alone on the passes of a 450x450 frame (202 500 x 64 and x 192 points), CUDA events, resident inputs.
python profiles/time_mlp.py [bf16|fp16x2|fp32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from ideal_nerf_b200 import ops

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_, net, fr, cam = bench.build_network(mode, dev)
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
with torch.no_grad():
    rays = ops.get_rays_packed(450, 450, net.focal, fr["pose"].to(dev)[:3, :4], net.near, net.far)
    for s, fn in ((64, net.face_nerf_coarse), (192, net.face_nerf_fine)):
        z = torch.sort(torch.rand(202500, s, device=dev) * 0.6 + net.near, -1)[0]
        reps = 10 if mode != "fp32" else 1
        for _ in range(3 if mode != "fp32" else 1):
            fn.query(rays, z, aud, expr, lat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn.query(rays, z, aud, expr, lat)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{mode} s={s}: {ms:8.3f} ms  {202500 * s * 1121280 / ms / 1e9:7.1f} TFLOP/s algorithmic")
