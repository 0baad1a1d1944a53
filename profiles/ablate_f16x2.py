"""Where the fp16x2 FaceNeRF kernel spends its time: the fine pass of a 450x450 frame (202 500 x 192 points) with parts switched off
(INERF_F16X2_ABL bits: 1 = weight ring never reloaded, 2 = epilogue without TMEM loads / conversion / stores, 4 = gamma(p) without
sincosf).  Outputs are garbage in the ablated runs; only the time is meaningful.   python profiles/ablate_f16x2.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
_, net, fr, cam = bench.build_network("fp16x2", dev)
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
with torch.no_grad():
    rays = ops.get_rays_packed(450, 450, net.focal, fr["pose"].to(dev)[:3, :4], net.near, net.far)
    z = torch.sort(torch.rand(202500, 192, device=dev) * 0.6 + net.near, -1)[0]
    for abl in (0, 1, 2, 4, 3, 7):
        os.environ["INERF_F16X2_ABL"] = str(abl)
        for _ in range(2):
            net.face_nerf_fine.query(rays, z, aud, expr, lat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            net.face_nerf_fine.query(rays, z, aud, expr, lat)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"abl={abl}: {ms:8.2f} ms per fine pass  ({3 * 202500 * 192 * 1121280 / ms / 1e9:7.0f} TFLOP/s of MMA issued)")
