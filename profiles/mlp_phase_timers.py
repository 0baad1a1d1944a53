"""Phase timers of the PRODUCTION bf16 FaceNeRF kernel (two clock reads per epilogue half in one warp, and around the issuer's waits).
    build here :  INERF_SO=$PWD/build/libinerf_ph.so INERF_EXTRA_NVCC=-DINERF_PHASE_TIMERS python profiles/mlp_phase_timers.py --build
    run on GPU :  INERF_SO=$PWD/build/libinerf_ph.so python profiles/mlp_phase_timers.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

if "--build" in sys.argv:
    M.build(force=True)
    print("built", M._lib.SO_PATH)
    sys.exit(0)

dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
net = M.FaceNeRF(dim_aud=64, dim_latent=32, dim_expr=76, mlp_mode="bf16")
torch.manual_seed(1)
net.apply(M.init_weights)
net = net.to(dev)
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
TRAIN = "--train" in sys.argv          # the activation-saving forward on an N_rand = 3072 batch instead of the inference kernel on a frame
if TRAIN:
    rays, z = rays[:3072].contiguous(), z[:3072].contiguous()
    params = [p.detach() for p in net.kernel_params()]
    cond = ops.fold_cond(net._dims, params, aud, expr, lat)
    packed = net.packed_weights(net.kernel_params())
    run = lambda: ops.mlp_fwd_train_bf16(net._dims, params, packed, cond, rays, z)
else:
    run = lambda: net.query(rays, z, aud, expr, lat)
with torch.no_grad():
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
buf = np.zeros((148, 16), np.uint64)
rc = M.lib().inerf_debug_phase_timers(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
b = buf.astype(np.float64)
b = b[b[:, 1] > 0]            # CTAs that issued (pair build: the leaders)
it = b[:, 1]
per = lambda col: float((b[:, col] / it).mean())
print(f"fine pass {e0.elapsed_time(e1):.3f} ms, {it.mean():.1f} iterations per CTA")
print(f"issuer, cycles per 256-point iteration: total {per(0):.0f} | waits: E0 of L2-4,6,7 {per(2):.0f} (x6 layers), E1 of the same {per(3):.0f}, "
      f"other layers' events {per(4):.0f}, weight stages {per(5):.0f}")
for name, o in (("warp 4 (slot 0, rows 0-31)", 6), ("warp 11 (slot 1, rows 96-127)", 11)):
    print(f"epilogue {name}, L1..L7, cycles per layer: h0 wait for C0 {per(o) / 7:.0f}, h0 work {per(o + 1) / 7:.0f} (of which waiting for C1 {per(o + 4) / 7:.0f}), "
          f"h1 wait for C2 {per(o + 2) / 7:.0f}, h1 work {per(o + 3) / 7:.0f}")
