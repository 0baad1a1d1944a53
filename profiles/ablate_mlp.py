"""Where does the bf16 FaceNeRF kernel's time go?  Times inerf_mlp_fwd (fine pass of a 450x450 frame: 202 500 x 192 points)
with parts of the kernel switched off (csrc/mlp_bf16.cu, template parameter ABL; outputs are garbage, only timing counts).

    build here :  INERF_SO=$PWD/build/libinerf_abl.so INERF_EXTRA_NVCC=-DINERF_ABLATION python profiles/ablate_mlp.py --build
    run on GPU :  INERF_SO=$PWD/build/libinerf_abl.so python profiles/ablate_mlp.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

if "--build" in sys.argv:
    M.build(force=True)
    print("built", M._lib.SO_PATH)
    sys.exit(0)

dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
a = M.default_args(dim_aud=64, dim_expr=76, perturb=1.0, mlp_mode="bf16", N_samples=64, N_importance=128, near=S.NEAR, far=S.FAR)
net = M.Network(450, 450, cam["focal"], S.NEAR, S.FAR, 1 << 20, None, 64, 128, args=a)
torch.manual_seed(1)
net.apply(M.init_weights)
net = net.to(dev).eval()
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
names = {0: "full kernel", 1: "epilogue off", 2: "weight streaming off", 4: "sincosf off", 3: "epilogue+weights off",
         5: "epilogue+sincosf off", 6: "weights+sincosf off", 7: "all three off (MMA issue + barrier protocol only)", 16: "epilogue: no bias/ReLU/convert math", 24: "epilogue: stores only", 8: "epilogue: no TMEM loads"}
pts = rays.shape[0] * 192
for abl in (0, 1, 2, 4, 7):
    os.environ["INERF_ABL"] = str(abl)
    with torch.no_grad():
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
    cyc = ms * 1e-3 * 1.83e9 / (pts / 256 / 148)       # 1.83 GHz: profiles/mlp_clock.py
    print(f"ABL={abl} {names[abl]:55s} {ms:8.3f} ms   {pts * 1121280 / ms / 1e9:7.1f} TFLOP/s   ~{cyc / 1e3:6.1f} kcyc / 256-point iteration")
