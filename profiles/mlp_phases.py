"""Where do the issuer and the epilogue warps of the bf16 FaceNeRF kernel wait?  Runs the TRACE build (clock64 phase timers in
csrc/mlp_bf16.cu) on the fine pass of a frame band and prints cycles per 256-point iteration.
    python profiles/mlp_phases.py [n_rays]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
cam, fr = S.camera(), S.frame_inputs(0)
net = M.FaceNeRF(dim_aud=64, dim_latent=32, dim_expr=76, mlp_mode="bf16")
torch.manual_seed(1)
net.apply(M.init_weights)
net = net.to(dev)
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)[:n].contiguous()
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(n, 192, device=dev), -1)[0].contiguous()
params = [p.detach() for p in net.kernel_params()]
cond = ops.fold_cond(net._dims, params, fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev))
packed = net.packed_weights(net.kernel_params())
for _ in range(2):
    raw, _ = ops.mlp_fwd_trace(M._lib.INERF_MLP_BF16, net._dims, params, packed, cond, rays, z)
torch.cuda.synchronize()
t = ops.mlp_fwd_trace.timing.cpu()
iss, epi = t[:148], t[148:296]
iss, epi = iss[iss[:, 4] > 0], epi[epi[:, 4] > 0]
it = iss[:, 4]
print(f"{n} rays x 192 samples, {float(it.mean()):.1f} iterations per CTA (trace build: clock64 reads add overhead)")
print("issuer, cycles per iteration: total %.0f | waits: epilogue events %.0f, PE %.0f, weight stages %.0f | issuing + other %.0f"
      % ((iss[:, 0] / it).mean(), (iss[:, 1] / it).mean(), (iss[:, 2] / it).mean(), (iss[:, 3] / it).mean(),
         ((iss[:, 0] - iss[:, 1] - iss[:, 2] - iss[:, 3]) / it).mean()))
lw = ops.mlp_fwd_trace.layer_waits.cpu()[:len(it)]
print("issuer wait for epilogue events by layer (L0..L7, V0..V2): " + " ".join("%.0f" % float((lw[:, j] / it).mean()) for j in range(11)))
ie = epi[:, 4]
print("epilogue warp 4, cycles per iteration: wait for MMAs %.0f | TMEM loads + convert %.0f | wait h1 reads (cbar1) %.0f | stores + fence + arrive %.0f"
      % tuple(float((epi[:, j] / ie).mean()) for j in range(4)))
