"""One fine-pass launch of the experimental cta_group::2 kernel (INERF_MLP_V2=1) for ncu source-level capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["INERF_MLP_V2"] = "1"
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S
dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
net = M.FaceNeRF(dim_aud=64, dim_latent=32, dim_expr=76, mlp_mode="bf16").to(dev)
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)[:40000]
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
with torch.no_grad():
    for _ in range(3):
        net.query(rays, z, fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev))
torch.cuda.synchronize()
print("ok")
