"""Small-shape pass over the kernels added in round 2, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tests/native/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

dev = "cuda:0"
cam = S.camera()
c2w = cam["c2w"].to(dev)
rays = ops.get_rays_range(450, 450, cam["focal"], c2w, S.NEAR, S.FAR, 101337, 133)
st = torch.tensor([5, 0], dtype=torch.int64, device=dev)
for s in (64, 45):
    z = ops.sample_coarse_rng(rays, s, st)
    w = torch.rand(133, s, device=dev)
    for n_imp in (128, 37):
        ops.importance_sample_rng(z, w, n_imp, st, want_samples=True)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
ops.flag_nonfinite([z, w, torch.zeros(0, device=dev)], flag)
net = M.Network(450, 450, cam["focal"], S.NEAR, S.FAR, 8192, None, 64, 128, args=M.default_args(dim_aud=64, dim_expr=76, perturb=1.0, nosmo_iters=0))
torch.manual_seed(0)
net.apply(M.init_weights)
net = net.to(dev)
fr = S.frame_inputs(0)
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
bc = torch.rand(133, 3, device=dev)
with torch.no_grad():
    for mode in ("fp16x2", "bf16", "fp32"):
        net.set_mlp_mode(mode)
        r = net.render_rays(rays, bc, aud, None, lat, expr, perturb=1.0)
        assert bool(torch.isfinite(r["rgb_map"]).all()), mode
auds = torch.randn(12, 16, 29, device=dev)
net.set_mlp_mode("bf16")
net.train()
a = net.audio_feature(auds, 5, 12, global_step=0)
r = net.render_rays(rays, bc, a, None, lat.clone().requires_grad_(True), expr, perturb=1.0)
(r["rgb_map"].sum() + r["rgb0"].sum()).backward()
assert net.aud_net.encoder_conv[0].weight.grad is not None
torch.cuda.synchronize()
print("sanitize_small OK")
