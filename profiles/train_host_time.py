import sys, time, json, os
sys.path.insert(0, '/root/repo')
import torch, bench
import ideal_nerf_b200 as M
# replicate train_step_bench but time the host side of each step
from ideal_nerf_b200 import synthetic as S, ops, train as T
dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
a = M.default_args(dim_aud=64, dim_expr=76, perturb=1.0, mlp_mode="bf16", N_samples=64, N_importance=128)
net = M.Network(450, 450, cam["focal"], S.NEAR, S.FAR, 8192, None, 64, 128, args=a)
torch.manual_seed(4321); net.apply(M.init_weights); net = net.to(dev).train()
g = torch.Generator().manual_seed(5)
idx = torch.randperm(202500, generator=g)[:3072].to(dev)
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)[idx].contiguous()
bc, tgt = fr["bc_rgb"].to(dev)[idx].contiguous(), torch.rand(3072, 3, generator=g).to(dev)
aud, expr = fr["aud"].to(dev), fr["expr"].to(dev)
lat = torch.ones(32, device=dev, requires_grad=True)
params = list(net.parameters()) + [lat]
opt = torch.optim.Adam(params, lr=3e-4, fused=True)
def step():
    opt.zero_grad(set_to_none=True)
    r = net.render_rays(rays, bc, aud, None, lat, expr)
    loss = T.head_loss(r, tgt, lat, 0.0005)[0]
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
hs = []
for _ in range(20):
    h0 = time.perf_counter(); step(); hs.append(time.perf_counter() - h0)
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host time per step {1e3*t_host/20:.3f} ms (median {1e3*sorted(hs)[10]:.3f}); wall incl. final sync {1e3*t_all/20:.3f} ms per step")
