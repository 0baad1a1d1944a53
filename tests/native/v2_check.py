"""Debug driver (GPU): cta_group::2 inference kernel (INERF_MLP_V2=1) against the v1 kernel, then timing on a frame's fine pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S
from oracle import render_oracle as O

dev = torch.device("cuda", 0)
b = O.synthetic_train_batch(0)
sd = O.init_face_nerf(7)
net = M.FaceNeRF(dim_aud=64, dim_latent=32, dim_expr=76, mlp_mode="bf16"); net.load_state_dict(sd); net = net.to(dev)
aud, expr, lat = b["aud"].to(dev), b["expr"].to(dev), b["latent"].to(dev)
def run(rays, z, v2):
    if v2: os.environ["INERF_MLP_V2"] = "1"
    else: os.environ.pop("INERF_MLP_V2", None)
    with torch.no_grad():
        return net.query(rays, z, aud, expr, lat)
for n, s in ((8, 64), (301, 64), (40, 192), (47, 45), (1000, 192)):
    rays = b["rays"][:n].to(dev)
    z = ops.sample_coarse(rays, s, torch.rand(n, s, device=dev))
    r1 = run(rays, z, False); r2 = run(rays, z, True)
    torch.cuda.synchronize()
    print(f"n={n} s={s}: v2 vs v1 max-abs {float((r1 - r2).abs().max()):.3e}  equal={bool(torch.equal(r1, r2))}  finite={bool(torch.isfinite(r2).all())}", flush=True)
cam, fr = S.camera(), S.frame_inputs(0)
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
for v2 in (False, True, False, True):
    for _ in range(2): run(rays, z, v2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = run(rays, z, v2)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{'v2' if v2 else 'v1'}: fine pass {ms:.3f} ms  {rays.shape[0] * 192 * 1121280 / ms / 1e9:.1f} TFLOP/s", flush=True)

# phase timings of v2 (trace buffer = profiling output)
os.environ["INERF_MLP_V2"] = "1"
params = [p.detach() for p in net.kernel_params()]
cond = ops.fold_cond(net._dims, params, aud, expr, lat)
packed = net.packed_weights(net.kernel_params())
raw, tr = ops.mlp_fwd_trace(M._lib.INERF_MLP_BF16, net._dims, params, packed, cond, rays, z)
torch.cuda.synchronize()
t = tr.reshape(-1)[:32].cpu()
it = float(t[3])
print(f"issuer (pair 0): total {float(t[0])/it:.0f} cycles/iteration, wait-epilogue {float(t[1])/it:.0f}, wait-weights/bias {float(t[2])/it:.0f}")
for i, nm in enumerate(("leader slot0", "leader slot1", "peer slot0", "peer slot1")):
    e = t[8 + 4 * i: 12 + 4 * i]
    print(f"  epilogue {nm}: per iteration wait-C {float(e[0])/it:.0f}  ld+convert+store {float(e[1])/it:.0f}  fence+arrive {float(e[2])/it:.0f}")
