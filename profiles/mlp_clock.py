"""Effective SM clock of the bf16 FaceNeRF kernel: issuer-thread clock64 span (trace build) over the CUDA-event time of the same launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S
dev = torch.device("cuda", 0)
cam, fr = S.camera(), S.frame_inputs(0)
a = M.default_args(dim_aud=64, dim_expr=76, perturb=1.0, mlp_mode="bf16", N_samples=64, N_importance=128, near=S.NEAR, far=S.FAR)
net = M.Network(450, 450, cam["focal"], S.NEAR, S.FAR, 1 << 20, None, 64, 128, args=a)
torch.manual_seed(1); net.apply(M.init_weights); net = net.to(dev).eval()
rays = ops.get_rays_packed(450, 450, cam["focal"], cam["c2w"].to(dev), S.NEAR, S.FAR)
z = torch.sort(S.NEAR + (S.FAR - S.NEAR) * torch.rand(rays.shape[0], 192, device=dev), -1)[0].contiguous()
aud, expr, lat = fr["aud"].to(dev), fr["expr"].to(dev), fr["latent"].to(dev)
f = net.face_nerf_fine
with torch.no_grad():
    params = [p.detach() for p in f.kernel_params()]
    cond = ops.fold_cond(f._dims, params, aud, expr, lat)
    packed = f.packed_weights(f.kernel_params())
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.mlp_fwd_trace(M._lib.INERF_MLP_BF16, f._dims, params, packed, cond, rays, z)
        e1.record()
        torch.cuda.synchronize()
        t = ops.mlp_fwd_trace.timing.cpu()
        ms = e0.elapsed_time(e1)
        cyc = float(t[:148, 0].max())
        it = float(t[:148, 4].mean())
        print(f"trace launch {rep}: {ms:.3f} ms, issuer span {cyc / 1e6:.2f} Mcycles -> {cyc / ms / 1e3:.0f} MHz; per iteration: total {float(t[:148,0].mean())/it:.0f} "
              f"wait-E {float(t[:148,1].mean())/it:.0f} wait-PE {float(t[:148,2].mean())/it:.0f} wait-W {float(t[:148,3].mean())/it:.0f} cycles")
        e = t[148:296]
        print(f"   epilogue warp 4, per iteration (22 half-layers): wait-C0/C2 {float(e[:,0].mean())/it:.0f}  tmem-ld+convert {float(e[:,1].mean())/it:.0f}  "
              f"wait-C1 {float(e[:,2].mean())/it:.0f}  st.shared+fence+arrive {float(e[:,3].mean())/it:.0f} cycles")
        lw = ops.mlp_fwd_trace.layer_waits.cpu()[:, :11].mean(0) / it
        print("   issuer wait-E per layer (L0..L7, V0..V2):", " ".join(f"{float(x):.0f}" for x in lw))
