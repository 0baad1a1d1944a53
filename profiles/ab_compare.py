"""A/B two builds of libinerf_b200.so on the same GPU, alternating, same inputs: the fine pass of a frame (202 500 x 192 points,
inerf_mlp_fwd) and the training forward of an N_rand = 3072 step (inerf_mlp_fwd_train_bf16).
    python profiles/ab_compare.py build/libinerf_a.so build/libinerf_b.so        (each run happens in its own process: INERF_SO)"""
import os
import subprocess
import sys

if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for rnd in range(3):
        for so in sys.argv[1:]:
            env = dict(os.environ, INERF_SO=os.path.abspath(so))
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
            print(f"round {rnd} {os.path.basename(so):28s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
    sys.exit(0)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ideal_nerf_b200 as M
from ideal_nerf_b200 import ops, synthetic as S

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
f = net.face_nerf_fine


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


with torch.no_grad():
    ms_r = timed(lambda: f.query(rays, z, aud, expr, lat), 5)
kp = f.kernel_params()
params = [p.detach() for p in kp]
packed = f.packed_weights(kp)
cond = ops.fold_cond(f._dims, params, aud, expr, lat)
r3, z3 = rays[:3072].contiguous(), z[:3072].contiguous()
ms_t = timed(lambda: ops.mlp_fwd_train_bf16(f._dims, params, packed, cond, r3, z3), 20)
print(f"fine pass {ms_r:7.3f} ms   training forward (3072 x 192) {ms_t:6.3f} ms")
